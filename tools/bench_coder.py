"""Time the rANS kernels alone (CUDA events): S streams of N symbols, realistic Gaussian symbol mix."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from progressivecodec_b200 import GaussianConditional, ans, get_scale_table

t = GaussianConditional(None)  # the codec's own 64-level tables (update() through the C-ABI quantiser)
t.update_scale_table(get_scale_table())
tables = ans.CdfTables(t._quantized_cdf, t._cdf_length, t._offset)
g = torch.Generator(device="cuda").manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
for S in (1, 2, 8, 32, 168, 672):
    # sigma log-uniform in [0.05, 8]: most symbols in the narrow tables, like the codec's slices
    sigma = torch.exp(torch.empty((S, N), device="cuda").uniform_(-3.0, 2.0, generator=g))
    idx = torch.bucketize(sigma.clamp_min(0.11), t.scale_table.cuda()[:-1]).int()
    sym = torch.round(torch.randn((S, N), generator=g, device="cuda") * sigma).int()
    data, offs = ans.encode_batch(sym, idx, tables)
    out = ans.decode_batch(data, offs, idx, tables)
    assert torch.equal(out, sym)
    offs_d = offs.cuda()
    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps
    te = timeit(lambda: ans.encode_batch(sym, idx, tables))
    td = timeit(lambda: ans.decode_batch(data, offs_d, idx, tables))
    bits = 8.0 * int(offs[-1]) / (S * N)
    print(f"S={S:4d} N={N}: encode {te:8.3f} ms  decode {td:8.3f} ms   ({bits:.2f} bit/sym; "
          f"decode {td * 1e6 / N:.1f} ns/sym/stream, {S * N / td / 1e3:.1f} Msym/s)")
