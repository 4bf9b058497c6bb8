import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from progressivecodec_b200 import GaussianConditional, ans, get_scale_table
t = GaussianConditional(None)  # the codec's own 64-level tables (update() through the C-ABI quantiser)
t.update_scale_table(get_scale_table())
tables = ans.CdfTables(t._quantized_cdf, t._cdf_length, t._offset)
g = torch.Generator(device="cuda").manual_seed(0)
S, N = 8, 49152
sigma = torch.exp(torch.empty((S, N), device="cuda").uniform_(-3.0, 2.0, generator=g))
idx = torch.bucketize(sigma.clamp_min(0.11), t.scale_table.cuda()[:-1]).int()
sym = torch.round(torch.randn((S, N), generator=g, device="cuda") * sigma).int()
for _ in range(2):
    data, offs = ans.encode_batch(sym, idx, tables)
    out = ans.decode_batch(data, offs.cuda(), idx, tables)
torch.cuda.synchronize()
assert torch.equal(out, sym)
print("ok")
