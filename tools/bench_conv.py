"""Time the tap-GEMM conv kernels alone on representative layer shapes (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200.engine import Engine, Act, new_act, pack_conv2d

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=8):
    fn(); fn(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
B = int(os.environ.get("B", "8"))
shapes = [("g_a 5x5s2 192->192 @256x384", 192, 192, 5, 2, (1, 256, 384)),
          ("cc L1 3x3 352->224 @32x48", 352, 224, 3, 1, (B, 32, 48)),
          ("cc L2 3x3 224->176", 224, 176, 3, 1, (B, 32, 48)),
          ("cc L3 3x3 176->128", 176, 128, 3, 1, (B, 32, 48)),
          ("cc L5 3x3 64->32", 64, 32, 3, 1, (B, 32, 48)),
          ("ru 1x1 192->96 @128x192", 192, 96, 1, 1, (B, 128, 192)),
          ("ru 3x3 96->96 @128x192", 96, 96, 3, 1, (B, 128, 192))]
for name, cin, cout, k, stride, (b, h, w) in shapes:
    m = nn.Conv2d(cin, cout, k, stride, k // 2)
    pc = pack_conv2d(m, dev, name).attach_tc(3)
    x = Act(torch.randn(b, h, w, cin, device=dev))
    out = new_act(b, h // stride, w // stride, cout, dev)
    flops = 2.0 * b * (h // stride) * (w // stride) * cout * cin * k * k
    res = []
    for impl, split in ((2, 3), (2, 1), (1, 3)):
        E = Engine(dev, impl); pc.tc_split = split
        t = timeit(lambda: E.conv(pc, [x], out))
        res.append(f"{'tc' if impl == 2 else 'simt'}{split if impl == 2 else ''}: {t*1e3:8.1f} us {flops/t/1e9:7.1f} TF/s")
    print(f"{name:32s} M={b*(h//stride)*(w//stride):7d} " + " | ".join(res), flush=True)
