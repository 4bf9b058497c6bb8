"""cout <= 64 layers: one-CTA-per-SM (PCODEC_TC_SMALL=0) vs two-CTAs-per-SM variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200 import _lib as L
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
E = Engine(dev, 2)
for name, cin, cout, k, (b, h, w), epi in [("L4 128->64", 128, 64, 3, (32, 32, 48), L.EPI_GELU), ("L5 64->32", 64, 32, 3, (32, 32, 48), L.EPI_LINEAR),
                                           ("L4 b8", 128, 64, 3, (8, 32, 48), L.EPI_GELU), ("L5 b8", 64, 32, 3, (8, 32, 48), L.EPI_LINEAR)]:
    row = []
    for small in ("0", "1"):
        os.environ["PCODEC_TC_SMALL"] = small
        m = nn.Conv2d(cin, cout, k, 1, k // 2)
        pc = pack_conv2d(m, dev, name).attach_tc(3)
        x = Act(torch.randn(b, h, w, cin, device=dev)); out = new_act(b, h, w, cout, dev)
        t = timeit(lambda: E.conv(pc, [x], out, epi))
        row.append(f"small={small}: {t*1e3:7.1f} us")
    print(f"{name:14s} " + " | ".join(row), flush=True)
