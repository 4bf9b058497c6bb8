"""Slope/intercept of conv_taps_tc_kernel time vs K length: time per CTA = F + n_steps * c."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200 import _lib as L
from progressivecodec_b200.engine import Engine, Act, new_act, pack_conv2d

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=5):
    fn(); fn(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
B = 37  # 444 M-tiles = 3 x 148
E = Engine(dev, 2)
modes = [int(v) for v in os.environ.get("MODES", "0,7").split(",")]
for cout in (224, 176, 128, 64, 32):
    for k in (1, 3):
        for cin in (64, 128, 256, 512):
            m = nn.Conv2d(cin, cout, k, 1, k // 2)
            pc = pack_conv2d(m, dev, "x").attach_tc(3)
            x = Act(torch.randn(B, 32, 48, cin, device=dev))
            out = new_act(B, 32, 48, cout, dev)
            ntiles = (B * 12) * (2 if cout > 128 else 1)
            waves = -(-ntiles // 148)
            slabs = k * k * ((cin + 31) // 32)
            row = []
            for dbg in modes:
                os.environ["PCODEC_TC_DEBUG"] = str(dbg)
                for split in (3, 1):
                    pc.tc_split = split
                    t = timeit(lambda: E.conv(pc, [x], out, L.EPI_GELU))
                    row.append(f"d{dbg}s{split}: {t*1e3:7.1f} us ({t*1e3/waves:6.1f}/cta)")
            os.environ["PCODEC_TC_DEBUG"] = "0"
            print(f"cout {cout:3d} k{k} cin {cin:3d} slabs {slabs:3d} waves {waves} | " + " | ".join(row), flush=True)
