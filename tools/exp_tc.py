"""Bottleneck decomposition of conv_taps_tc_kernel with the PCODEC_TC_DEBUG knobs (results are WRONG under the
knobs; timing only).  bit0: no A global loads, bit1: no B TMA loads, bit2: no converter TMEM stores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200 import _lib as L
from progressivecodec_b200.engine import Engine, Act, new_act, pack_conv2d

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=6):
    fn(); fn(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
B = int(os.environ.get("B", "32"))
shapes = [("cc L1 3x3 512->224 @32x48", 512, 224, 3, 1, (B, 32, 48), L.EPI_GELU, False),
          ("cc L2 3x3 224->176", 224, 176, 3, 1, (B, 32, 48), L.EPI_GELU, False),
          ("cc L4 3x3 128->64", 128, 64, 3, 1, (B, 32, 48), L.EPI_GELU, False),
          ("ru 1x1 96->192 addgelu @128x192", 96, 192, 1, 1, (B, 128, 192), L.EPI_ADD_GELU, True),
          ("ru 1x1 192->96 gelu @128x192", 192, 96, 1, 1, (B, 128, 192), L.EPI_GELU, False),
          ("ru 3x3 96->96 @128x192", 96, 96, 3, 1, (B, 128, 192), L.EPI_GELU, False)]
for name, cin, cout, k, stride, (b, h, w), epi, res in shapes:
    m = nn.Conv2d(cin, cout, k, stride, k // 2)
    pc = pack_conv2d(m, dev, name).attach_tc(3)
    x = Act(torch.randn(b, h, w, cin, device=dev))
    out = new_act(b, h // stride, w // stride, cout, dev)
    r1 = Act(torch.randn(b, h, w, cout, device=dev)) if res else None
    flops = 2.0 * b * (h // stride) * (w // stride) * cout * cin * k * k
    E = Engine(dev, 2)
    row = []
    for dbg in [int(v) for v in os.environ.get('MODES', '0,1,2,3,4,7').split(',')]:
        os.environ["PCODEC_TC_DEBUG"] = str(dbg)
        for split in (3, 1):
            pc.tc_split = split
            t = timeit(lambda: E.conv(pc, [x], out, epi, r1))
            row.append(f"d{dbg}s{split}:{t*1e3:7.1f}")
    os.environ["PCODEC_TC_DEBUG"] = "0"
    print(f"{name:34s} {flops/1e9:7.1f} GF | " + " ".join(row), flush=True)
