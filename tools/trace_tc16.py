#!/usr/bin/env python
"""clock64 timeline of one CTA of the fp16 conv kernel (needs the PCODEC_EXPERIMENTS build, see tools/exp_tc16.sh)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PCODEC_TC16_DEBUG"] = os.environ.get("PCODEC_TC16_DEBUG", "192")
import torch
import torch.nn as nn

from progressivecodec_b200 import _lib as L
from progressivecodec_b200.engine import Act, Engine, new_act, pack_conv2d

dev = torch.device("cuda", 0)
for cin, cout, k, hw, B, epi in ((512, 224, 3, (32, 48), 32, L.EPI_GELU), (192, 96, 1, (128, 192), 8, L.EPI_GELU),
                                 (64, 32, 3, (32, 48), 32, L.EPI_LINEAR)):
    m = nn.Conv2d(cin, cout, k, 1, k // 2)
    pc = pack_conv2d(m, dev, "t").attach_tc(3)
    x = Act(torch.randn(B, hw[0], hw[1], cin, device=dev))
    E = Engine(dev, 3)
    out = E.act(B, hw[0], hw[1], cout)
    for _ in range(3):
        E.conv(pc, [x], out, epi)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 128)()
    lib = ctypes.CDLL(L.LIB_PATH)
    assert lib.pcodec_debug_tc16_trace(buf, 128) == 128
    t = list(buf)
    t0 = t[0]
    names = ["entry", "setup done", "first stage landed", "last MMA committed", "accumulators complete", "epilogue done", "exit"]
    print(f"== {k}x{k} {cin}->{cout} @{hw} B={B}")
    for i, n in enumerate(names):
        print(f"   {n:24s} {t[i] - t0:8d} clk")
    slabs = [v - t0 for v in t[8:108] if v > t0]
    if len(slabs) > 4:
        d = [b - a for a, b in zip(slabs, slabs[1:])]
        print(f"   slab issue period: first {d[:6]} ... median {sorted(d)[len(d) // 2]} max {max(d)} (n={len(d)})")
