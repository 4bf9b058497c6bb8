"""The launch bench.py's roofline times (5x5 s2 192->192, 8 x 256x384 -> 128x192, fp32 output only), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200.engine import Engine, Act, pack_conv2d
dev = torch.device("cuda", 0)
cin, cout, k, stride, (b, h, w) = 192, 192, 5, 2, (8, 256, 384)
m = nn.Conv2d(cin, cout, k, stride, k // 2)
pc = pack_conv2d(m, dev, "x").attach_tc(3)
E = Engine(dev, int(os.environ.get("IMPL", "3")))
x = Act(torch.randn(b, h, w, cin, device=dev))
E.planes(x)
out = E.act(b, h // stride, w // stride, cout, fmt=1)
for _ in range(3): E.conv(pc, [x], out, fmt=1)
torch.cuda.synchronize(); print("ok")
