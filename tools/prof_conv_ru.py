"""A bandwidth-class launch (ResidualUnit 1x1 96->192 + residual + GELU at 32 x 128x192): for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200 import _lib as L
from progressivecodec_b200.engine import Engine, Act, new_act, pack_conv2d
dev = torch.device("cuda", 0)
m = nn.Conv2d(96, 192, 1)
pc = pack_conv2d(m, dev, "ru_c3").attach_tc(3)
x = Act(torch.randn(32, 128, 192, 96, device=dev)); out = new_act(32, 128, 192, 192, dev)
r1 = Act(torch.randn(32, 128, 192, 192, device=dev))
E = Engine(dev, 2)
for _ in range(3): E.conv(pc, [x], out, L.EPI_ADD_GELU, r1)
torch.cuda.synchronize(); print("ok")
