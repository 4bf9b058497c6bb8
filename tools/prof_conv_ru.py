"""A bandwidth-class launch as the model issues it (ResidualUnit 1x1 96->192 + residual + GELU at 32 x 128x192 on the fp16
kernel: planes in, plane residual, planes out): for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200 import _lib as L
from progressivecodec_b200.engine import Engine, Act, pack_conv2d
dev = torch.device("cuda", 0)
m = nn.Conv2d(96, 192, 1)
pc = pack_conv2d(m, dev, "ru_c3").attach_tc(3)
E = Engine(dev, int(os.environ.get("IMPL", "3")))
x = Act(torch.randn(32, 128, 192, 96, device=dev)); E.planes(x)
r1 = Act(torch.randn(32, 128, 192, 192, device=dev)); E.planes(r1)
xp = Act(base=0, shape=(32, 128, 192, 96)); xp._root = x._root      # planes-only views of the same buffers
rp = Act(base=0, shape=(32, 128, 192, 192)); rp._root = r1._root
out = E.act(32, 128, 192, 192, fmt=2)
try:
    for _ in range(3): E.conv(pc, [xp], out, L.EPI_ADD_GELU, rp, fmt=2)
except Exception as e:  # fall back to the fp32-backed operands if the view construction above is not accepted
    print("planes-only views rejected:", e)
    out = E.act(32, 128, 192, 192, fmt=2)
    for _ in range(3): E.conv(pc, [x], out, L.EPI_ADD_GELU, r1, fmt=2)
torch.cuda.synchronize(); print("ok")
