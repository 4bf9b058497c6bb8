"""The launch class that carries most of the step (slice-net first layer, 3x3 512->224 at 32 x 32x48): for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200 import _lib as L
from progressivecodec_b200.engine import Engine, Act, new_act, pack_conv2d
dev = torch.device("cuda", 0)
m = nn.Conv2d(512, 224, 3, 1, 1)
pc = pack_conv2d(m, dev, "cc_l1").attach_tc(3)
x = Act(torch.randn(32, 32, 48, 512, device=dev)); out = new_act(32, 32, 48, 224, dev)
E = Engine(dev, int(os.environ.get("IMPL", "3")))
for _ in range(3): E.conv(pc, [x], out, L.EPI_GELU)
torch.cuda.synchronize(); print("ok")
