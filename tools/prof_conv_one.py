import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200.engine import Engine, Act, new_act, pack_conv2d
dev = torch.device("cuda", 0)
cin, cout, k, stride, (b, h, w) = 192, 192, 5, 2, (8, 256, 384)  # the launch bench.py's roofline times
m = nn.Conv2d(cin, cout, k, stride, k // 2)
pc = pack_conv2d(m, dev, "x").attach_tc(3)
x = Act(torch.randn(b, h, w, cin, device=dev)); out = new_act(b, h // stride, w // stride, cout, dev)
E = Engine(dev, 2)
for _ in range(3): E.conv(pc, [x], out)
torch.cuda.synchronize(); print("ok")
