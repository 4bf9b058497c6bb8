import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200.engine import Act, Engine, pack_conv2d
dev = torch.device("cuda", 0)
B = int(sys.argv[1]); cin, cout = int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(1)
m = nn.Conv2d(cin, cout, 3, 1, 1)
x = torch.randn(B, cin, 32, 48)
E = Engine(dev, 3)
pc = pack_conv2d(m, dev, "t").attach_tc(3)
torch.cuda.synchronize(); print("weights ok", flush=True)
xa = Act(x.permute(0, 2, 3, 1).contiguous().cuda())
E.planes(xa); torch.cuda.synchronize(); print("planes ok", flush=True)
out = E.conv_new(pc, [xa]); torch.cuda.synchronize(); print("conv ok", flush=True)
ref = torch.nn.functional.conv2d(x[:1].double(), m.weight.double(), m.bias.double(), 1, 1)
got = out.t[:1].permute(0, 3, 1, 2).double().cpu()
print("rms err", ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item())
