import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200.engine import Act, Engine, pack_conv2d
dev = torch.device("cuda", 0)
E = Engine(dev, 3)
for cin, cout in [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]:
    torch.manual_seed(1)
    m = nn.Conv2d(cin, cout, 3, 1, 1)
    x = torch.randn(3, cin, 32, 48)
    pc = pack_conv2d(m, dev, "t").attach_tc(3)
    xa = Act(x.permute(0, 2, 3, 1).contiguous().cuda())
    try:
        out = E.conv_new(pc, [xa]); torch.cuda.synchronize()
    except Exception as e:
        print(cin, cout, "FAILED", str(e).split("\n")[0]); break
    ref = torch.nn.functional.conv2d(x[:1].double(), m.weight.double(), m.bias.double(), 1, 1)
    got = out.t[:1].permute(0, 3, 1, 2).double().cpu()
    print(cin, cout, "ok rms err", ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(), flush=True)
