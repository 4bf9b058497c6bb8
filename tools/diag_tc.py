import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200.engine import Engine, Act, pack_conv2d

def nhwc(x): return Act(x.permute(0, 2, 3, 1).contiguous().cuda())
def nchw(a): return a.t[..., a.c0:a.c0 + a.C].permute(0, 3, 1, 2).cpu()
E2, E1 = Engine(torch.device("cuda", 0), 2), Engine(torch.device("cuda", 0), 1)
for (cin, cout, k, stride, hw) in [(192, 96, 1, 1, (16, 16)), (32, 32, 1, 1, (16, 16)), (64, 32, 3, 1, (16, 16)), (192, 192, 3, 1, (16, 16)),
                                   (192, 192, 5, 2, (32, 48)), (320, 640, 5, 2, (8, 8))]:
    torch.manual_seed(1)
    m = nn.Conv2d(cin, cout, k, stride, k // 2)
    x = torch.randn(2, cin, *hw)
    ref = torch.nn.functional.conv2d(x.double(), m.weight.double(), m.bias.double(), stride, k // 2)
    scale = ref.pow(2).mean().sqrt().item()
    pc = pack_conv2d(m, E2.device, "t").attach_tc(3)
    r3 = nchw(E2.conv_new(pc, [nhwc(x)])).double()
    pc.tc_split = 1
    r1 = nchw(E2.conv_new(pc, [nhwc(x)])).double()
    rs = nchw(E1.conv_new(pc, [nhwc(x)])).double()
    # variants: zero the low bits of the input on the host to see what the HW does with them
    xt = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    pc.tc_split = 1
    r1t = nchw(E2.conv_new(pc, [nhwc(xt)])).double()
    reft = torch.nn.functional.conv2d(xt.double(), m.weight.double(), m.bias.double(), stride, k // 2)
    f = lambda a, b: ((a - b).abs().max().item() / scale, ((a - b).pow(2).mean().sqrt().item()) / scale)
    print(f"K={cin*k*k:5d} cout={cout}: split3 max/rms {f(r3, ref)}  split1 {f(r1, ref)}  simt {f(rs, ref)}  split1(trunc-x vs trunc-ref) {f(r1t, reft)}  mean signed err3 {((r3-ref).mean().item())/scale:.2e}")
