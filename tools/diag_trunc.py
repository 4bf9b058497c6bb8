"""Does kind::tf32 truncate (ignore the low 13 mantissa bits of) fp32 operands?  Compare plain-TF32 results on raw
inputs against the same inputs with the low bits cleared / rounded on the host."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200.engine import Engine, Act, pack_conv2d
def nhwc(x): return Act(x.permute(0, 2, 3, 1).contiguous().cuda())
def nchw(a): return a.t[..., a.c0:a.c0 + a.C].permute(0, 3, 1, 2).cpu()
E = Engine(torch.device("cuda", 0), 2)
torch.manual_seed(0)
m = nn.Conv2d(192, 96, 3, 1, 1)
pc = pack_conv2d(m, E.device, "t").attach_tc(1)
x = torch.randn(2, 192, 16, 16)
xi = x.view(torch.int32)
x_trunc = (xi & ~0x1FFF).view(torch.float32)
x_rne = ((xi + 0xFFF + ((xi >> 13) & 1)) & ~0x1FFF).view(torch.float32)
r_raw, r_tr, r_rn = (nchw(E.conv_new(pc, [nhwc(t)])) for t in (x, x_trunc, x_rne))
print("raw vs truncated-on-host: identical =", torch.equal(r_raw, r_tr), " max diff", (r_raw - r_tr).abs().max().item())
print("raw vs RNE-on-host      : identical =", torch.equal(r_raw, r_rn), " max diff", (r_raw - r_rn).abs().max().item())
