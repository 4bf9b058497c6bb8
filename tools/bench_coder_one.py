import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import entropy_port as EP
from progressivecodec_b200 import ans
t = EP.GaussianTables.build()
tables = ans.CdfTables(t.cdf, t.cdf_length, t.offset)
g = torch.Generator(device="cuda").manual_seed(0)
S, N = 8, 49152
sigma = torch.exp(torch.empty((S, N), device="cuda").uniform_(-3.0, 2.0, generator=g))
idx = torch.bucketize(sigma.clamp_min(0.11), t.scale_table.cuda()[:-1]).int()
sym = torch.round(torch.randn((S, N), generator=g, device="cuda") * sigma).int()
for _ in range(2):
    data, offs = ans.encode_batch(sym, idx, tables)
    out = ans.decode_batch(data, offs.cuda(), idx, tables)
torch.cuda.synchronize()
assert torch.equal(out, sym)
print("ok")
