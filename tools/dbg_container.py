import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from conftest import build_pair
from progressivecodec_b200.synthetic import synthetic_image
from progressivecodec_b200 import container as C
net, orc = build_pair("allscalable", "cuda")
x = synthetic_image((2, 3, 128, 192), seed=21)
LEVELS = (0.05, 0.5, 1.25, 5, 10)
blobs = C.encode_progressive(net, x.cuda(), LEVELS)
for k, q in enumerate(LEVELS, start=1):
    dbg = {}
    out = C.decode_progressive(net, [C.truncate(b, k) for b in blobs], debug=dbg)
    f = net.forward_single_quality(x.cuda(), q, training=False)
    yh = f["y_hat"]
    d = (dbg["y_hat"] - yh).abs()
    print("level", q, "y_hat maxdiff", float(d.max()), "n diff", int((d > 1e-6).sum()), "per-slice", [int((d[:, 32*i:32*i+32] > 1e-6).sum()) for i in range(10)])
    dbg2 = {}
    c = net.compress(x.cuda(), quality=q, debug=dbg2)
    sym = dbg2["symbols"][10:]  # prog
    ypre = torch.stack(dbg["y_pre"], 0)  # [10,B,32,h,w]
    mu = f["mu"]  # [B,320,h,w]
    rec_sym = (ypre.permute(1,0,2,3,4).reshape(mu.shape) - mu).round().int()
    enc_sym = sym.reshape(10, 2, 32, mu.shape[2], mu.shape[3]).permute(1,0,2,3,4).reshape(mu.shape)
    print("   symbols differ:", int((rec_sym != enc_sym).sum()), "nonzero enc", int((enc_sym != 0).sum()), "nonzero dec", int((rec_sym != 0).sum()))
