import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import AUTHORS
from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights, graphs
from progressivecodec_b200.synthetic import synthetic_image
net = ChannelProgresssiveWACNN(**AUTHORS).eval(); apply_synthetic_weights(net, seed=0); net.update(force=True); net = net.cuda()
x = synthetic_image((1, 3, 128, 192), seed=19).cuda()
st = torch.cuda.Stream()
for q in (0, 0.05, 10):
    ref = net.compress(x, quality=q, _planes_only=True)
    ref = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ref.items()}
    c_ref = net.compress(x, quality=q, return_device_streams=True)
    xr = net.decompress(c_ref, c_ref["shape"], quality=q)["x_hat"].clone()
    ge = graphs.GraphedCompress(net, tuple(x.shape), q, None, st)
    with torch.cuda.stream(st):
        for rep in range(2):
            c = ge(x, return_device_streams=True)
            st.synchronize()
            print(f"q={q} rep {rep}: sym equal {torch.equal(ge.planes['sym'], ref['sym'])} idx equal {torch.equal(ge.planes['idx'], ref['idx'])} "
                  f"z equal {torch.equal(ge.planes['z_sym'], ref['z_sym'])} bytes {int(c['streams'][1][-1])} vs {int(c_ref['streams'][1][-1])}")
    gd = graphs.GraphedDecompress(net, (2, 3), q, None, 1, net.ns0 if q <= 0 else net.ns1, 32 * 8 * 12, 0, st)
    with torch.cuda.stream(st):
        for rep in range(2):
            xh = gd(c_ref)
            st.synchronize()
            print(f"q={q} rep {rep}: decode equal {torch.equal(xh, xr)} max diff {(xh - xr).abs().max().item():.3e}")
