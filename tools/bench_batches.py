"""Throughput of compress / decompress (device-resident) vs batch size and decode group count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import AUTHORS, H, W
from progressivecodec_b200.synthetic import synthetic_image
from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights

net = ChannelProgresssiveWACNN(**AUTHORS).eval()
apply_synthetic_weights(net, seed=0)
net.update(force=True)
net = net.cuda()
q = 5
configs = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]] or [(8, 1), (8, 2), (16, 2), (16, 4), (32, 4)]
for B, G in configs:
    x = torch.cat([synthetic_image((1, 3, H, W), seed=i) for i in range(B)]).cuda()
    net.decode_groups = G
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        c = net.compress(x, quality=q, return_device_streams=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        r = net.decompress(c, c["shape"], quality=q)["x_hat"]
        torch.cuda.synchronize(); t2 = time.perf_counter()
    ref = None
    if G > 1:
        net.decode_groups = 1
        ref = net.decompress(c, c["shape"], quality=q)["x_hat"]
        same = torch.equal(ref, r)
    else:
        same = True
    print(f"B={B:3d} groups={G}: compress {1e3*(t1-t0):7.1f} ms  decompress {1e3*(t2-t1):7.1f} ms  -> {B/(t2-t0):6.1f} img-q/s  "
          f"(grouped == single-stream: {same}; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB)", flush=True)
    del x, c, r, ref
    torch.cuda.empty_cache()
