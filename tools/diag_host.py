import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import AUTHORS, H, W
from progressivecodec_b200.synthetic import synthetic_image
from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights
net = ChannelProgresssiveWACNN(**AUTHORS).eval(); apply_synthetic_weights(net, seed=0); net.update(force=True); net = net.cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.cat([synthetic_image((1, 3, H, W), seed=i) for i in range(B)]).cuda()
for G in (1, 8):
    net.decode_groups = G
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        c = net.compress(x, quality=5, return_device_streams=True)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        r = net.decompress(c, c["shape"], quality=5)["x_hat"]
        t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"B={B} G={G}: compress host {1e3*(t1-t0):.1f} ms (+{1e3*(t2-t1):.1f} wait) | decompress host {1e3*(t3-t2):.1f} ms (+{1e3*(t4-t3):.1f} wait)  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
