"""One compress+decompress of a 768x512 batch at one quality (after a warm-up) — the command profiled with ncu.

    python tools/profile_step.py [--batch B] [--quality Q] [--warmup N]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import AUTHORS, H, W
from progressivecodec_b200.synthetic import synthetic_image
from progressivecodec_b200 import ChannelProgresssiveWACNN, _lib, apply_synthetic_weights

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--quality", type=float, default=5)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
net = ChannelProgresssiveWACNN(**AUTHORS).eval()
apply_synthetic_weights(net, seed=0)
net.update(force=True)
net = net.cuda()
x = torch.cat([synthetic_image((1, 3, H, W), seed=i) for i in range(a.batch)]).cuda()
for _ in range(a.warmup):
    c = net.compress(x, quality=a.quality, return_device_streams=True)
    net.decompress(c, c["shape"], quality=a.quality)
torch.cuda.synchronize()
_lib.lib().pcodec_reset_launch_count()
torch.cuda.cudart().cudaProfilerStart()
t0 = time.perf_counter()
c = net.compress(x, quality=a.quality, return_device_streams=True)
torch.cuda.synchronize()
t1 = time.perf_counter()
net.decompress(c, c["shape"], quality=a.quality)
torch.cuda.synchronize()
t2 = time.perf_counter()
torch.cuda.cudart().cudaProfilerStop()
print(f"batch {a.batch} q {a.quality}: compress {1e3 * (t1 - t0):.1f} ms, decompress {1e3 * (t2 - t1):.1f} ms, "
      f"launches {_lib.lib().pcodec_launch_count()}")
