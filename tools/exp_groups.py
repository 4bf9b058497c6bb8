"""Pipelined sweep throughput vs number of decode groups (image groups decoded on separate streams)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import AUTHORS, H, W, QUALITIES
from progressivecodec_b200.synthetic import synthetic_image
from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights, pipeline
net = ChannelProgresssiveWACNN(**AUTHORS).eval(); apply_synthetic_weights(net, seed=0); net.update(force=True); net = net.cuda()
B = int(os.environ.get("B", "32"))
x = torch.cat([synthetic_image((1, 3, H, W), seed=i) for i in range(B)]).cuda()
for g in (1, 2, 4, 8):
    net.decode_groups = g
    pipeline.sweep(net, x, QUALITIES, keep=False); torch.cuda.synchronize()
    t0 = time.perf_counter(); pipeline.sweep(net, x, QUALITIES, keep=False); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"groups {g}: {B * len(QUALITIES) / t:7.1f} image-qualities/s", flush=True)
