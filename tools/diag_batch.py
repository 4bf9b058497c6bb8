#!/usr/bin/env python
"""Is image b of a large batch coded exactly like the same image alone?  python tools/diag_batch.py [--impl 0]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights
from progressivecodec_b200.synthetic import synthetic_image

ap = argparse.ArgumentParser()
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--batches", default="8,32,64")
args = ap.parse_args()
AUTHORS = dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True, delta_encode=True,
               support_progressive_slices=5, mask_policy="point-based-std")
net = ChannelProgresssiveWACNN(**AUTHORS).eval()
apply_synthetic_weights(net, seed=0)
net.update(force=True)
net = net.cuda()
net.prepare()["eng"].conv_impl = args.impl
x = torch.cat([synthetic_image((1, 3, 512, 768), seed=200 + i) for i in range(64)]).cuda()
q = 0.5
single = {}
for b in (0, 3, 61):
    d = {}
    net.compress(x[b:b + 1], quality=q, debug=d)
    single[b] = {k: v.clone() for k, v in d.items()}
for B in [int(v) for v in args.batches.split(",")]:
    d = {}
    net.compress(x[:B], quality=q, debug=d)
    for b in (0, 3, 61):
        if b >= B:
            continue
        s = single[b]
        ydiff = (d["y"][b] - s["y"][0]).abs().max().item()
        zdiff = int((d["z_symbols"][b] != s["z_symbols"][0]).sum())
        sdiff = [int((d["symbols"][k, b] != s["symbols"][k, 0]).sum()) for k in range(d["symbols"].shape[0])]
        print(f"B={B} image {b}: max|dy|={ydiff:.3e} z symbols differing={zdiff} y symbols differing per slice={sdiff}", flush=True)
