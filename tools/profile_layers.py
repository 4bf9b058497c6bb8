"""Per-layer time of every tap-GEMM launch of one compress+decompress (CUDA events around each launch, serialised).

    python tools/profile_layers.py [--batch B] [--quality Q]

Prints one row per distinct conv shape: launches, total ms, share, algorithmic TFLOP/s.
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import AUTHORS, H, W
from progressivecodec_b200.synthetic import synthetic_image
from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights
from progressivecodec_b200 import engine as eng

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--quality", type=float, default=5)
a = ap.parse_args()
net = ChannelProgresssiveWACNN(**AUTHORS).eval()
apply_synthetic_weights(net, seed=0)
net.update(force=True)
net = net.cuda()
net.decode_groups = 1
x = torch.cat([synthetic_image((1, 3, H, W), seed=i) for i in range(a.batch)]).cuda()
c = net.compress(x, quality=a.quality, return_device_streams=True)
net.decompress(c, c["shape"], quality=a.quality)
torch.cuda.synchronize()

records = []
orig = eng.Engine.conv


def timed(self, pc, segs, out, epi=0, r1=None, r2=None, flags=0, fmt=3, square_planes=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()  # (includes the split-plane conversions this launch triggers for inputs no epilogue produced)
    r = orig(self, pc, segs, out, epi, r1, r2, flags, fmt, square_planes)
    e1.record()
    a0 = segs[0]
    if pc.out_step == 1:
        M = a0.B * (a0.H // pc.in_step) * (a0.W // pc.in_step)
    else:
        M = a0.B * a0.H * a0.W
    records.append((pc.name, len(pc.taps), pc.cin, pc.cout, pc.in_step, pc.out_step, M, epi, pc.tc_split, e0, e1))
    return r


eng.Engine.conv = timed
phase_marks = {}
for phase, fn in (("compress", lambda: net.compress(x, quality=a.quality, return_device_streams=True)),
                  ("decompress", lambda: net.decompress(c, c["shape"], quality=a.quality))):
    n0 = len(records)
    fn()
    torch.cuda.synchronize()
    phase_marks[phase] = (n0, len(records))
eng.Engine.conv = orig

for phase, (lo, hi) in phase_marks.items():
    agg = collections.OrderedDict()
    tot = 0.0
    for name, taps, cin, cout, ins, outs, M, epi, split, e0, e1 in records[lo:hi]:
        ms = e0.elapsed_time(e1)
        fam = name.split(".")[0]
        key = (fam, taps, cin, cout, ins, outs, M, epi, split)
        v = agg.setdefault(key, [0, 0.0, 0.0])
        v[0] += 1
        v[1] += ms
        v[2] += 2.0 * M * taps * cin * cout
        tot += ms
    print(f"== {phase}: {hi - lo} conv launches, {tot:.1f} ms (event-timed one by one)")
    print(f"{'family':26s} {'taps':>4s} {'cin':>4s} {'cout':>4s} {'is':>2s} {'os':>2s} {'M':>8s} {'epi':>3s} {'n':>4s} "
          f"{'ms':>8s} {'share':>6s} {'us/launch':>9s} {'TF/s':>7s}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fam, taps, cin, cout, ins, outs, M, epi, split = key
        print(f"{fam:26s} {taps:4d} {cin:4d} {cout:4d} {ins:2d} {outs:2d} {M:8d} {epi:3d} {v[0]:4d} {v[1]:8.2f} "
              f"{100 * v[1] / tot:5.1f}% {1e3 * v[1] / v[0]:9.1f} {v[2] / v[1] / 1e9:7.1f}")
