"""One 768x512 image through the 13-level sweep (BASELINE configs[1] taken literally): eager vs graph replay, by the
number of decode workers.  Prints image-qualities/s (13 levels / sweep time, CUDA events around the whole sweep)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import AUTHORS, QUALITIES
from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights, pipeline
from progressivecodec_b200.synthetic import synthetic_image

net = ChannelProgresssiveWACNN(**AUTHORS).eval()
apply_synthetic_weights(net, seed=0)
net.update(force=True)
net = net.cuda()
x = synthetic_image((1, 3, 512, 768), seed=19).cuda()
combos = [(g, w, 3) for g in (False, True) for w in (4, 6, 8)]
if len(sys.argv) > 1:  # graphs,workers,encoders ...
    combos = [(bool(int(a.split(",")[0])), int(a.split(",")[1]), int(a.split(",")[2])) for a in sys.argv[1:]]
for graphs, workers, encoders in combos:
    for _ in range(3):
        pipeline.sweep(net, x, QUALITIES, graphs=graphs, decode_workers=workers, encoders=encoders, keep=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        pipeline.sweep(net, x, QUALITIES, graphs=graphs, decode_workers=workers, encoders=encoders, keep=False)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"graphs={graphs} workers {workers} encoders {encoders}: sweep {1e3 * ts[2]:.0f} ms = {len(QUALITIES) / ts[2]:.1f} image-qualities/s "
          f"(min {1e3 * ts[0]:.0f} ms)", flush=True)
# single calls at q = 5
for q in (5,):
    c = net.compress(x, quality=q, return_device_streams=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        c = net.compress(x, quality=q, return_device_streams=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(5):
        net.decompress(c, c["shape"], quality=q)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"eager q={q}: compress {200 * (t1 - t0):.1f} ms, decompress {200 * (t2 - t1):.1f} ms", flush=True)
