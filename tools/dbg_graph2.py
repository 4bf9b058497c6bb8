"""Replays test_pipelined_sweep_with_several_decode_workers followed by test_graphed_small_batch_sweep_equals_sequential_calls
on ONE model (the tests share a cached model) and reports every mismatch instead of stopping at the first."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import AUTHORS
from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights, pipeline
from progressivecodec_b200.synthetic import synthetic_image

net = ChannelProgresssiveWACNN(**AUTHORS).eval()
apply_synthetic_weights(net, seed=0)
net.update(force=True)
net = net.cuda()

x = synthetic_image((1, 3, 128, 192), seed=19).cuda()
qs = [0, 0.05, 0.5, 1.25, 5, 10]
ref = []
for q in qs:
    c = net.compress(x, quality=q)
    ref.append(net.decompress(c["strings"], c["shape"], quality=q)["x_hat"])
for workers in (None, 3):
    got = pipeline.sweep(net, x, qs, decode_workers=workers)
    for i in range(len(qs)):
        print(f"A workers={workers} q={qs[i]}: equal {torch.equal(got[i], ref[i])}", flush=True)

for batch in (1, 2):
    x = synthetic_image((batch, 3, 128, 192), seed=41).cuda()
    qs = [0, 0.05, 1.25, 10]
    seq = []
    for q in qs:
        c = net.compress(x, quality=q)
        seq.append((c["strings"], net.decompress(c["strings"], c["shape"], quality=q)["x_hat"]))
    for rep in range(3):
        x_in = x if rep < 2 else torch.flip(x, dims=[3]).contiguous()
        seen = {}
        got = pipeline.sweep(net, x_in, qs, graphs=True, host_strings=(rep == 1),
                             on_result=(lambda q, c, r: seen.__setitem__(q, c["strings"])) if rep == 1 else None)
        for i, q in enumerate(qs):
            if rep < 2:
                want = seq[i][1]
                extra = ""
                if rep == 1:
                    extra = f" y-strings equal {seen[q][0] == seq[i][0][0]} z equal {seen[q][1] == seq[i][0][1]}"
            else:
                c = net.compress(x_in, quality=q)
                want = net.decompress(c["strings"], c["shape"], quality=q)["x_hat"]
                extra = ""
            print(f"B batch={batch} rep={rep} q={q}: equal {torch.equal(got[i], want)} "
                  f"max diff {(got[i] - want).abs().max().item():.3e}{extra}", flush=True)
