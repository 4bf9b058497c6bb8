#!/usr/bin/env python
"""Accuracy and time of the three convolution kernels (1 = fp32 SIMT, 2 = 3xTF32 tcgen05, 3 = fp16-split tcgen05) on
the layer shapes that matter, against a float64 CPU reference:  python tools/acc_conv.py [--batch 8]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from progressivecodec_b200.engine import Act, Engine, pack_conv2d

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--impls", default="1,2,3")
args = ap.parse_args()
dev = torch.device("cuda", 0)
shapes = [(512, 224, 3, 1, (32, 48)), (224, 176, 3, 1, (32, 48)), (176, 128, 3, 1, (32, 48)), (128, 64, 3, 1, (32, 48)),
          (64, 32, 3, 1, (32, 48)), (192, 192, 5, 2, (256, 384)), (192, 96, 1, 1, (128, 192)), (96, 192, 1, 1, (128, 192)),
          (192, 192, 1, 1, (128, 192)), (640, 320, 3, 1, (32, 48))]
for cin, cout, k, stride, hw in shapes:
    torch.manual_seed(1)
    m = nn.Conv2d(cin, cout, k, stride, k // 2)
    B = args.batch if hw[0] <= 64 else max(1, args.batch // 4)
    x = torch.randn(B, cin, *hw)
    ref = torch.nn.functional.conv2d(x[:1].double(), m.weight.double(), m.bias.double(), stride, k // 2)
    rms = ref.pow(2).mean().sqrt().item()
    flops = 2.0 * B * (hw[0] // stride) * (hw[1] // stride) * cout * cin * k * k
    line = f"{k}x{k}s{stride} {cin:4d}->{cout:4d} @{hw[0]}x{hw[1]} B={B}:"
    for impl in [int(v) for v in args.impls.split(",")]:
        E = Engine(dev, impl)
        pc = pack_conv2d(m, dev, "t").attach_tc(3)
        xa = Act(x.permute(0, 2, 3, 1).contiguous().cuda())
        try:
            out = E.conv_new(pc, [xa])
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            line += f"  impl{impl}: {type(e).__name__}"
            continue
        got = out.t[:1].permute(0, 3, 1, 2).double().cpu()
        err = (got - ref)
        for _ in range(2):
            E.conv(pc, [xa], out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            E.conv(pc, [xa], out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        line += f"  impl{impl}: rms {err.pow(2).mean().sqrt().item() / rms:.2e} max {err.abs().max().item() / rms:.2e} {flops / ms / 1e9:6.1f} TF/s"
    print(line, flush=True)
