#!/usr/bin/env python
"""Benchmark of the progressive codec hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, C-ABI library)
    python bench.py --impl reference --steps K --warmup W  # reference arm: CPU oracle port on the host cores

Metric: 768x512 images/s, compress+decompress, per quality ("image-qualities per second"): one *step* is the
full 13-level quality sweep of the reference's evaluation protocol (training/step.py:322-337, pr_list of
train.py:293) — compress() then decompress() at every level — over a batch of synthetic 768x512 images.
`value` counts (images x qualities) / second with inputs resident in HBM and streams kept on the device;
`e2e` is the same sweep through the public API with HOST buffers (pinned image in, python `bytes` strings out
of compress(), bytes into decompress(), x_hat copied back).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QUALITIES = [0, 0.05, 0.1, 0.25, 0.5, 0.6, 0.75, 1, 1.25, 2, 3, 5, 10]  # reference train.py:293
AUTHORS = dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True, delta_encode=True,
               support_progressive_slices=5, mask_policy="point-based-std")
H, W = 512, 768


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md 'clocks line')."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (oracle port on the host cores)
# ----------------------------------------------------------------------------------------------------------
def _oracle_model():
    import torch

    from oracle.codec_port import CodecConfig, OracleCodec
    from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights

    net = ChannelProgresssiveWACNN(**AUTHORS).eval()
    apply_synthetic_weights(net, seed=0)
    net.update(force=True)
    return net, OracleCodec(net.state_dict(), CodecConfig(**AUTHORS))


def _cpu_time_sweep(orc, x, qualities):
    t0 = time.perf_counter()
    for q in qualities:
        c = orc.compress(x, quality=q)
        orc.decompress(c["strings"], c["shape"], quality=q)
    return time.perf_counter() - t0


def run_reference(args):
    import torch

    from oracle.gen_golden import synthetic_image

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _net, orc = _oracle_model()
    x = synthetic_image((1, 3, H, W), seed=0)
    # one step = one image at one quality level (compress + decompress), cycling through the sweep
    qs = [QUALITIES[(3 * i) % len(QUALITIES)] for i in range(args.warmup + args.steps)]
    for q in qs[:args.warmup]:
        _cpu_time_sweep(orc, x, [q])
    t = 0.0
    for q in qs[args.warmup:]:
        t += _cpu_time_sweep(orc, x, [q])
    value = args.steps / t
    line = {"impl": "reference", "metric": "768x512 img/s compress+decompress per quality", "value": value,
            "unit": "image-qualities/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "compress+decompress, 13-level quality sweep, 768x512 (configs[1]); "
                                   "reference arm: CPU oracle port (torch fp32 restatement of the reference, "
                                   "pinned bit-exact by tests/golden) — the python reference cannot travel to the GPU box",
                       "model": "ChannelProgresssiveWACNN authors' flags, synthetic calibrated weights"},
            "cpu_baseline": {"value": value, "unit": "image-qualities/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps of 1 image x 1 quality (cycling the 13-level sweep)"},
            "e2e": {"value": value, "unit": "image-qualities/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def conv_roofline(net, peaks, peak_kind):
    """Dominant kernel = the tap-GEMM convolution.  Time it alone (CUDA events on the launch stream, L2 flushed
    between launches) on the layer that carries most FLOPs of the transforms: 5x5 stride-2 conv 192->192 at
    256x384 -> 128x192 (45.3 GFLOP, g_a[2])."""
    import torch

    from progressivecodec_b200.engine import Act, new_act

    P = net.prepare()
    E = P["eng"]
    E.begin(0)  # latch the CURRENT stream for this thread's launches (the CUDA events below are recorded on it)
    pc = P["g_a"][0]["c2"]
    nb = 8  # images per launch: 3072 CTAs = 20.8 waves of 148, as in the batched run (one image is 2.6 waves)
    x = Act(torch.randn(nb, 256, 384, 192, device=E.device))
    out = new_act(nb, 128, 192, 192, E.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=E.device)
    flops = 2.0 * nb * 128 * 192 * 192 * (25 * 192)
    for _ in range(3):
        E.conv(pc, [x], out)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        E.conv(pc, [x], out)
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    avg = sum(times) / len(times)
    achieved = flops / avg / 1e12
    peak = peaks["bf16_tflops"]
    # DRAM traffic of this exact launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set
    # full` capture of tools/prof_conv_one.py, recorded in profiles/roofline_traffic.json; null if not captured.
    # Algorithmic bytes = NHWC input + TF32 hi/lo weights + output.
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    alg_bytes = 4.0 * (nb * 256 * 384 * 192 + 2 * 25 * 192 * 192 + nb * 128 * 192 * 192)
    return {"bound": "tensor", "kernel": f"conv_taps_tc_kernel (tcgen05 3xTF32; 5x5 s2 192->192, {nb} x 256x384 -> 128x192)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_unit": "bytes/launch (ncu dram read+write)", "algorithmic_bytes": alg_bytes,
            "peak_kind": peak_kind + " bf16 burst (cuBLAS); this kernel issues 3 TF32 MMAs per algorithmic MAC, "
                         "TF32 runs at half the bf16 rate, so frac <= 1/6",
            "tensor_pipe_frac_of_tf32_peak": 3.0 * achieved / (peak / 2.0),
            "flops_per_launch": flops, "avg_launch_ms": avg * 1e3}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from oracle.gen_golden import synthetic_image  # input generator only (not the checker)
    from progressivecodec_b200 import ChannelProgresssiveWACNN, _lib, apply_synthetic_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    net = ChannelProgresssiveWACNN(**AUTHORS).eval()
    apply_synthetic_weights(net, seed=0)
    net.update(force=True)
    net = net.to(dev)
    B = args.batch
    x_host = torch.cat([synthetic_image((1, 3, H, W), seed=rank * 1000 + i) for i in range(B)]).pin_memory()
    x_dev = x_host.to(dev)
    xhat_host = torch.empty_like(x_host).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    from progressivecodec_b200 import pipeline

    def step_device():
        if args.no_pipeline:
            for q in QUALITIES:
                c = net.compress(x_dev, quality=q, return_device_streams=True)
                net.decompress(c, c["shape"], quality=q)
        else:  # same calls, compress(q+1) overlapped with decompress(q) on two streams / host threads
            pipeline.sweep(net, x_dev, QUALITIES, keep=False)

    h2d = [0]
    d2h = [0]

    def step_e2e():
        h2d[0] = d2h[0] = 0
        if args.no_pipeline:
            for q in QUALITIES:
                xd = x_host.to(dev, non_blocking=True)
                h2d[0] += x_host.numel() * 4
                c = net.compress(xd, quality=q)
                nbytes = sum(len(s) for sl in c["strings"][0] for s in sl) + sum(len(s) for s in c["strings"][1])
                d2h[0] += nbytes
                r = net.decompress(c["strings"], c["shape"], quality=q)["x_hat"]
                xhat_host.copy_(r, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                h2d[0] += nbytes
                d2h[0] += r.numel() * 4
            return
        # host image in (pinned -> device), python `bytes` strings between the stages, x_hat copied back to the host
        xd = x_host.to(dev, non_blocking=True)
        h2d[0] += x_host.numel() * 4

        def on_result(q, c, r):
            nbytes = sum(len(s) for sl in c["strings"][0] for s in sl) + sum(len(s) for s in c["strings"][1])
            d2h[0] += nbytes + r["x_hat"].numel() * 4
            h2d[0] += nbytes
            xhat_host.copy_(r["x_hat"], non_blocking=True)  # pinned destination
            torch.cuda.current_stream().synchronize()

        pipeline.sweep(net, xd, QUALITIES, host_strings=True, on_result=on_result, keep=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        lib.pcodec_reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            flush.zero_()
            fn()
        e1.record()
        barrier()
        t = e0.elapsed_time(e1) * 1e-3
        launches = lib.pcodec_launch_count()
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            tt = torch.tensor([t], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t, launches, clocks

    t_dev, launches, clocks = timed(step_device, args.steps, args.warmup)
    t_e2e, _l2, _c2 = timed(step_e2e, max(1, args.steps // 2), 1)
    e2e_steps = max(1, args.steps // 2)
    units = B * len(QUALITIES)
    value = world * units * args.steps / t_dev
    e2e_value = world * units * e2e_steps / t_e2e

    if rank == 0:
        peaks, kind = _peaks()
        roof = conv_roofline(net, peaks, kind)
        # algorithmic work of the whole step (914.2 GFLOP per image-quality at q>0, 596.7 at q=0)
        flops_step = B * (596.7e9 + 12 * 914.2e9)
        step_tflops = flops_step * args.steps / t_dev / 1e12
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            _n, orc = _oracle_model()
            xs = x_host[:1].clone()
            sample_q = [0, 1.25, 10]
            _cpu_time_sweep(orc, xs, [5])
            tc = _cpu_time_sweep(orc, xs, sample_q)
            cpu = {"value": len(sample_q) / tc, "unit": "image-qualities/s", "cores": cores, "kind": "port",
                   "sample": "1 image 768x512 x qualities [0,1.25,10] (compress+decompress), after 1 warm-up call"}
        lat = None
        if world == 1:
            x1 = x_dev[:1].contiguous()
            for _ in range(2):
                c1 = net.compress(x1, quality=5, return_device_streams=True)
                net.decompress(c1, c1["shape"], quality=5)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c1 = net.compress(x1, quality=5, return_device_streams=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            net.decompress(c1, c1["shape"], quality=5)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            lat = {"batch": 1, "quality": 5, "compress_ms": 1e3 * (t1 - t0), "decompress_ms": 1e3 * (t2 - t1)}
            # configs[1] taken literally: ONE 768x512 image, the full 13-level sweep (pipelined; the decode chains of
            # up to four levels run concurrently because one image leaves the GPU almost idle)
            pipeline.sweep(net, x1, QUALITIES, keep=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipeline.sweep(net, x1, QUALITIES, keep=False)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            lat["sweep_13_levels_ms"] = 1e3 * (t1 - t0)
            lat["sweep_image_qualities_per_s"] = len(QUALITIES) / (t1 - t0)
        line = {"metric": "768x512 img/s compress+decompress per quality", "value": value, "unit": "image-qualities/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"compress+decompress, 13-level quality sweep (configs[1]), batch {B} x 768x512 per GPU",
                           "model": "ChannelProgresssiveWACNN authors' flags (mdmh-mem5-de), synthetic calibrated weights",
                           "batch_per_gpu": B, "qualities": QUALITIES,
                           "l2": "256 MiB buffer written between steps; working set (0.6 GB weights + activations) >> 126 MB L2",
                           "pipeline": ("none" if args.no_pipeline else
                                        "compress(q+1) overlaps decompress(q): 2 host threads / CUDA streams (pipeline.sweep)"),
                           "parallelism": f"dp{world} (images sharded, no collective in the timed region)"},
                "e2e": {"value": e2e_value, "unit": "image-qualities/s", "h2d_bytes_per_step": h2d[0],
                        "d2h_bytes_per_step": d2h[0], "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
                "step_algorithmic_tflops": step_tflops, "single_image_latency": lat, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="768x512 images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="compress/decompress strictly back to back")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
