#!/bin/bash
# 8-GPU validation: the driver's command (shorter), then BASELINE configs[4] (data set sharded over the ranks).
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"
cut -c 1-260 gpurun_out/r02_bench_n8.json
timeout 400 $TR bench.py --gpus 8 --dataset 1024 > gpurun_out/r02_dataset1024_n8.json 2> gpurun_out/r02_dataset1024_n8.err; echo "dataset 1024 rc=$?"
cat gpurun_out/r02_dataset1024_n8.json
timeout 600 $TR bench.py --gpus 8 --dataset 4096 > gpurun_out/r02_dataset4096_n8.json 2> gpurun_out/r02_dataset4096_n8.err; echo "dataset 4096 rc=$?"
cat gpurun_out/r02_dataset4096_n8.json
grep -h "FAILED\|Error" gpurun_out/r02_*n8.err | head -5
