#!/bin/bash
# 8 ranks on one box: the headline training bench (overlapped two-bucket all-reduce), then one of the GPUs alone for the efficiency
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
echo "8gpu rc=$?"; tail -2 gpurun_out/r02_bench_8gpu.err; grep '^{' gpurun_out/r02_bench_8gpu.json | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1of8.json 2> gpurun_out/r02_bench_1of8.err
echo "1gpu rc=$?"; cut -c1-300 gpurun_out/r02_bench_1of8.json
