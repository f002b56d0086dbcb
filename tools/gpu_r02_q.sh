#!/bin/bash
# round 2, call Q (2 GPUs): data-parallel step with the overlapped two-bucket all-reduce -- correctness and timing
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r02q_dp_check.log 2>&1; echo "dp_check rc=$?"; grep "rank" gpurun_out/r02q_dp_check.log | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02q_bench_2gpu.json 2> gpurun_out/r02q_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/r02q_bench_2gpu.json; tail -3 gpurun_out/r02q_bench_2gpu.err
NS_NO_AR_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02q_bench_2gpu_single.json 2>> gpurun_out/r02q_bench_2gpu.err; cut -c1-300 gpurun_out/r02q_bench_2gpu_single.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02q_bench_ref_2gpu.json 2>> gpurun_out/r02q_bench_2gpu.err; cut -c1-200 gpurun_out/r02q_bench_ref_2gpu.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
