#!/bin/bash
# ncu evidence: (1) per-launch durations of one profiled step, (2) full-set capture of the top kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 40 -c 4 -o gpurun_out/prof_gemm_nt -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 12 -c 3 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/ | tail; tail -2 gpurun_out/ncu_bench.log
