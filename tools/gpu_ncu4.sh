#!/bin/bash
# ncu launch list of one training step of the final round-1 build (run under gpurun, one GPU).  The eager step is profiled
# (NS_NO_TRAIN_GRAPH=1): ncu serialises kernels anyway and a graph replay would hide the per-kernel names of the warm-up.
mkdir -p gpurun_out
NS_NO_TRAIN_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 460 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/launches.csv
