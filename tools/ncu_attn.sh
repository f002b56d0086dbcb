#!/bin/bash
# ncu full capture of the attention kernels at a reduced batch (kernel replay is ~40x); run under gpurun.
set -x
K=${1:-attn_bwd_fused}
B=${2:-16}
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_$K -f \
    python tools/kbench.py attn --B $B --iters 1 > gpurun_out/ncu_$K.log 2>&1
tail -5 gpurun_out/ncu_$K.log
