#!/bin/bash
# ncu evidence for the round-1 build (run under gpurun; one GPU).  Reports are converted to CSV on the box and only the
# small ones travel back (gpurun_out/ is capped at 64 MiB).
#  (1) per-launch device time of every kernel of one training step (cold-cache, serialised: compare SHARES)
#  (2) --set full captures: the first 12 tcgen05 GEMM launches of a step (stem convs + encoder layer 0 forward), and one
#      launch each of the encoder attention forward / fused backward at the bench shape
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 460 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:gemm_nt_kernel -s 574 -c 12 -o /tmp/prof_gemm_nt -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm_nt.ncu-rep --page raw --csv > gpurun_out/prof_gemm_nt.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused_kernel -s 4 -c 1 -o gpurun_out/prof_attn_bwd_fused -f \
    python tools/kbench.py attn --B 64 --iters 1 > gpurun_out/ncu_attn_bwd.log 2>&1
ncu --set full --clock-control none -k regex:attn_fwd_tc_kernel -s 4 -c 1 -o /tmp/prof_attn_fwd -f \
    python tools/kbench.py attn --B 64 --iters 1 > gpurun_out/ncu_attn_fwd.log 2>&1
ncu -i /tmp/prof_attn_fwd.ncu-rep --page raw --csv > gpurun_out/prof_attn_fwd.csv 2>/dev/null
ncu -i gpurun_out/prof_attn_bwd_fused.ncu-rep --page raw --csv > gpurun_out/prof_attn_bwd_fused.csv 2>/dev/null
ls -la gpurun_out/ | tail -12; du -sh gpurun_out
