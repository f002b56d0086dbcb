#!/bin/bash
# ncu evidence, round 1 third pass (r01c; run under gpurun, one GPU).  Reports are converted to CSV on the box.
#  (1) per-launch device time of every kernel of one training step (cold-cache, serialised: compare SHARES)
#  (2) --set full: the first 12 tcgen05 GEMM launches of a step (stem convs + encoder layer 0 forward)
#  (3) --set full: the augmentation pass at the bench shape
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 460 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:gemm_nt_kernel -s 574 -c 12 -o /tmp/prof_gemm_nt -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm_nt.ncu-rep --page raw --csv > gpurun_out/prof_gemm_nt.csv 2>/dev/null
ncu --set full --clock-control none -k regex:aug_btc_kernel -s 2 -c 1 -o /tmp/prof_aug -f \
    python tools/kbench.py aug --iters 2 > gpurun_out/ncu_aug.log 2>&1
ncu -i /tmp/prof_aug.ncu-rep --page raw --csv > gpurun_out/prof_aug.csv 2>/dev/null
ls -la gpurun_out/ | tail -8; du -sh gpurun_out
