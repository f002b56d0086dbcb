#!/bin/bash
# Per-instruction stall samples of the tcgen05 GEMM on the fc1 shape (developer aid; run under gpurun, one GPU).
mkdir -p gpurun_out
for k in plain gelu; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 1 -c 1 -o /tmp/prof_epi_$k -f \
      python tools/gemm_epi_trace.py $k 1 > gpurun_out/ncu_epi_$k.log 2>&1
  ncu -i /tmp/prof_epi_$k.ncu-rep --page source --csv > gpurun_out/prof_epi_${k}_source.csv 2>/dev/null
  ncu -i /tmp/prof_epi_$k.ncu-rep --page raw --csv > gpurun_out/prof_epi_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out/prof_epi_*
