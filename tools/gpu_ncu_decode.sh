#!/bin/bash
# part (3) of tools/gpu_ncu_r02.sh alone: ncu --set full of the kernels of one greedy-decode position (B = 128, eeg_ch = 273)
mkdir -p gpurun_out
TAG=${1:-r02f}
NS_DECODE_NCU=1 ncu --set full --clock-control none -k regex:'attn_decode|cross_absorbed|gemm_nt_kernel|greedy|ln_fwd|embed' -s 1000 -c 80 -f -o /tmp/${TAG}_full_decode \
    python tools/bench_decode.py --B 128 --max-length 48 --batches 1 > gpurun_out/${TAG}_ncu_decode.log 2>&1
ncu -i /tmp/${TAG}_full_decode.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_decode.csv 2>/dev/null
ls -la gpurun_out/${TAG}_full_decode.csv
