#!/bin/bash
# round 2, call D: ncu --set full of the LoRA side kernels (one launch each, K=512, G=1 and 3)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lora_ -c 10 -f -o gpurun_out/r02d_lora python tools/lora_bench.py --once > gpurun_out/r02d_ncu.log 2>&1
tail -5 gpurun_out/r02d_ncu.log; ls -la gpurun_out/*.ncu-rep
