#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lora_ -c 5 -f -o gpurun_out/r02g_lora python tools/lora_bench.py --once > gpurun_out/r02g_ncu.log 2>&1
tail -3 gpurun_out/r02g_ncu.log
