#!/bin/bash
# round 2, call B: full GPU suite with the LoRA-dropout kernels, their timings, bench with dropout
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -25 gpurun_out/r02b_pytest.log
python tools/lora_bench.py > gpurun_out/r02b_lora_bench.log 2>&1; tail -8 gpurun_out/r02b_lora_bench.log
