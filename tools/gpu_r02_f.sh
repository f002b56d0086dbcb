#!/bin/bash
# round 2, call F: LoRA kernel tests + timings (quick loop for kernel tuning)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "lora or dropout" > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -4 gpurun_out/r02f_pytest.log
python tools/lora_bench.py > gpurun_out/r02f_lora_bench.log 2>&1; tail -4 gpurun_out/r02f_lora_bench.log
