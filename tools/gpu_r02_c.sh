#!/bin/bash
# round 2, call E: LoRA kernel tests + timings after tuning, bench with lora_dropout=0.05 (+ per-launch profile)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "lora or dropout" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -5 gpurun_out/r02e_pytest.log
python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "dropout" >> gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -3 gpurun_out/r02e_pytest.log
python tools/lora_bench.py > gpurun_out/r02e_lora_bench.log 2>&1; tail -4 gpurun_out/r02e_lora_bench.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02e_profile.json > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; cut -c1-300 gpurun_out/r02e_bench.json; tail -3 gpurun_out/r02e_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lora-dropout 0 > gpurun_out/r02e_bench_p0.json 2>> gpurun_out/r02e_bench.err; cut -c1-200 gpurun_out/r02e_bench_p0.json
