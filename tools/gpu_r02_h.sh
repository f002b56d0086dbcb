#!/bin/bash
# round 2, call H: full GPU suite + LoRA kernel timings + bench (dropout on / off) with per-launch profile
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -6 gpurun_out/r02h_pytest.log
python tools/lora_bench.py > gpurun_out/r02h_lora_bench.log 2>&1; tail -3 gpurun_out/r02h_lora_bench.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02h_profile.json > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; cut -c1-260 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lora-dropout 0 > gpurun_out/r02h_bench_p0.json 2>> gpurun_out/r02h_bench.err; cut -c1-200 gpurun_out/r02h_bench_p0.json
