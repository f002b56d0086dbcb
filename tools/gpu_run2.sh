#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -c 3000 gpurun_out/bench.log
