#!/bin/bash
# round 2, call M: mask stages (TMA -> mask in shared memory -> tcgen05) for the LoRA down product and dA
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -12 gpurun_out/r02m_pytest.log
timeout 300 python tools/lora_bench.py > gpurun_out/r02m_lora_bench.log 2>&1; tail -3 gpurun_out/r02m_lora_bench.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02m_profile.json > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; cut -c1-260 gpurun_out/r02m_bench.json; tail -3 gpurun_out/r02m_bench.err
NS_NO_MASK_STAGE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_nostage.json 2>> gpurun_out/r02m_bench.err; cut -c1-200 gpurun_out/r02m_bench_nostage.json
