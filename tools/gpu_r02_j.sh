#!/bin/bash
# round 2, call J: row-major dropout plane + masked second product in the input-gradient GEMMs
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -12 gpurun_out/r02j_pytest.log
python tools/lora_bench.py > gpurun_out/r02j_lora_bench.log 2>&1; tail -3 gpurun_out/r02j_lora_bench.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02j_profile.json > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; cut -c1-260 gpurun_out/r02j_bench.json; tail -3 gpurun_out/r02j_bench.err
NS_NO_GEMM_MASK=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_nomask.json 2>> gpurun_out/r02j_bench.err; cut -c1-200 gpurun_out/r02j_bench_nomask.json
timeout 600 python bench.py --config pipeline --steps 16 --no-cpu-baseline > gpurun_out/r02j_bench_pipeline.json 2> gpurun_out/r02j_bench_pipeline.err; cut -c1-300 gpurun_out/r02j_bench_pipeline.json; tail -2 gpurun_out/r02j_bench_pipeline.err
