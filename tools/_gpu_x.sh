#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log; tail -3 gpurun_out/r02x_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench_ov$i.json 2> gpurun_out/r02x_bench.err; cut -c100-240 gpurun_out/r02x_bench_ov$i.json; tail -2 gpurun_out/r02x_bench.err
NS_NO_PLANE_OVERLAP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench_noov$i.json 2>> gpurun_out/r02x_bench.err; cut -c100-240 gpurun_out/r02x_bench_noov$i.json
done
