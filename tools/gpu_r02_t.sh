#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest.log
tail -4 gpurun_out/r02t_pytest.log
timeout 600 python tools/bench_decode.py --B 128 --max-length 448 --batches 3 > gpurun_out/r02t_decode.json 2> gpurun_out/r02t_decode.err; python -c "
import json; d=json.load(open('gpurun_out/r02t_decode.json')); print('eager(native)', d['eager']); print('graphs', d['cuda_graphs'], d['graphs_match_eager'])"; tail -2 gpurun_out/r02t_decode.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02t_profile.json > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; cut -c1-260 gpurun_out/r02t_bench.json; tail -3 gpurun_out/r02t_bench.err
timeout 600 python bench.py --config large --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench_large.json 2> gpurun_out/r02t_bench_large.err; cut -c1-260 gpurun_out/r02t_bench_large.json; tail -3 gpurun_out/r02t_bench_large.err
