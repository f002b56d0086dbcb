#!/bin/bash
# round 2, call N: programmatic dependent launch in the decode loop, vectorised greedy pick; regression of the training bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -8 gpurun_out/r02n_pytest.log
timeout 600 python bench.py --config decode --steps 5 --warmup 3 > gpurun_out/r02n_bench_decode.json 2> gpurun_out/r02n_bench_decode.err; cut -c1-330 gpurun_out/r02n_bench_decode.json; tail -2 gpurun_out/r02n_bench_decode.err
NS_NO_PDL=1 timeout 600 python bench.py --config decode --steps 5 --warmup 3 > gpurun_out/r02n_bench_decode_nopdl.json 2>> gpurun_out/r02n_bench_decode.err; cut -c1-330 gpurun_out/r02n_bench_decode_nopdl.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02n_profile.json > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; cut -c1-260 gpurun_out/r02n_bench.json; tail -3 gpurun_out/r02n_bench.err
