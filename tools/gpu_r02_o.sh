#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -5 gpurun_out/r02o_pytest.log
for bn in 64 32; do
NS_GEMM_MIN_BN=$bn timeout 600 python bench.py --config decode --steps 3 --warmup 3 > gpurun_out/r02o_bench_decode_bn$bn.json 2> gpurun_out/r02o_bench_decode.err; cut -c100-330 gpurun_out/r02o_bench_decode_bn$bn.json; tail -2 gpurun_out/r02o_bench_decode.err
done
NS_DECODE_PROFILE=1 timeout 600 python tools/bench_decode.py --B 128 --max-length 448 --batches 1 > gpurun_out/r02o_decode_profile.json 2> gpurun_out/r02o_decode_profile.err; python -c "
import json; d=json.load(open('gpurun_out/r02o_decode_profile.json')); print(d.get('profile_ms')); print(d['cuda_graphs'])"
