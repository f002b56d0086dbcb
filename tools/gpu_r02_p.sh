#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "greedy or beam or decode or golden or generate or module" > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -4 gpurun_out/r02p_pytest.log
NS_GEMM_MIN_BN=32 timeout 600 python bench.py --config decode --steps 3 --warmup 3 > gpurun_out/r02p_bench_decode.json 2> gpurun_out/r02p_bench_decode.err; cut -c100-330 gpurun_out/r02p_bench_decode.json; tail -2 gpurun_out/r02p_bench_decode.err
timeout 600 python tools/bench_decode.py --B 128 --max-length 448 --batches 2 --beams 5 > gpurun_out/r02p_decode_beam.json 2> gpurun_out/r02p_decode_beam.err; python -c "
import json; d=json.load(open('gpurun_out/r02p_decode_beam.json')); print(d['cuda_graphs']); print(d.get('beam5'))"
