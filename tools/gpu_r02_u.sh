#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log
tail -4 gpurun_out/r02u_pytest.log
timeout 600 python tools/bench_decode.py --B 128 --max-length 448 --batches 3 --beams 5 > gpurun_out/r02u_decode.json 2> gpurun_out/r02u_decode.err; python -c "
import json; d=json.load(open('gpurun_out/r02u_decode.json')); print('eager(native)', d['eager']); print('graphs', d['cuda_graphs'], d['graphs_match_eager']); print(d.get('beam5'))"; tail -2 gpurun_out/r02u_decode.err
