#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest.log
tail -4 gpurun_out/r02s_pytest.log
timeout 600 python tools/bench_decode.py --B 128 --max-length 448 --batches 3 > gpurun_out/r02s_decode.json 2> gpurun_out/r02s_decode.err; python -c "
import json; d=json.load(open('gpurun_out/r02s_decode.json')); print('eager(native)', d['eager']); print('graphs', d['cuda_graphs'], d['graphs_match_eager'])"; tail -2 gpurun_out/r02s_decode.err
