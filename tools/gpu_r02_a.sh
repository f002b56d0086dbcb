#!/bin/bash
# round 2, call A: new parity tests, per-tensor bf16 gradient errors, baseline bench of the round-1 build on this box
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "config4" > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
python tools/grad_err_report.py TINY MID BASE > gpurun_out/r02a_grad_err.log 2>&1; tail -40 gpurun_out/r02a_grad_err.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; cat gpurun_out/r02a_bench.json | cut -c1-600
