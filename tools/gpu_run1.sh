#!/bin/bash
# first GPU validation: diag ladder (with timeouts), then the kernel and model parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 python tools/diag_gemm.py nt > gpurun_out/diag_nt.log 2>&1; echo "diag nt exit $?" >> gpurun_out/diag_nt.log
timeout 120 python tools/diag_gemm.py tn > gpurun_out/diag_tn.log 2>&1; echo "diag tn exit $?" >> gpurun_out/diag_tn.log
timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 2>&1 | tail -150 > gpurun_out/pytest_kernels.log
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 2>&1 | tail -150 > gpurun_out/pytest_model.log
tail -5 gpurun_out/diag_nt.log gpurun_out/diag_tn.log gpurun_out/pytest_kernels.log gpurun_out/pytest_model.log
