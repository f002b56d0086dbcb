#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02a2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a2_pytest.log; tail -4 gpurun_out/r02a2_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c100-240
NS_NO_GELU_DERIV=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c100-240
done
python tools/kbench.py gemm --iters 30 2>/dev/null | grep "^## gemm" | sed -n 3,7p
