#!/bin/bash
# round-2 late pass: full GPU suite, then the training bench with / without the one-pass B-side LoRA backward
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log; tail -5 gpurun_out/r02z_pytest.log
for v in fused nofused; do
  if [ $v = nofused ]; then export NS_NO_FUSED_BWD_B=1; else unset NS_NO_FUSED_BWD_B; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02z_profile_$v.json > gpurun_out/r02z_bench_$v.json 2> gpurun_out/r02z_bench_$v.err
  echo "$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_$v.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['kernel_shares'])"; tail -2 gpurun_out/r02z_bench_$v.err
done
