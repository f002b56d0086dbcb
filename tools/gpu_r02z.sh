#!/bin/bash
# round-2 late pass: kernel + model tests, the training bench with / without the plane drawn in the mask stage
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log; tail -5 gpurun_out/r02z_pytest.log
for v in fuse nofuse; do
  if [ $v = nofuse ]; then export NS_NO_PLANE_FUSE=1; else unset NS_NO_PLANE_FUSE; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r02z_profile_$v.json > gpurun_out/r02z_bench_$v.json 2> gpurun_out/r02z_bench_$v.err
  echo "$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_$v.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['kernel_shares'])"; tail -2 gpurun_out/r02z_bench_$v.err
done
