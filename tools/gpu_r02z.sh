#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log; tail -6 gpurun_out/r02z_pytest.log
for v in fold nofold; do
  if [ $v = nofold ]; then export NS_NO_LN_FOLD=1; else unset NS_NO_LN_FOLD; fi
  timeout 300 python bench.py --config decode --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_decode_$v.json 2> gpurun_out/r02z_decode_$v.err
  echo "$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02z_decode_$v.json')); print(d['value'], d['ms_per_token_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"; tail -2 gpurun_out/r02z_decode_$v.err
done
