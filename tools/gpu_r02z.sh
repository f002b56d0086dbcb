#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "absorbed or greedy or decode or config4" > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log; tail -4 gpurun_out/r02z_pytest.log
timeout 120 python tools/absorbed_bench.py 2>&1 | tail -4
for i in 1 2; do
  timeout 300 python bench.py --config decode --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_decode_$i.json 2> gpurun_out/r02z_decode_$i.err
  echo "run $i rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02z_decode_$i.json')); print(d['value'], d['ms_per_token_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/r02z_decode_$i.err
done
