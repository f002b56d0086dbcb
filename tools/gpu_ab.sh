#!/bin/bash
# A/B of one environment switch on the training bench, alternating runs on the same box: tools/gpu_ab.sh NS_NO_PLANE_FUSE [rounds]
mkdir -p gpurun_out
sw=$1; n=${2:-3}
for i in $(seq 1 $n); do
  for v in on off; do
    if [ $v = off ]; then export $sw=1; else unset $sw; fi
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v$i.json 2> gpurun_out/ab.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_$v$i.json')); print('$sw', '$v' == 'off' and 'SET  ' or 'unset', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'])"
  done
done
