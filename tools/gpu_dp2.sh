#!/bin/bash
# data-parallel overhead on ONE box: 1-GPU bench, then 2 ranks (overlapped two-bucket all-reduce / single all-reduce), alternating
mkdir -p gpurun_out
run1() { timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/dp_$1.json 2> gpurun_out/dp_$1.err; }
run2() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/dp_$1.json 2> gpurun_out/dp_$1.err; }
show() { python -c "
import json
for l in open('gpurun_out/dp_$1.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['n_gpus'], round(d['ms_per_step'],3), 'ms', round(d['value'],1), d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1))"; tail -1 gpurun_out/dp_$1.err; }
for i in 1 2; do
  run1 n1_$i; show n1_$i
  run2 n2_ov_$i; show n2_ov_$i
  NS_NO_AR_OVERLAP=1 run2 n2_single_$i; show n2_single_$i
done
