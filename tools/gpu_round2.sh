#!/bin/bash
# Round-2 measurement pass (run under gpurun, ONE GPU): the GPU test suite, the headline bench line and the other configurations
# of BASELINE.json through bench.py --config.  The JSON lines land in gpurun_out/r02_bench_*.json (copied to profiles/ afterwards).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log; tail -3 gpurun_out/r02_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r02_profile.json > gpurun_out/r02_bench_train.json 2> gpurun_out/r02_bench_train.err; cut -c1-200 gpurun_out/r02_bench_train.json
timeout 300 python bench.py --steps 10 --warmup 3 --lora-dropout 0 --no-cpu-baseline > gpurun_out/r02_bench_train_p0.json 2>> gpurun_out/r02_bench_train.err; cut -c1-200 gpurun_out/r02_bench_train_p0.json
for c in c273 large decode pipeline; do
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; echo "$c rc=$?"; cut -c1-200 gpurun_out/r02_bench_$c.json; tail -1 gpurun_out/r02_bench_$c.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 600 python tools/bench_decode.py --B 32 --max-length 448 --batches 2 --beams 5 > gpurun_out/r02_decode_beam.json 2> gpurun_out/r02_decode_beam.err; python -c "
import json; d=json.load(open('gpurun_out/r02_decode_beam.json')); print(d.get('beam5'))"
