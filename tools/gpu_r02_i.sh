#!/bin/bash
# round 2, call I: full GPU suite (ragged-store aug pass, resident loader) + the bench configurations the driver does not run
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -6 gpurun_out/r02i_pytest.log
for c in pipeline decode c273 large; do
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_$c.json 2> gpurun_out/r02i_bench_$c.err
  echo "$c rc=$?"; cut -c1-400 gpurun_out/r02i_bench_$c.json; tail -2 gpurun_out/r02i_bench_$c.err
done
