#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y_pytest.log; tail -4 gpurun_out/r02y_pytest.log
for sk in 0 1; do echo "NS_NO_SKINNY=$sk"; if [ $sk = 1 ]; then export NS_NO_SKINNY=1; fi; timeout 600 python bench.py --config decode --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_token_step'], d['roofline']['frac'], d['e2e']['value'])"; done
