#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "greedy or decode or beam or attention or golden or programmatic or module" 2>&1 | tail -2
for w in 8 4 0; do echo "NS_DECODE_WARPS=$w"; NS_DECODE_WARPS=$w timeout 600 python bench.py --config decode --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_token_step'], d['roofline']['frac'])"; done
