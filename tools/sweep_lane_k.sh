#!/bin/bash
# words per thread of the lane interpreter (CLG_LANE_K) on programs of different stack sizes
for wl in "mul16 24" "adder32 24" "dag_thin 21" "mul64 20"; do
  set -- $wl
  for k in 1 2 4; do
    CLG_LANE_K=$k python bench.py --workload $1 --mode lane --log2-vectors $2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 2^$2 K=$k', round(d['ms_per_step'],4), d['roofline'].get('kernel'), 'slots', d['config']['slots'])" 2>&1 | tail -1
  done
done
