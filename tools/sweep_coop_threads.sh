#!/bin/bash
# usage: tools/sweep_coop_threads.sh "1024 960 ..." [extra bench args]   -- tuning experiment, not part of the product
for t in $1; do
  CLG_COOP_THREADS=$t python bench.py --no-cpu-baseline --no-e2e --steps 3 --log2-vectors 22 ${@:2} 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d['ms_per_step'],3))" $t
done
