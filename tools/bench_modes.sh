#!/bin/bash
# usage: tools/bench_modes.sh <workload> "<modes>" [extra bench args]  -- one line per execution mode
for m in $2; do
  python bench.py --workload $1 --mode $m --no-cpu-baseline --no-e2e --steps 5 ${@:3} 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(sys.argv[1], d['config']['mode'], 'K=%d' % d['config']['words_per_thread_or_cta'], 'ms=%.4f' % d['ms_per_step'], 'gate-evals/s=%.3e' % d['value'], 'hbm=%.3f' % r['hbm_frac'], 'lop3_issued=%.3f' % r['frac_issued_of_nominal_lop3'], 'lop3_alg=%.3f' % (r['algorithmic_lane_ops_per_launch']/ (d['ms_per_step']*1e-3) / 18.61248e12))" $1 || echo "$1 $m failed"
done
