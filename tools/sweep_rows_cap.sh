for v in "X=1" "CLG_ROWS_CAP=40" "CLG_ROWS_CAP=42" "CLG_ROWS_CAP=46" "CLG_ROWS_CAP=48" "CLG_ROWS_CAP=52" "CLG_ROWS_WIN=2048" "CLG_ROWS_WIN=512"; do
  env $v python bench.py --steps 3 --warmup 3 --log2-vectors 22 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],2), d['config']['slots'], d['roofline']['kernel'])"
done
