for cfg in "400 3" "625 3" "400 4" "625 2"; do
  set -- $cfg
  echo "== e2e-chunk $1 streams $2"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk $1 --e2e-streams $2 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('dev ms',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'raw_only ms',round(d['raw_only']['ms_per_step'],2))
"
done
