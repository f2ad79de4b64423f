for c in 1000 1250 1600 2000; do
  echo "== e2e-chunk $c"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk $c 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('dev ms',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'raw_only ms',round(d['raw_only']['ms_per_step'],2))
"
done
