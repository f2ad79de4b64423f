# host-streaming sweep: chunk size x compute streams  (usage: tools/sweep_e2e.sh ["chunk streams" ...])
[ $# -eq 0 ] && set -- "400 3" "300 3" "400 4" "250 4" "200 4"
for cfg in "$@"; do
  set -- $cfg
  echo "== e2e-chunk $1 streams $2"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-chunk $1 --e2e-streams $2 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('dev ms',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'h2d-only ms',round(d['e2e']['h2d_only_ms_per_step'],2))
"
done
