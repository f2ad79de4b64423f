for cfg in "1024 1024" "4096 4096" "20000 10000" "256 256" "128 128"; do
  set -- $cfg
  echo "== FEAT=$1 NR=$2"
  DYS_FEAT_SUBBATCH=$1 DYS_NR_SUBBATCH=$2 python bench.py --steps 5 --warmup 3 --device-only 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('ms_per_step',round(d['ms_per_step'],2)); print(d['roofline']['kernel_ms_per_step'])
"
done
