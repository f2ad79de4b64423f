#!/bin/bash
# A/B of library builds on one box: for every variants/*.so, install it as the package's libdysb200.so, run the
# device-resident bench and print the step time and the per-kernel times.  usage: tools/ab_variants.sh [steps]
# Build a variant with:  touch <pkg>/csrc/*.cu && make VARIANT_FLAGS=-D... LIB=variants/<name>.so
STEPS=${1:-10}
PKG=recognizing-speech-dysfluencies-in-stuttering_b200
mkdir -p gpurun_out
for so in variants/*.so; do
  name=$(basename $so .so)
  cp $so $PKG/libdysb200.so
  python bench.py --steps $STEPS --warmup 3 --device-only --no-cpu-baseline 2> gpurun_out/ab_$name.err | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['kernel_ms_per_step']
        print('$name', 'ms_per_step', round(d['ms_per_step'],3), {a: round(b,3) for a,b in k.items() if b > 0.1})
"
done
