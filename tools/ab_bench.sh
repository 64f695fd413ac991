#!/bin/bash
# A/B bench of the library variants built by tools/build_variants.py (one gpurun call).
mkdir -p gpurun_out
out=gpurun_out/ab_bench.txt
: > $out
for lib in clsim_b200/variants/*.so; do
  echo "== $lib" >> $out
  CLSIMCU_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants $AB_ARGS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g ms/step %.2f seg/photon %.2f hitfrac %.5f frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['segments_per_photon'], d['config']['hit_fraction'], d['roofline']['frac'], d['clocks']))
    else: print(l)
" >> $out 2>&1
done
cat $out
