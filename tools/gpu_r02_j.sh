#!/bin/bash
# round 2, call J: A/B of the approximate sphere pre-test (default) against the z-range pre-test; fast-kernel parity tests on the new default
mkdir -p gpurun_out
out=gpurun_out/ab_r02_j_pretest.txt; : > $out
for lib in clsim_b200/variants/*.so; do
  echo "== $lib" >> $out
  CLSIMCU_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g ms/step %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))
        for k,v in (d.get('other_configs') or {}).items(): print(k, '%.4g'%v['value'], 'hitfrac %.5f seg/ph %.2f frac %.3f ms %.2f'%(v['hit_fraction'], v['segments_per_photon'], v['roofline_frac'], v['ms_per_step']))
" >> $out 2>&1
done
cat $out
timeout 1500 python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_engine.py -m gpu -q > gpurun_out/pytest_gpu_r02_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_j.log
tail -6 gpurun_out/pytest_gpu_r02_j.log
