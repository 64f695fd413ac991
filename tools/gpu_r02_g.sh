#!/bin/bash
# round 2, call G: table-maker variant with the warp-cooperative sink: tests, rate; C3 with block transforms + tilt grid
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tabulator.py -m gpu -q > gpurun_out/pytest_gpu_r02_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_g.log
tail -15 gpurun_out/pytest_gpu_r02_g.log
timeout 300 python tools/tabulator_rate.py 32768 20 3 fast > gpurun_out/tabulator_rate_r02_g_fast.json 2>&1; cat gpurun_out/tabulator_rate_r02_g_fast.json
timeout 300 python tools/tabulator_rate.py 262144 20 3 fast > gpurun_out/tabulator_rate_r02_g_fast_big.json 2>&1; cat gpurun_out/tabulator_rate_r02_g_fast_big.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_g.json 2> gpurun_out/bench_r02_g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_g.json').read().strip().split('\n')[-1])
print("value %.4g e2e %.4g frac %.3f"%(d['value'], d['e2e']['value'], d['roofline']['frac']))
for k,v in (d.get('other_configs') or {}).items(): print(k, "%.4g"%v['value'], "hitfrac %.5f seg/ph %.2f frac %.3f ms %.2f"%(v['hit_fraction'], v['segments_per_photon'], v['roofline_frac'], v['ms_per_step']))
PY
timeout 900 python -m pytest tests/test_gpu_fast_kernel.py -m gpu -q -x -k "spice_lea or tilt or statist" 2>&1 | tail -4
