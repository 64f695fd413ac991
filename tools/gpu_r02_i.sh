#!/bin/bash
# round 2, call I: full GPU suite on the tree with step parts + fast coordinate math; table-maker rate; bench (other_configs: config1 with step parts)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_i.log
tail -8 gpurun_out/pytest_gpu_r02_i.log
timeout 300 python tools/tabulator_rate.py 262144 20 3 fast > gpurun_out/tabulator_rate_r02_i_fast_big.json 2>&1; cat gpurun_out/tabulator_rate_r02_i_fast_big.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_i.json 2> gpurun_out/bench_r02_i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_i.json').read().strip().split('\n')[-1])
print("value %.4g e2e %.4g frac %.3f"%(d['value'], d['e2e']['value'], d['roofline']['frac']))
for k,v in (d.get('other_configs') or {}).items(): print(k, "%.4g"%v['value'], "hitfrac %.5f seg/ph %.2f frac %.3f ms %.2f"%(v['hit_fraction'], v['segments_per_photon'], v['roofline_frac'], v['ms_per_step']))
PY
