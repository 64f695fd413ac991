#!/bin/bash
# round 2, call C: full GPU suite (fast-kernel non-stop + history are new), C++ host tests, short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_c.log
tail -15 gpurun_out/pytest_gpu_r02_c.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; tail -3 gpurun_out/bench_r02_c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_c.json').read().strip().split('\n')[-1])
print("value %.4g e2e %.4g frac %.3f"%(d['value'], d['e2e']['value'], d['roofline']['frac']))
for k,v in (d.get('other_configs') or {}).items(): print(k, "%.4g"%v['value'], "hitfrac %.5f seg/ph %.2f frac %.3f ms %.2f"%(v['hit_fraction'], v['segments_per_photon'], v['roofline_frac'], v['ms_per_step']))
PY
