#!/bin/bash
# round 2, call B: bench with the new legs at N=1 (other_configs + the multi-GPU legs' first point), reference arm
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 5 --warmup 3 --multi-gpu-legs > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err ) 2> gpurun_out/bench_r02_b.time
tail -5 gpurun_out/bench_r02_b.err; cat gpurun_out/bench_r02_b.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_b.json').read().strip().split('\n')[-1])
print("value %.4g e2e %.4g frac %.3f cpu %s"%(d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline')))
for k,v in (d.get('other_configs') or {}).items(): print(k, "%.4g"%v['value'], "hitfrac %.5f seg/ph %.2f frac %.3f ms %.2f"%(v['hit_fraction'], v['segments_per_photon'], v['roofline_frac'], v['ms_per_step']))
for k,v in (d.get('multi_gpu') or {}).items(): print(k, {a:b for a,b in v.items() if a!='workload'})
PY
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r02_b.json 2> gpurun_out/bench_ref_r02_b.err ) 2>> gpurun_out/bench_r02_b.time
cat gpurun_out/bench_ref_r02_b.json | cut -c1-400; tail -3 gpurun_out/bench_r02_b.time
