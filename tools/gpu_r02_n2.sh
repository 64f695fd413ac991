#!/bin/bash
# round 2: the bench at N=2 under torchrun (weak-scaling headline + the three multi-GPU legs), then the reference arm the same way
mkdir -p gpurun_out
N=${1:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err ) 2> gpurun_out/bench_r02_n$N.time
tail -3 gpurun_out/bench_r02_n$N.err; cat gpurun_out/bench_r02_n$N.time
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_n$N.json').read().strip().split('\n')[-1])
print("n_gpus", d['n_gpus'], "value %.4g e2e %.4g frac %.3f"%(d['value'], d['e2e']['value'], d['roofline']['frac']))
for k,v in (d.get('multi_gpu') or {}).items(): print(k, {a:b for a,b in v.items() if a!='workload'})
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ref_r02_n$N.json 2> gpurun_out/bench_ref_r02_n$N.err ) 2>> gpurun_out/bench_r02_n$N.time
cut -c1-300 gpurun_out/bench_ref_r02_n$N.json; tail -4 gpurun_out/bench_r02_n$N.time
