#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, ncu launch list.  Outputs in gpurun_out/.
mkdir -p gpurun_out
tag=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$tag.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.log; tail -2 gpurun_out/smoke_$tag.log; cat gpurun_out/bench_$tag.json
