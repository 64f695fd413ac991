#!/bin/bash
# round 2, call F: ncu --set full of the C2 kernel (v34) and of the table-maker variant on the persistent kernel; launch list of a bench run
mkdir -p gpurun_out
bash tools/gpu_prof.sh r02_v34
timeout 600 ncu --set full --import-source on --clock-control none -k regex:propagate_persistent -s 1 -c 1 -f -o gpurun_out/r02_v34_tab \
  python tools/tabulator_rate.py 32768 20 1 fast > gpurun_out/ncu_r02_v34_tab.log 2>&1
tail -3 gpurun_out/ncu_r02_v34_tab.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_v34.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r02_v34.log 2>&1
tail -2 gpurun_out/ncu_bench_r02_v34.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_tabulator.py -m gpu -q -k "non_stop or persistent_kernel" 2>&1 | tail -3
