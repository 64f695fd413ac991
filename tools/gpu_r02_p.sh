#!/bin/bash
# round 2, call P: table-maker with RED adds: tests, rate, ncu with per-instruction counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tabulator.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/tabulator_rate.py 262144 20 3 fast > gpurun_out/tabulator_rate_r02_p_fast_big.json 2>&1; cat gpurun_out/tabulator_rate_r02_p_fast_big.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:propagate_persistent -s 1 -c 1 -f -o gpurun_out/r02_v38_tab \
  python tools/tabulator_rate.py 131072 20 1 fast > gpurun_out/ncu_r02_v38_tab.log 2>&1
tail -2 gpurun_out/ncu_r02_v38_tab.log
