#!/bin/bash
# round 2, call L: table-maker with the branch-free bin index: tests + rate; ncu --set full of the v37 C2 kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tabulator.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/tabulator_rate.py 262144 20 3 fast > gpurun_out/tabulator_rate_r02_l_fast_big.json 2>&1; cat gpurun_out/tabulator_rate_r02_l_fast_big.json
bash tools/gpu_prof.sh r02_v37
