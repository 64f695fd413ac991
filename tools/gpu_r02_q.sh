#!/bin/bash
# round 2, call Q: table-maker with the branch-free bin index: tests and rate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tabulator.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/tabulator_rate.py 262144 20 3 fast > gpurun_out/tabulator_rate_r02_q_fast_big.json 2>&1; cat gpurun_out/tabulator_rate_r02_q_fast_big.json
timeout 300 python tools/tabulator_rate.py 262144 20 3 reference > gpurun_out/tabulator_rate_r02_q_reference_big.json 2>&1; cat gpurun_out/tabulator_rate_r02_q_reference_big.json
