#!/bin/bash
# round 2, call E: non-stop tests again (twin's DOM de-duplication restored), table-maker on the persistent kernel, its rate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_reference_kernel.py tests/test_gpu_tabulator.py -m gpu -q -k "non_stop or non_stopping or tabul or table" > gpurun_out/pytest_gpu_r02_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_e.log
tail -25 gpurun_out/pytest_gpu_r02_e.log
timeout 300 python tools/tabulator_rate.py 32768 20 3 reference > gpurun_out/tabulator_rate_r02_reference.json 2>&1; cat gpurun_out/tabulator_rate_r02_reference.json
timeout 300 python tools/tabulator_rate.py 32768 20 3 fast > gpurun_out/tabulator_rate_r02_fast.json 2>&1; cat gpurun_out/tabulator_rate_r02_fast.json
timeout 300 python tools/tabulator_rate.py 262144 20 3 fast > gpurun_out/tabulator_rate_r02_fast_big.json 2>&1; cat gpurun_out/tabulator_rate_r02_fast_big.json
