#!/bin/bash
# quick GPU check: parity tests of the product path + A/B bench of the variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_engine.py tests/test_host_class.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
tail -15 gpurun_out/pytest_quick.log
bash tools/ab_bench.sh
