#!/bin/bash
# round 2, call D: GPU suite on the fixed-offset layout; non-stop test with the compiled-in mix; A/B bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_d.log
tail -15 gpurun_out/pytest_gpu_r02_d.log
CLSIMCU_LIB=$PWD/clsim_b200/variants/libclsimcuda_T1024_B1_CLSIMCU_NONSTOP_MIXED.so timeout 600 python -m pytest tests/test_gpu_fast_kernel.py -m gpu -q -k non_stop > gpurun_out/pytest_gpu_r02_d_nonstop_mixed.log 2>&1
tail -5 gpurun_out/pytest_gpu_r02_d_nonstop_mixed.log
bash tools/ab_bench.sh > /dev/null 2>&1
cp gpurun_out/ab_bench.txt gpurun_out/ab_r02_d_plain.txt
cat gpurun_out/ab_r02_d_plain.txt
