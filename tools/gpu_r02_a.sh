#!/bin/bash
# round 2, call A: full GPU test-suite, then A/B of kernel variants on C2 and on SpiceLea + tilt
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_r02_a.txt 2>&1
nproc >> gpurun_out/smi_r02_a.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_a.log
tail -5 gpurun_out/pytest_gpu_r02_a.log
bash tools/ab_bench.sh > /dev/null 2>&1
cp gpurun_out/ab_bench.txt gpurun_out/ab_r02_a_plain.txt
AB_ARGS="--ice spice_lea --tilt" bash tools/ab_bench.sh > /dev/null 2>&1
cp gpurun_out/ab_bench.txt gpurun_out/ab_r02_a_spice_lea_tilt.txt
cat gpurun_out/ab_r02_a_plain.txt gpurun_out/ab_r02_a_spice_lea_tilt.txt
