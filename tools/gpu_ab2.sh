#!/bin/bash
# quick parity tests, then the A/B of the variants on the headline workload and on SpiceLea + tilt + anisotropy
bash tools/gpu_quick.sh
cp gpurun_out/ab_bench.txt gpurun_out/ab_bench_plain.txt
AB_ARGS="--ice spice_lea --tilt" bash tools/ab_bench.sh
cp gpurun_out/ab_bench.txt gpurun_out/ab_bench_spice_lea_tilt.txt
