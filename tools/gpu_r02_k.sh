#!/bin/bash
# round 2, call K: high-statistics p-value scans of the v37 kernel against the reference-order kernel (2.1e8 photons per arm and repetition)
mkdir -p gpurun_out
for spec in "homogeneous 8" "spice_mie 12" "spice_lea 12" "spice_mie_tilt 8" "spice_lea_notilt 8"; do
  set -- $spec
  timeout 900 python tests/tools/pvalue_scan.py $1 524288 $2 > gpurun_out/pvalue_scan_r02_v37_$1.txt 2>&1
  tail -2 gpurun_out/pvalue_scan_r02_v37_$1.txt
done
