#!/bin/bash
# round 2: compute-sanitizer (memcheck, racecheck, synccheck) over every kernel of the library on small inputs
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_${tool}_r02.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|sanitize smoke|Error" gpurun_out/sanitizer_${tool}_r02.log | tail -5
done
