#!/bin/bash
# ncu --set full capture of the fast kernel on a full bench bunch (one launch), report to gpurun_out/
tag=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:propagate_persistent -s 2 -c 1 -f -o gpurun_out/$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
