#!/bin/bash
# round 2, call M: ncu of config 5 after the sphere pre-test
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:propagate_persistent -s 2 -c 1 -f -o gpurun_out/r02_v37_config5 python tools/prof_config.py config5 > gpurun_out/ncu_r02_v37_config5.log 2>&1
tail -2 gpurun_out/ncu_r02_v37_config5.log
