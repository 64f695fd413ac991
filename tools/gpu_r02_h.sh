#!/bin/bash
# round 2, call H: ncu --set full of config 3, config 5 and the cooperative table-maker kernel
mkdir -p gpurun_out
for c in config3 config5; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:propagate_persistent -s 2 -c 1 -f -o gpurun_out/r02_v35_$c python tools/prof_config.py $c > gpurun_out/ncu_r02_v35_$c.log 2>&1
  tail -2 gpurun_out/ncu_r02_v35_$c.log
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:propagate_persistent -s 1 -c 1 -f -o gpurun_out/r02_v35_tab \
  python tools/tabulator_rate.py 262144 20 1 fast > gpurun_out/ncu_r02_v35_tab.log 2>&1
tail -3 gpurun_out/ncu_r02_v35_tab.log
