#!/bin/bash
# quick GPU check: parity tests of the product path + A/B bench of the variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_engine.py tests/test_host_class.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
tail -15 gpurun_out/pytest_quick.log
bash tools/ab_bench.sh
for b in 8192 16384; do echo "== pixel budget $b"; CLSIMCU_PIXEL_BUDGET=$b python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
"; done
CLSIMCU_DEBUG_STATS=1 CLSIMCU_LIB=$PWD/clsim_b200/variants/libclsimcuda_T1024_B1_CLSIMCU_DEBUG_COUNTERS.so python bench.py --steps 1 --warmup 3 --no-cpu-baseline --bunch 262144 2>&1 | grep -v "^{"
