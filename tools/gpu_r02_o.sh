#!/bin/bash
# round 2, call O: the generic-transform hook (test + timing), then the full GPU suite on the final tree
mkdir -p gpurun_out
python tools/prof_config.py config3; CLSIMCU_GENERIC_TRANSFORMS=1 python tools/prof_config.py config3
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02_v38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_v38.log
tail -5 gpurun_out/pytest_gpu_r02_v38.log
