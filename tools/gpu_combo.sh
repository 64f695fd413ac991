#!/bin/bash
# One gpurun call: quick parity tests, A/B of the variants, the SpiceLea + tilt + anisotropy side measurement,
# and an ncu --set full capture of the default library.  usage: bash tools/gpu_combo.sh <tag>
tag=${1:-combo}
bash tools/gpu_quick.sh
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants --ice spice_lea --tilt > gpurun_out/bench_${tag}_spice_lea_tilt.json 2> gpurun_out/bench_${tag}_spice_lea_tilt.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${tag}_spice_lea_tilt.json').read().strip().splitlines()[-1]); print('spice_lea+tilt value %.4g e2e %.4g seg/photon %.2f' % (d['value'], d['e2e']['value'], d['roofline']['segments_per_photon']))"
bash tools/gpu_prof.sh $tag
