"""Throughput of the table-maker variant on the default spherical layout (202 x 38 x 102 x 107 bins, 335 MB in HBM).
usage (GPU box): python tools/tabulator_rate.py [steps] [photons_per_step] [bunches] [fast|reference]"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clsim_b200 import ice, mcpe, steps, tabulator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
pps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bunches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = {"fast": 0, "reference": 1}[sys.argv[4]] if len(sys.argv) > 4 else 0
medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_mie", useTiltIfAvailable=False)
axes = tabulator.default_axes()
conv = tabulator.I3CLSimStepToTableConverter(0, axes, 0, False, medium, None, math.pi * 0.1651 ** 2, ice.GetIceCubeDOMAcceptance(),
                                             mcpe.GetIceCubeDOMAngularSensitivity(), 1, maxNumWorkitems=n, kernelMode=mode)
ref = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)
bunch = steps.cascade_steps(n, photons_per_step=pps, energy_gev=1e3, zenith_deg=180.0, azimuth_deg=0.0, seed=1)
conv.EnqueueSteps(bunch, ref)
conv.Finish()
t0 = time.perf_counter()
for _ in range(bunches):
    conv.EnqueueSteps(bunch, ref)
conv.Finish()
dt = time.perf_counter() - t0
table, _ = conv.GetTable()
entries_per_photon = float(table.sum()) / ((bunches + 1) * n * pps)
print(json.dumps({"workload": "table-maker variant, default spherical axes (%d bins), SpiceMie, cascade steps" % axes.GetNBins(),
                  "kernel": "fast persistent" if conv.kernelMode == 0 else "reference-order",
                  "photons": bunches * n * pps, "seconds": dt, "photons_per_s": bunches * n * pps / dt,
                  "summed_weight_per_photon": entries_per_photon}))
