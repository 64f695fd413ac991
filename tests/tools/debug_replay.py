import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from clsim_b200 import capi, steps
from clsim_b200.description import KERNEL_FAST
from oracle import pyoracle
from tests.scenes import make_scene
name = sys.argv[1] if len(sys.argv) > 1 else "spice_mie"
sc = make_scene(name)
bunch = steps.muon_track_steps(256, photons_per_step=40, seed=21)
bunch["identifier"] = np.arange(len(bunch))
opt = sc.options(kernel_mode=KERNEL_FAST, stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=1.0,
                 max_num_workitems=len(bunch), output_photons_per_workitem=40, rng_seed=5)
with capi.Engine(sc.medium, None, sc.generators, sc.bias, opt) as eng:
    eng.upload_resident(bunch)
    res = eng.run_resident(1)
    photons = eng.download_resident()
    tags_x, tags_a = eng.download_resident_rng_tags(len(photons))
osc = pyoracle.Scene(sc.medium, None, sc.generators, sc.bias, opt)
rows = []
for i, p in enumerate(photons):
    s = int(p["identifier"])
    saved, q = osc.single_photon_split(bunch[s], tags_x[i, 0], tags_a[i, 0], tags_x[i, 1], tags_a[i, 1])
    dev = max(abs(float(q[k]) - float(p[k])) for k in ("x", "y", "z"))
    rows.append((int(q["num_scatters"]), int(p["num_scatters"]), dev, float(q["cherenkov_dist"]), float(p["cherenkov_dist"]),
                 float(q["dist_in_abs_lens"]), float(p["dist_in_abs_lens"]), float(q["t"]), float(p["t"]), float(q["z"]), float(p["z"])))
r = np.array(rows)
same_sc = r[:, 0] == r[:, 1]
print("same scatter count:", same_sc.mean())
d = r[same_sc, 2]
print("dev percentiles (same scatters):", np.percentile(d, [50, 90, 99, 99.9, 100]))
print("rel path dev:", np.percentile(np.abs(r[same_sc, 3] - r[same_sc, 4]) / np.maximum(1, r[same_sc, 3]), [50, 90, 99, 100]))
print("abs lens dev:", np.percentile(np.abs(r[same_sc, 5] - r[same_sc, 6]), [50, 90, 99, 100]))
print("t dev:", np.percentile(np.abs(r[same_sc, 7] - r[same_sc, 8]), [50, 90, 99, 100]))
bad = ~same_sc
print("different scatter count examples (oracle, gpu, dev, path_o, path_g, abs_o, abs_g):")
for row in r[bad][:15]:
    print("  ", row[:7])
# does the mismatch correlate with scatter count?
for lo, hi in ((0, 5), (5, 15), (15, 30), (30, 60), (60, 1000)):
    sel = (r[:, 0] >= lo) & (r[:, 0] < hi)
    if sel.sum():
        print("oracle scatters [%d,%d): n=%d same=%.3f" % (lo, hi, sel.sum(), same_sc[sel].mean()))
