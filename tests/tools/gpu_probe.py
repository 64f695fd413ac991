"""First-contact GPU script: reference-order kernel vs oracle, fast kernel rate, OpenCL ICD probe."""
import ctypes, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from clsim_b200 import capi, geometry, ice, steps
from clsim_b200.description import ConverterOptions, KERNEL_FAST, KERNEL_REFERENCE
from oracle import pyoracle
from tests.scenes import make_scene, rng_streams, sort_photons

print("nproc", os.cpu_count())
print("OpenCL vendors:", os.listdir("/etc/OpenCL/vendors") if os.path.isdir("/etc/OpenCL/vendors") else None)
try:
    cl = ctypes.CDLL("libOpenCL.so.1")
    n = ctypes.c_uint(0)
    rc = cl.clGetPlatformIDs(0, None, ctypes.byref(n))
    print("clGetPlatformIDs rc", rc, "platforms", n.value)
except OSError as e:
    print("no libOpenCL:", e)

sc = make_scene("spice_mie")
bunch = steps.muon_track_steps(4096, seed=2)
a, x = rng_streams(len(bunch))
opt = sc.options(kernel_mode=KERNEL_REFERENCE, max_num_workitems=len(bunch), rng_n=len(bunch), rng_a=a, rng_x=x)
t = time.time()
with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
    print("engine create %.2fs" % (time.time() - t))
    t = time.time()
    eng.enqueue(bunch, 5)
    r = eng.get_result()
    print("reference kernel: %d hits in %.3fs, stats %s" % (len(r.photons), time.time() - t, eng.statistics()))
    x_after, _ = eng.rng_get(len(bunch))
osc = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, opt)
t = time.time()
want, cnt, st, x_or, _ = osc.propagate(bunch, x, a, num_threads=os.cpu_count())
print("oracle: %d hits in %.2fs (%d photons, %.1f seg/photon)" % (cnt, time.time() - t, st["photons"], st["segments"] / st["photons"]))
print("rng states equal after run: %.4f" % np.mean(x_after == x_or))
g, w = sort_photons(r.photons), sort_photons(want)
if len(g) == len(w):
    same = (g["string_id"] == w["string_id"]) & (g["om_id"] == w["om_id"])
    print("same DOM:", same.mean(), "max |dt|", np.abs(g["t"] - w["t"])[same].max(), "max |dpos|",
          max(np.abs(g[k] - w[k])[same].max() for k in ("x", "y", "z")))
    print("bit-identical records:", np.mean([g[i].tobytes() == w[i].tobytes() for i in range(len(g))]))
else:
    print("hit count differs", len(g), len(w))

for n_steps in (1 << 14, 1 << 17, 1 << 20):
    big = steps.muon_track_steps(n_steps, seed=3)
    opt = sc.options(kernel_mode=KERNEL_FAST, max_num_workitems=n_steps, rng_seed=9)
    t = time.time()
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        print("fast engine create %.2fs" % (time.time() - t))
        eng.upload_resident(big)
        eng.run_resident(1)
        res = eng.run_resident(2)
        print("fast kernel n=%d: %.3e photons/s, hits/photon %.5f, seg/photon %.2f, %.2f ms/launch" % (
            n_steps, res["photons"] / (res["kernel_ms"] * 1e-3), res["hits"] / res["photons"], res["segments"] / res["photons"], res["kernel_ms"] / 2))
    if n_steps == 1 << 17:
        opt = sc.options(kernel_mode=KERNEL_REFERENCE, max_num_workitems=n_steps, rng_seed=9)
        with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
            eng.upload_resident(big)
            eng.run_resident(1)
            res = eng.run_resident(1)
            print("reference-order kernel n=%d: %.3e photons/s, hits/photon %.5f seg/photon %.2f" % (
                n_steps, res["photons"] / (res["kernel_ms"] * 1e-3), res["hits"] / res["photons"], res["segments"] / res["photons"]))
