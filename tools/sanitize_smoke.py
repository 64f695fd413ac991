"""Small end-to-end pass over every kernel of the library, sized for compute-sanitizer (memcheck / racecheck):
fast kernel (plain, tilt+anisotropy, save-all), reference-order kernel (hits, history, table mode), photon -> MCPE,
step generator, attached conversion, device-generated bunch."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clsim_b200 import capi, ice, mcpe, stepgen, steps, tabulator
from clsim_b200.description import KERNEL_FAST, KERNEL_REFERENCE
from tests.scenes import make_scene

n = 2048
for name in ("spice_mie", "spice_lea"):
    sc = make_scene(name)
    for mode in (KERNEL_FAST, KERNEL_REFERENCE):
        opt = sc.options(kernel_mode=mode, max_num_workitems=n, rng_seed=3, photon_history_entries=(3 if mode == KERNEL_REFERENCE else 0))
        with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
            eng.enqueue(steps.muon_track_steps(n, photons_per_step=50, seed=1), 1)
            r = eng.get_result()
            print(name, mode, "hits", len(r.photons))
    # the fast kernel's rare variants: non-stop detection, photon history (with and without stopping), both together
    for stop, hist in ((False, 0), (True, 3), (False, 2)):
        opt = sc.options(kernel_mode=KERNEL_FAST, max_num_workitems=n, rng_seed=6, stop_detected_photons=stop, photon_history_entries=hist, output_photons_per_workitem=4)
        with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
            eng.enqueue(steps.muon_track_steps(n, photons_per_step=50, seed=5), 3)
            r = eng.get_result()
            print(name, "fast stop", stop, "history", hist, "hits", len(r.photons))
    # a bunch small enough to be cut into step parts
    opt = sc.options(kernel_mode=KERNEL_FAST, max_num_workitems=64, rng_seed=7)
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        eng.enqueue(steps.muon_track_steps(64, photons_per_step=333, seed=6), 4)
        print(name, "small bunch hits", len(eng.get_result().photons))
    opt = sc.options(kernel_mode=KERNEL_FAST, stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=0.01, max_num_workitems=n, rng_seed=4)
    with capi.Engine(sc.medium, None, sc.generators, sc.bias, opt) as eng:
        eng.enqueue(steps.muon_track_steps(n, photons_per_step=20, seed=2), 2)
        print(name, "save-all records", len(eng.get_result().photons))
sc = make_scene("spice_mie")
ang = mcpe.GetIceCubeDOMAngularSensitivity()
acc_of = {(int(s), int(o)): sc.bias for s, o in zip(sc.geo.stringIDs, sc.geo.domIDs)}
pe = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(1, acc_of, ang)
gen = stepgen.StepGenerator(rng_seed=2)
opt = sc.options(kernel_mode=KERNEL_FAST, max_num_workitems=n, rng_seed=5)
with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
    pe.attach_to(eng, keep_photons=True)
    gen.enqueue_into(eng, steps.muon_track_sources(n, photons_per_step=50), 7)
    r = eng.get_result()
    print("generated bunch: hits", len(r.photons), "mcpes", len(r.mcpes))
    print("standalone convert", len(mcpe.I3CLSimPhotonToMCPEConverterForDOMs(2, acc_of, ang).Convert(r.photons)))
print("standalone steps", len(gen.generate(steps.muon_track_sources(500, photons_per_step=50))))
axes = tabulator.SphericalAxes([tabulator.PowerAxis(0, 300, 10, 2), tabulator.LinearAxis(0, 180, 4), tabulator.LinearAxis(-1, 1, 5), tabulator.PowerAxis(0, 3e3, 10, 2)])
medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=True)
tab = tabulator.I3CLSimStepToTableConverter(0, axes, 0, True, medium, None, math.pi * 0.1651 ** 2, ice.GetIceCubeDOMAcceptance(), ang, 3, maxNumWorkitems=256)
tab.EnqueueSteps(steps.point_source_steps(256, 5, seed=3), (0, 0, 0, 0, 0, 0, 1))
tab.Finish()
print("table sum (persistent kernel)" if tab.kernelMode == KERNEL_FAST else "table sum", float(tab.GetTable()[0].sum()))
tab.close()
tab = tabulator.I3CLSimStepToTableConverter(0, axes, 0, True, medium, None, math.pi * 0.1651 ** 2, ice.GetIceCubeDOMAcceptance(), ang, 3, maxNumWorkitems=256, kernelMode=KERNEL_REFERENCE)
tab.EnqueueSteps(steps.point_source_steps(256, 5, seed=3), (0, 0, 0, 0, 0, 0, 1))
tab.Finish()
print("table sum (reference-order kernel)", float(tab.GetTable()[0].sum()))
print("sanitize smoke: done")
