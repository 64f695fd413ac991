"""One of bench.py's other_configs alone (resident bunch, kernel-only), for an ncu capture:
   ncu --set full --import-source on --clock-control none -k regex:propagate_persistent -s 2 -c 1 -o out python tools/prof_config.py config5 [steps_log2]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clsim_b200 import capi, geometry, ice, steps
from clsim_b200.description import KERNEL_FAST, ConverterOptions

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
lea = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=True)
if name == "config3":
    geo = geometry.make_ic86_like_geometry(oversize=5.0)
    bias = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0)
    gens = [ice.makeCherenkovWavelengthGenerator(bias, False, lea)]
    bunch, pancake = steps.muon_bundle_steps(n, num_muons=100, seed=3), 5.0
elif name == "config5":
    geo = geometry.make_ic86_like_geometry(oversize=1.0)
    bias = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS)
    wl, val = ice.GetFlasherLED405Spectrum()
    gens = [ice.makeCherenkovWavelengthGenerator(bias, False, lea), ice.makeWavelengthGenerator(wl, val, bias, lea)]
    i = int(np.argmin((geo.posX - 0.0) ** 2 + (geo.posY - 0.0) ** 2 + (geo.posZ + 200.0) ** 2))
    bunch, pancake = steps.flasher_steps(n, np.array([geo.posX[i], geo.posY[i], geo.posZ[i]]), seed=5), 1.0
elif name in ("config2", "config2_nonstop", "config2_history"):
    lea = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_mie", useTiltIfAvailable=False)   # (config 2's ice)
    geo = geometry.make_ic86_like_geometry(oversize=5.0)
    bias = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0)
    gens = [ice.makeCherenkovWavelengthGenerator(bias, False, lea)]
    bunch, pancake = steps.muon_track_steps(n, seed=1000), 5.0
else:
    raise SystemExit("config2 | config2_nonstop | config2_history | config3 | config5")
opt = ConverterOptions(device=0, stop_detected_photons=(name != "config2_nonstop"), pancake_factor=pancake, kernel_mode=KERNEL_FAST, max_num_workitems=len(bunch),
                       rng_seed=777, photon_history_entries=(4 if name == "config2_history" else 0), output_photons_per_workitem=2)
with capi.Engine(lea, geo, gens, bias, opt) as eng:
    eng.upload_resident(bunch)
    eng.run_resident(2)
    r = eng.run_resident(2)
print(json.dumps({"config": name, "photons_per_s": r["photons"] / (r["kernel_ms"] * 1e-3), "ms": r["kernel_ms"] / 2, "segments": r["segments"] / 2, "hits": r["hits"] / 2}))
