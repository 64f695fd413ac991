"""BASELINE config 4 end to end on one GPU: a 1 PeV cascade in SpiceLea ice (tilt + anisotropy), steps made on the
device from the cascade's step-generation queue entry, photo-electrons out (photons never leave the device).
usage (GPU box): python tools/config4_cascade.py [energy_GeV] [bunch_steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clsim_b200 import capi, geometry, ice, mcpe, stepgen
from clsim_b200.description import KERNEL_FAST, ConverterOptions
from clsim_b200.sharding import mcpe_row_offset, stepgen_row_offset

energy = float(sys.argv[1]) if len(sys.argv) > 1 else 1e6
bunch = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=True)
geo = geometry.make_ic86_like_geometry(oversize=5.0)
ang = mcpe.GetIceCubeDOMAngularSensitivity()
acc = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0, efficiency=0.9 * mcpe.GetHoleIcePeak())
gen = ice.makeCherenkovWavelengthGenerator(acc, False, medium)
opt = ConverterOptions(stop_detected_photons=True, pancake_factor=5.0, kernel_mode=KERNEL_FAST, enable_double_buffering=True,
                       max_num_workitems=bunch, rng_seed=1, output_photons_per_workitem=2)
conv = stepgen.I3CLSimLightSourceToStepConverterPPC(photonsPerStep=200)
conv.SetMediumProperties(medium)
conv.SetWlenBias(acc)
conv.SetRandomService(4)
conv.Initialize(rngFirstMultiplierRow=stepgen_row_offset(0))
pe = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(5, {(int(s), int(o)): acc for s, o in zip(geo.stringIDs, geo.domIDs)}, ang,
                                             rngFirstMultiplierRow=mcpe_row_offset(0))
with capi.Engine(medium, geo, [gen], acc, opt) as eng:
    pe.attach_to(eng)
    # warm up: one small cascade
    conv.EnqueueLightSource(stepgen.Particle("EMinus", 1e3, (20.0, -30.0, -250.0), (0.3, 0.2, -0.93)), 0)
    while conv.EnqueueInto(eng, 0):
        eng.get_result()
    conv.EnqueueLightSource(stepgen.Particle("EMinus", energy, (20.0, -30.0, -250.0), (0.3, 0.2, -0.93)), 1)
    conv.EnqueueBarrier()
    t0 = time.perf_counter()
    sent = pending = photons = hits = pes = 0
    while True:
        n = conv.EnqueueInto(eng, 100 + sent)
        if n == 0:
            break
        sent += 1
        pending += 1
        while eng.more_photons_available():
            r = eng.get_result(); pending -= 1
            photons += r.num_photons_generated; hits += r.num_hits_counted; pes += len(r.mcpes)
    while pending:
        r = eng.get_result(); pending -= 1
        photons += r.num_photons_generated; hits += r.num_hits_counted; pes += len(r.mcpes)
    dt = time.perf_counter() - t0
print(json.dumps({"workload": "config4: %.3g GeV e- cascade, SpiceLea tilt+anisotropy, IC86-like, oversize 5, steps made on the device, MCPEs out" % energy,
                  "bunches": sent, "photons": photons, "hits": hits, "mcpes": pes, "seconds": dt, "photons_per_s": photons / dt,
                  "mean_photons_per_m_biased": conv.meanPhotonsPerMeter}))
