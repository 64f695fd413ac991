"""Photon -> MCPE on the device against the oracle: the stand-alone conversion with explicit
uniforms (every survivor, bit for bit), with the device's MWC streams (draw assignment replayed by the oracle),
both converter flavours, the reference's fatal conditions, and the in-stream conversion attached to an engine."""
import math

import numpy as np
import pytest

from clsim_b200 import capi, geometry, ice, mcpe, steps
from clsim_b200.description import PHOTON_DTYPE
from oracle import mcpe_oracle
from clsim_b200.sharding import mcpe_row_offset, stepgen_row_offset
from tests.scenes import DEVICE_KERNEL, Scene, make_scene, sized
from tests.test_mcpe_oracle import golden_angular, photons_on_sphere

pytestmark = pytest.mark.gpu

N_STEPS = sized(1 << 15, 1 << 12)     # steps per bunch where a test only needs detected photons
MIN_HITS = sized(3000, 300)
RATIO = (np.array([250.0, 400.0, 700.0]), np.array([1.30, 1.35, 1.40]))  # synthetic stand-in for ice-models' wv.rde


def detector(oversize=5.0, unshadowed=0.9):
    """Acceptances the way traysegments/common.py:181-213 sets them up: two DOM classes, bias = envelope."""
    g = golden_angular()
    eff = unshadowed * g["peak"]
    icecube = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * oversize, efficiency=eff)
    deepcore = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * oversize, efficiency=eff, highQE=True, highQERatio=RATIO)
    sc = make_scene("spice_mie", oversize=oversize)
    sc.bias = ice.envelope([icecube, deepcore])
    sc.generators = [ice.makeCherenkovWavelengthGenerator(sc.bias, False, sc.medium)]
    acc_of = {(int(s), int(o)): (deepcore if s > 78 else icecube) for s, o in zip(sc.geo.stringIDs, sc.geo.domIDs)}
    return sc, acc_of, mcpe.GetIceCubeDOMAngularSensitivity()


def detected_photons(sc, n_steps=N_STEPS, seed=21, pancake=None):
    opt = sc.options(kernel_mode=DEVICE_KERNEL, max_num_workitems=n_steps, rng_seed=5)
    if pancake is not None:
        opt = sc.options(kernel_mode=DEVICE_KERNEL, max_num_workitems=n_steps, rng_seed=5, pancake_factor=pancake)
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        eng.enqueue(steps.muon_track_steps(n_steps, seed=seed), 1)
        return eng.get_result().photons


def as_set(m):
    return sorted(zip(m["identifier"].tolist(), m["string_id"].tolist(), m["om_id"].tolist(), m["time"].tolist(), m["npe"].tolist()))


def expected(photons, keep, time):
    out = np.zeros(int(keep.sum()), dtype=mcpe.MCPE_DTYPE)
    sel = photons[keep]
    out["string_id"], out["om_id"], out["identifier"] = sel["string_id"], sel["om_id"], sel["identifier"]
    out["time"] = time[keep].astype(np.float32)
    out["npe"] = 1
    return out


def test_inloop_converter_on_propagated_photons_explicit_uniforms():
    sc, acc_of, ang = detector()
    photons = detected_photons(sc)
    assert len(photons) > MIN_HITS
    # the propagation kernel leaves the photon on the surface of the real-size DOM (pancake undone)
    r = np.sqrt(photons["x"].astype(np.float64) ** 2 + photons["y"].astype(np.float64) ** 2 + photons["z"].astype(np.float64) ** 2)
    assert np.abs(r - 0.1651).max() < 0.005
    conv = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(11, acc_of, ang)
    u = np.random.default_rng(1).uniform(size=len(photons)).astype(np.float32)
    got = conv.Convert(photons, uniforms=u)
    keep, prob, t = mcpe_oracle.convert_inloop(photons, acc_of, ang.coefficients, u)
    assert prob.max() <= 1.0 and 0.05 < keep.mean() < 0.6
    assert as_set(got) == as_set(expected(photons, keep, t))     # every survivor, exact times
    # u == 0 keeps every photon with p > 0; u just below 1 keeps none (p <= peak < 1)
    assert len(conv.Convert(photons, uniforms=np.zeros(len(photons), np.float32))) == int((prob > 0).sum())
    assert len(conv.Convert(photons, uniforms=np.full(len(photons), 0.999, np.float32))) == 0
    assert len(conv.Convert(photons[:0])) == 0
    conv.close()


def test_device_rng_draw_assignment_is_replayable():
    sc, acc_of, ang = detector()
    photons = np.concatenate([detected_photons(sc, seed=31), detected_photons(sc, seed=32)])
    conv = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(12345, acc_of, ang, rngFirstMultiplierRow=mcpe_row_offset(1))
    x0, a = conv.rng_state()
    assert len(x0) == len(a) and len(x0) % 256 == 0 and len(np.unique(a)) == len(a)
    assert np.array_equal(a, capi.safeprime_multipliers(mcpe_row_offset(1), len(a)))
    big = np.concatenate([photons] * 8)      # several draws per stream
    assert len(big) > 2 * len(x0)
    got = conv.Convert(big)
    u, x1 = mcpe_oracle.mwc_uniforms(x0, a, len(big))
    keep, _, t = mcpe_oracle.convert_inloop(big, acc_of, ang.coefficients, u)
    assert as_set(got) == as_set(expected(big, keep, t))
    assert np.array_equal(conv.rng_state()[0], x1)
    # a second conversion continues the streams
    got2 = conv.Convert(photons)
    u2, _ = mcpe_oracle.mwc_uniforms(x1, a, len(photons))
    keep2, _, t2 = mcpe_oracle.convert_inloop(photons, acc_of, ang.coefficients, u2)
    assert as_set(got2) == as_set(expected(photons, keep2, t2))
    conv.close()


def test_module_converter_oversized_spheres_time_correction():
    # DOMOversizeFactor 5 without pancake: photons sit on the 0.8255 m sphere and are brought forward in time
    g = golden_angular()
    sc = make_scene("spice_mie", oversize=5.0)
    photons = detected_photons(sc, pancake=1.0)
    r = np.sqrt(photons["x"].astype(np.float64) ** 2 + photons["y"].astype(np.float64) ** 2 + photons["z"].astype(np.float64) ** 2)
    assert np.abs(r - 0.8255).max() < 0.01
    keys = [(int(s), int(o)) for s, o in zip(sc.geo.stringIDs, sc.geo.domIDs)]
    rde = {k: (1.2 if k[0] > 78 else 1.0) for k in keys}
    acc = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0, efficiency=0.9 * g["peak"] / 1.2)
    ang = mcpe.GetIceCubeDOMAngularSensitivity()
    conv = mcpe.I3PhotonToMCPEConverter(3, sc.geo, acc, ang, DOMOversizeFactor=5.0, DOMPancakeFactor=1.0, RelativeDOMEfficiencies=rde)
    u = np.random.default_rng(2).uniform(size=len(photons)).astype(np.float32)
    got = conv.Convert(photons, uniforms=u)
    keep, prob, t = mcpe_oracle.convert_module(photons, acc, ang.coefficients, rde, u, oversize=5.0, pancake=1.0)
    assert as_set(got) == as_set(expected(photons, keep, t))
    shift = t - photons["t"]
    assert shift.min() > -1e-3 and 1.0 < shift.max() < 0.8255 * 0.8 / 0.2 # brought FORWARD along the direction of travel
    conv.close()
    # same photons, converter told the DOMs are pancakes: no position check, no time shift
    conv = mcpe.I3PhotonToMCPEConverter(3, sc.geo, acc, ang, DOMOversizeFactor=5.0, DOMPancakeFactor=5.0, RelativeDOMEfficiencies=rde)
    got = conv.Convert(photons, uniforms=u)
    keep, _, t = mcpe_oracle.convert_module(photons, acc, ang.coefficients, rde, u, oversize=5.0, pancake=5.0)
    assert as_set(got) == as_set(expected(photons, keep, t)) and np.array_equal(t, photons["t"].astype(np.float64))
    conv.close()
    # spherical DOMs of the wrong size: fatal in the reference (…cxx:416-451), an error here
    conv = mcpe.I3PhotonToMCPEConverter(3, sc.geo, acc, ang, DOMOversizeFactor=1.0, DOMPancakeFactor=1.0, RelativeDOMEfficiencies=rde)
    with pytest.raises(capi.ClsimCudaError, match="distance not"):
        conv.Convert(photons, uniforms=u)
    conv.close()
    conv = mcpe.I3PhotonToMCPEConverter(3, sc.geo, acc, ang, DOMOversizeFactor=1.0, DOMPancakeFactor=1.0, RelativeDOMEfficiencies=rde,
                                        OnlyWarnAboutInvalidPhotonPositions=True)
    assert len(conv.Convert(photons, uniforms=u)) == int(keep.sum())
    conv.close()


def test_fatal_conditions_are_errors():
    sc, acc_of, ang = detector()
    conv = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(1, acc_of, ang)
    p = photons_on_sphere(4096, 0.1651, 5)
    p["weight"] = 100.0
    u = np.full(len(p), 0.1, np.float32)
    want = int(mcpe_oracle.convert_inloop(p, acc_of, ang.coefficients, u)[0].sum())
    assert want > 100 and len(conv.Convert(p, uniforms=u)) == want
    bad = p.copy(); bad["weight"][7] = -1.0
    with pytest.raises(capi.ClsimCudaError, match="negative weight"):
        conv.Convert(bad, uniforms=u)
    bad = p.copy(); bad["weight"][9] = 1e9; bad["theta"][9] = math.pi
    with pytest.raises(capi.ClsimCudaError, match="hit weights are too high"):
        conv.Convert(bad, uniforms=u)
    bad = p.copy()
    for k in "xyz":
        bad[k][11] *= 1.3
    with pytest.raises(capi.ClsimCudaError, match="distance not"):
        conv.Convert(bad, uniforms=u)
    bad = p.copy(); bad["string_id"][13] = 99
    with pytest.raises(capi.ClsimCudaError, match="No wavelength acceptance"):
        conv.Convert(bad, uniforms=u)
    zero = p.copy(); zero["weight"] = 0.0
    assert len(conv.Convert(zero, uniforms=np.zeros(len(p), np.float32))) == 0
    conv.close()


def test_conversion_attached_to_the_engine():
    sc, acc_of, ang = detector()
    n_steps = N_STEPS
    opt = sc.options(kernel_mode=DEVICE_KERNEL, max_num_workitems=n_steps, rng_seed=5, enable_double_buffering=True)
    conv = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(77, acc_of, ang, rngFirstMultiplierRow=mcpe_row_offset(2))
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        conv.attach_to(eng, keep_photons=True)
        with pytest.raises(capi.ClsimCudaError, match="attached once"):
            conv.attach_to(eng, keep_photons=True)
        x, a = conv.rng_state()
        for i in range(3):
            eng.enqueue(steps.muon_track_steps(n_steps, seed=40 + i), i)
        results = sorted((eng.get_result() for _ in range(3)), key=lambda r: r.identifier)  # launches run in enqueue order
        for res in results:
            assert len(res.photons) > MIN_HITS and res.num_hits_counted == len(res.photons)
            u, x = mcpe_oracle.mwc_uniforms(x, a, len(res.photons))
            keep, _, t = mcpe_oracle.convert_inloop(res.photons, acc_of, ang.coefficients, u)
            assert as_set(res.mcpes) == as_set(expected(res.photons, keep, t))
    conv.close()
    # photo-electrons only: the photons stay on the device
    conv = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(78, acc_of, ang, rngFirstMultiplierRow=mcpe_row_offset(2))
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        conv.attach_to(eng)
        eng.enqueue(steps.muon_track_steps(n_steps, seed=40), 9)
        res = eng.get_result()
        assert len(res.photons) == 0 and res.num_hits_counted > MIN_HITS
        frac = len(res.mcpes) / float(res.num_hits_counted)
        assert abs(frac - len(results[0].mcpes) / float(len(results[0].photons))) < 0.03
        assert np.all(res.mcpes["npe"] == 1) and np.all(res.mcpes["identifier"] < n_steps)
        # attaching after the first bunch is a life-cycle error
        other = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(79, acc_of, ang)
        with pytest.raises(capi.ClsimCudaError, match="before the first EnqueueSteps"):
            other.attach_to(eng)
        other.close()
    conv.close()


def test_attached_converter_errors_surface_through_the_engine():
    # a converter that does not know the detector's DOMs: fatal in the reference, the engine reports it
    sc, acc_of, ang = detector()
    only_string_1 = {k: v for k, v in acc_of.items() if k[0] == 1}
    conv = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(5, only_string_1, ang)
    opt = sc.options(kernel_mode=DEVICE_KERNEL, max_num_workitems=sized(1 << 14, 1 << 11), rng_seed=5)
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        conv.attach_to(eng)
        eng.enqueue(steps.muon_track_steps(sized(1 << 14, 1 << 11), seed=3), 1)
        with pytest.raises(capi.ClsimCudaError, match="No wavelength acceptance"):
            eng.get_result()
    conv.close()
