"""Photon -> MCPE conversion without a GPU: the oracle against the golden inputs generated
from the reference, the semantics of the two converters, and the C ABI's argument checks."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from clsim_b200 import capi, ice, mcpe
from clsim_b200.description import PHOTON_DTYPE
from oracle import mcpe_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_angular():
    with open(os.path.join(GOLDEN, "angular_acceptance.json")) as f:
        return json.load(f)


def photons_on_sphere(n, radius, seed, wl=(300e-9, 600e-9)):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=PHOTON_DTYPE)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    p["x"], p["y"], p["z"] = (radius * v).T
    p["t"] = rng.uniform(0, 3000, n)
    p["theta"] = np.arccos(rng.uniform(-1, 1, n))
    p["phi"] = rng.uniform(0, 2 * math.pi, n)
    p["wavelength"] = rng.uniform(wl[0], wl[1], n)
    p["weight"] = 1.0
    p["identifier"] = rng.integers(0, 50, n)
    p["string_id"] = rng.integers(1, 87, n)
    p["om_id"] = rng.integers(1, 61, n)
    p["group_velocity"] = 0.22
    return p


def test_angular_polynomial_matches_reference_file():
    g = golden_angular()
    assert g["peak"] == 0.75 and len(g["coefficients"]) == 11
    f = mcpe.GetIceCubeDOMAngularSensitivity()
    assert f.coefficients == g["coefficients"]
    assert mcpe.GetHoleIcePeak() == g["peak"]
    got = mcpe_oracle.polynomial(g["coefficients"], g["cos"])
    assert np.allclose(got, g["value"], rtol=0, atol=1e-15)
    assert np.allclose([f.GetValue(x) for x in g["cos"]], g["value"], rtol=0, atol=1e-15)
    # the curve never exceeds its advertised peak on [-1, 1] and is non-negative
    dense = mcpe_oracle.polynomial(g["coefficients"], np.linspace(-1, 1, 4001))
    assert dense.max() <= g["peak"] + 1e-3 and dense.min() >= 0.0


def test_table_lookup_equals_host_twin():
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * 5.0)
    with open(os.path.join(GOLDEN, "dom_acceptance.json")) as f:
        g = json.load(f)
    # the golden table is for the plain DOM radius; ours scales with 1/r^2
    assert np.allclose(np.asarray(acc.values) * 25.0, g["values"], rtol=1e-12)
    wl = np.concatenate([np.linspace(200e-9, 750e-9, 2001), [260e-9, 680e-9, 259.9999e-9, 680.0001e-9]])
    got = mcpe_oracle.from_table(acc.values, acc.start_wlen, acc.wlen_step, wl)
    want = np.array([acc.GetValue(w) for w in wl])
    assert np.array_equal(got, want)


def test_inloop_converter_semantics():
    g = golden_angular()
    icecube = ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * 5.0, efficiency=0.9 * g["peak"])
    ratio = (np.array([250.0, 400.0, 700.0]), np.array([1.30, 1.35, 1.40]))  # synthetic stand-in for ice-models' wv.rde
    deepcore = ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * 5.0, efficiency=0.9 * g["peak"], highQE=True, highQERatio=ratio)
    env = ice.envelope([icecube, deepcore])
    assert np.array_equal(env.values, deepcore.values)
    acc_of = {(s, o): (deepcore if s > 78 else icecube) for s in range(1, 87) for o in range(1, 61)}
    p = photons_on_sphere(20000, 0.1651, 3)
    p["weight"] = (1.0 / mcpe_oracle.acceptance_value(env, p["wavelength"].astype(np.float64))).astype(np.float32)
    p["weight"][:100] = 0.0
    u = np.random.default_rng(4).uniform(size=len(p)).astype(np.float32)
    keep, prob, t = mcpe_oracle.convert_inloop(p, acc_of, g["coefficients"], u)
    assert not keep[:100].any()
    live = p["weight"] != 0
    # weighted photons: probability = angular(-cos theta) * (own acceptance / envelope)
    ang = mcpe_oracle.polynomial(g["coefficients"], np.clip(-np.cos(p["theta"].astype(np.float64)), -1, 1))
    dc = p["string_id"] > 78
    assert np.allclose(prob[live & dc], ang[live & dc], rtol=1e-6)
    own = mcpe_oracle.acceptance_value(icecube, p["wavelength"].astype(np.float64)) / mcpe_oracle.acceptance_value(env, p["wavelength"].astype(np.float64))
    assert np.allclose(prob[live & ~dc], (ang * own)[live & ~dc], rtol=1e-6)
    assert np.array_equal(keep, live & (prob > u))
    assert np.array_equal(t, p["t"].astype(np.float64))       # no time correction in the in-loop converter
    assert 0.2 < keep.mean() < 0.5
    # the reference's fatal conditions
    bad = p.copy(); bad["weight"][5] = -1.0
    with pytest.raises(mcpe_oracle.Fatal, match="negative weight"):
        mcpe_oracle.convert_inloop(bad, acc_of, g["coefficients"], u)
    bad = p.copy(); bad["weight"][500] = 1e5; bad["theta"][500] = math.pi
    with pytest.raises(mcpe_oracle.Fatal, match="hitProbability"):
        mcpe_oracle.convert_inloop(bad, acc_of, g["coefficients"], u)
    bad = p.copy()
    for k in "xyz":
        bad[k][700] *= 1.3
    with pytest.raises(mcpe_oracle.Fatal, match="distance"):
        mcpe_oracle.convert_inloop(bad, acc_of, g["coefficients"], u)
    bad = p.copy(); bad["string_id"][900] = 99
    with pytest.raises(mcpe_oracle.Fatal, match="No wavelength acceptance"):
        mcpe_oracle.convert_inloop(bad, acc_of, g["coefficients"], u)


def test_module_converter_time_correction_and_efficiency():
    g = golden_angular()
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * 5.0)
    p = photons_on_sphere(5000, 0.1651 * 5.0, 7)
    p["weight"] = (1.0 / mcpe_oracle.acceptance_value(acc, p["wavelength"].astype(np.float64))).astype(np.float32)
    eff = {(s, o): (1.35 if s > 78 else 1.0) for s in range(1, 87) for o in range(1, 61)}
    u = np.random.default_rng(8).uniform(size=len(p)).astype(np.float32)
    keep, prob, t = mcpe_oracle.convert_module(p, acc, g["coefficients"], eff, u, oversize=5.0, pancake=1.0)
    # bringForward = (p . d) (1 - pancake/oversize) / v_g with p pointing from the photon to the DOM centre
    th, ph = p["theta"].astype(np.float64), p["phi"].astype(np.float64)
    d = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], 1)
    r = np.stack([p["x"], p["y"], p["z"]], 1).astype(np.float64)
    want = p["t"].astype(np.float64) + (-(r * d).sum(1)) * 0.8 / float(np.float32(0.22))
    assert np.allclose(t, want, rtol=0, atol=1e-9)
    assert np.abs(t - p["t"]).max() < 0.8255 * 0.8 / 0.22 + 1e-6
    # pancake == oversize: no correction, no position check (positions on the small sphere are fine)
    q = photons_on_sphere(100, 0.1651, 9)
    q["weight"] = 0.5
    k2, _, t2 = mcpe_oracle.convert_module(q, acc, g["coefficients"], eff, np.zeros(100, np.float32), oversize=5.0, pancake=5.0)
    assert np.array_equal(t2, q["t"].astype(np.float64))
    with pytest.raises(mcpe_oracle.Fatal, match="distance"):
        mcpe_oracle.convert_module(q, acc, g["coefficients"], eff, np.zeros(100, np.float32), oversize=5.0, pancake=1.0)
    # DeepCore efficiency scales the probability
    dc = p["string_id"] > 78
    base = mcpe_oracle.convert_module(p, acc, g["coefficients"], {k: 1.0 for k in eff}, u, oversize=5.0, pancake=1.0)[1]
    assert np.allclose(prob[dc], 1.35 * base[dc]) and np.array_equal(prob[~dc], base[~dc])


def test_mwc_draw_assignment():
    a = capi.safeprime_multipliers(100, 8)
    x = (np.arange(8, dtype=np.uint64) + np.uint64(12345)) * np.uint64(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF) % (a.astype(np.uint64) << np.uint64(32))
    u, x_after = mcpe_oracle.mwc_uniforms(x, a, 21)
    # scalar restatement of mwcrng_kernel.cl:12-28
    xs = [int(v) for v in x]
    want = []
    for j in range(21):
        s = j % 8
        xs[s] = (xs[s] & 0xFFFFFFFF) * int(a[s]) + (xs[s] >> 32)
        w = xs[s] & 0xFFFFFFFF
        # round toward zero to 24 bits
        if w >= (1 << 24):
            sh = w.bit_length() - 24
            w = (w >> sh) << sh
        want.append(np.float32(w) * np.float32(2.3283064365386963e-10))
    assert np.array_equal(u, np.array(want, dtype=np.float32))
    assert [int(v) for v in x_after] == xs
    assert u.max() < 1.0


def test_abi_argument_checks_without_gpu(has_gpu):
    lib = mcpe._lib()
    h = C.c_void_p()
    cfg = mcpe.McpeConfigStruct()
    cfg.struct_size = 4
    assert lib.clsimcu_mcpe_create(C.byref(cfg), C.byref(h)) == -1
    assert b"struct_size" in lib.clsimcu_last_error()
    cfg.struct_size = C.sizeof(mcpe.McpeConfigStruct)
    assert lib.clsimcu_mcpe_create(C.byref(cfg), C.byref(h)) == -1
    assert b"WavelengthAcceptance" in lib.clsimcu_last_error()
    n = C.c_size_t(0)
    assert lib.clsimcu_mcpe_convert(None, None, 0, None, None, 0, C.byref(n)) == -4
    assert lib.clsimcu_attach_mcpe_converter(None, None, 0) == -4
    if not has_gpu:
        acc = ice.GetIceCubeDOMAcceptance()
        with pytest.raises(capi.ClsimCudaError) as e:
            mcpe.I3CLSimPhotonToMCPEConverterForDOMs(1, {(1, 1): acc}, mcpe.GetIceCubeDOMAngularSensitivity())
        assert e.value.code == -3 and "no CPU fallback" in str(e.value)


# ----------------------------------------------------------------------------------------------------------------
# ... against the reference's own converters: private/clsim/dom/I3PhotonToMCPEConverter.cxx compiled unmodified
# (oracle/_ref/libclsim_ref_mcpe.so; IceTray's module protocol and data classes are stand-ins, oracle/ref_shim/host_mcpe/)
# ----------------------------------------------------------------------------------------------------------------
from oracle import pyoracle  # noqa: E402

needs_ref = pytest.mark.skipif(not pyoracle.ref_mcpe_available(), reason="oracle/_ref not built (no /root/reference at build time)")


@needs_ref
def test_inloop_converter_equals_the_references():
    """I3CLSimPhotonToMCPEConverterForDOMs::Convert: the same survivors, the same times, from the same uniforms -- 20 000
    photons, weights that put the probability anywhere in (0, 1], some weight-0 photons (which draw nothing)."""
    g = golden_angular()
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * 5.0, efficiency=0.9 * g["peak"])
    p = photons_on_sphere(20000, 0.1651, 3)
    rng = np.random.default_rng(4)
    p["weight"] = (rng.uniform(0.05, 1.0, len(p)) / mcpe_oracle.acceptance_value(acc, p["wavelength"].astype(np.float64))).astype(np.float32)
    p["weight"][:100] = 0.0
    u = rng.uniform(size=len(p))
    keep, prob, t = mcpe_oracle.convert_inloop(p, {(s, o): acc for s in range(1, 87) for o in range(1, 61)}, g["coefficients"], u)
    ref_keep, ref_t, used = pyoracle.ref_mcpe_convert_inloop(p, acc, g["coefficients"], u)
    assert np.array_equal(ref_keep, keep) and 2000 < keep.sum() < 12000
    assert np.array_equal(ref_t[keep], t[keep])
    assert used == int((p["weight"] != 0).sum())                  # one draw per photon that has a weight
    # a uniform EQUAL to the probability does not survive (hitProbability <= Uniform(): out)
    u2 = u.copy()
    u2[200:300] = prob[200:300]
    assert not pyoracle.ref_mcpe_convert_inloop(p, acc, g["coefficients"], u2)[0][200:300].any()
    assert not mcpe_oracle.convert_inloop(p, {(s, o): acc for s in range(1, 87) for o in range(1, 61)}, g["coefficients"], u2)[0][200:300].any()
    # the reference's fatal conditions are the oracle's
    for field, value, what in (("weight", -1.0, "negative weight"), ("weight", 1e5, "cannot continue"), ("x", 0.4, "distance not")):
        bad = p.copy()
        bad[field][500] = value
        if field == "weight" and value > 0:
            bad["theta"][500] = math.pi
        with pytest.raises(RuntimeError, match=what):
            pyoracle.ref_mcpe_convert_inloop(bad, acc, g["coefficients"], u)
        with pytest.raises(mcpe_oracle.Fatal):
            mcpe_oracle.convert_inloop(bad, {(s, o): acc for s in range(1, 87) for o in range(1, 61)}, g["coefficients"], u)


def _module_order(p):
    """The order in which the module meets the photons: DOMs in (string, om) order, photons of a DOM in input order."""
    return np.lexsort((np.arange(len(p)), p["om_id"], p["string_id"]))


@needs_ref
@pytest.mark.parametrize("oversize,pancake", [(5.0, 1.0), (5.0, 5.0), (1.0, 1.0), (16.0, 4.0)])
def test_module_converter_equals_the_references(oversize, pancake):
    """The I3PhotonToMCPEConverter module on a frame (absolute positions, geometry, calibration): survivors, corrected times and
    the per-DOM time ordering of its output against convert_module."""
    g = golden_angular()
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * oversize)
    n = 6000
    p = photons_on_sphere(n, 0.1651 * oversize, 7)
    rng = np.random.default_rng(8)
    p["weight"] = (rng.uniform(0.05, 0.7, n) / mcpe_oracle.acceptance_value(acc, p["wavelength"].astype(np.float64))).astype(np.float32)
    p["weight"][::97] = 0.0
    p["group_velocity"] = rng.uniform(0.21, 0.23, n).astype(np.float32)
    p["num_scatters"] = 1      # (an UNSCATTERED I3Photon is also checked against its start position, see the last test)
    eff_of = {(s, o): (1.35 if s > 78 else 1.0) for s in range(1, 87) for o in range(1, 61)}
    u = rng.uniform(size=n)
    keep, prob, t = mcpe_oracle.convert_module(p, acc, g["coefficients"], eff_of, u, oversize=oversize, pancake=pancake)
    # the frame: every DOM somewhere in the detector, the calibration carrying the efficiencies
    dom_pos = np.stack([p["string_id"] * 10.0 - 400.0, (p["string_id"] % 7) * 30.0, 500.0 - 17.0 * p["om_id"]], axis=1)
    eff = np.array([eff_of[(int(s), int(o))] for s, o in zip(p["string_id"], p["om_id"])])
    order = _module_order(p)
    in_order = order[p["weight"][order] != 0]
    s, o, rt, idx, used = pyoracle.ref_mcpe_convert_module(p, dom_pos, acc, g["coefficients"], eff, u[in_order], oversize=oversize, pancake=pancake)
    assert used == len(in_order)
    assert sorted(idx.tolist()) == sorted(np.nonzero(keep)[0].tolist()) and 500 < len(idx) < 4000
    assert np.array_equal(s, p["string_id"][idx]) and np.array_equal(o, p["om_id"][idx])
    # (the module subtracts absolute positions, the oracle works in DOM-relative ones: the last bits of the dot product differ)
    assert np.abs(rt - t[idx]).max() < 2e-10 * 3000
    # output order: DOMs in key order, within a DOM by time
    key = s.astype(np.int64) * 1000 + o
    assert np.all(np.diff(key) >= 0)
    same = np.diff(key) == 0
    assert np.all(np.diff(rt)[same] >= 0)


@needs_ref
def test_module_converter_efficiency_sources_and_position_check():
    g = golden_angular()
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651)
    p = photons_on_sphere(400, 0.1651, 11)
    p["weight"] = 0.3
    p["string_id"], p["om_id"] = 5, np.arange(400) % 60 + 1
    dom_pos = np.zeros((400, 3))
    dom_pos[:, 2] = -17.0 * p["om_id"]
    order = _module_order(p)
    u = np.full(400, 0.05)
    # (the module takes the angle against the PMT axis, -(d . (0, 0, -1)) = +cos(theta); the in-loop converter takes -cos(theta))
    ang = mcpe_oracle.polynomial(g["coefficients"], np.clip(np.cos(p["theta"].astype(np.float64)), -1, 1))
    base = 0.3 * mcpe_oracle.acceptance_value(acc, p["wavelength"].astype(np.float64)) * ang
    # no calibration entry: the default applies; ReplaceRelativeDOMEfficiencyWithDefault: the default always applies
    for eff, default, replace, factor in ((np.full(400, np.nan), 0.5, False, 0.5), (np.full(400, 2.0), 0.5, True, 0.5), (np.full(400, 2.0), 0.5, False, 2.0)):
        _, _, _, idx, _ = pyoracle.ref_mcpe_convert_module(p, dom_pos, acc, g["coefficients"], eff, u[order], default_efficiency=default,
                                                          replace_with_default=replace)
        assert sorted(idx.tolist()) == np.nonzero(base * factor > 0.05)[0].tolist()
    # a photon off the sphere is fatal when the pancake factor is 1 -- unless the check is made a warning
    q = p.copy()
    q["x"][7] += 0.2
    with pytest.raises(RuntimeError, match="distance not"):
        pyoracle.ref_mcpe_convert_module(q, dom_pos, acc, g["coefficients"], np.ones(400), u[order])
    pyoracle.ref_mcpe_convert_module(q, dom_pos, acc, g["coefficients"], np.ones(400), u[order], only_warn=True)
    with pytest.raises(mcpe_oracle.Fatal, match="distance"):
        mcpe_oracle.convert_module(q, acc, g["coefficients"], {(5, o): 1.0 for o in range(1, 61)}, u)
    # CheckSanity<I3Photon> (:213-231), a consistency check of the full photon record that neither the oracle nor the CUDA
    # converter restates (they see the compressed record): an unscattered photon more than a metre from its start must point
    # away from it
    far = photons_on_sphere(50, 0.1651 * 16.0, 12)
    far["weight"], far["string_id"], far["om_id"] = 0.01, 5, np.arange(50) + 1
    far["theta"], far["phi"] = np.pi / 2, 0.0                      # all travel along +x; the start is the DOM centre
    far["x"], far["y"], far["z"] = 0.0, 0.1651 * 16.0, 0.0         # ... but they sit at +y
    with pytest.raises(RuntimeError, match="unscattered photon direction is inconsistent"):
        pyoracle.ref_mcpe_convert_module(far, np.zeros((50, 3)), acc, g["coefficients"], np.ones(50), np.full(50, 0.99), oversize=16.0, pancake=16.0)
    far["num_scatters"] = 3
    pyoracle.ref_mcpe_convert_module(far, np.zeros((50, 3)), acc, g["coefficients"], np.ones(50), np.full(50, 0.99), oversize=16.0, pancake=16.0)
