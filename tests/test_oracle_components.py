"""The oracle, function by function, against (i) the ppc formulas the reference's own tests
restate (golden vectors, tolerances from resources/tests/*.py), (ii) host double-precision twins
of each generated device function (the reference testers' EvaluateReferenceFunction), and
(iii) analytic properties of the samplers.  CPU only."""
import json
import math
import os

import numpy as np
import pytest
from scipy import stats as sps

from clsim_b200 import ice
from oracle import pyoracle
from tests.scenes import make_scene, rng_streams

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def lea():
    sc = make_scene("spice_lea")
    return sc, pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, sc.options())


@pytest.fixture(scope="module")
def mie():
    sc = make_scene("spice_mie")
    return sc, pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, sc.options())


def test_rng_is_the_mwc_recurrence():
    a, x = rng_streams(4)
    u, x_after = pyoracle.rng_uniform_co(x[0], a[0], 1000)
    state = int(x[0])
    ref = []
    for _ in range(1000):
        state = (state & 0xffffffff) * int(a[0]) + (state >> 32)
        lo = state & 0xffffffff
        # round toward zero to 24 bits, then /2^32 (mwcrng_kernel.cl:14-19)
        if lo:
            top = lo.bit_length() - 1
            if top > 23:
                lo &= ~((1 << (top - 23)) - 1)
        ref.append(np.float32(lo) / np.float32(4294967296.0))
    assert state == x_after
    assert np.array_equal(u, np.array(ref, dtype=np.float32))
    assert u.max() < 1.0 and u.min() >= 0.0
    big, _ = pyoracle.rng_uniform_co(x[1], a[1], 200000)
    assert sps.kstest(big, "uniform").pvalue > 1e-3
    assert big.max() < 1.0


def test_rejection_rule_of_seed_states():
    a, x = rng_streams(5000, seed=99)
    hi, lo = (x >> np.uint64(32)).astype(np.uint64), (x & np.uint64(0xffffffff))
    assert np.all(x != 0) and np.all(hi < a.astype(np.uint64) - 1) and np.all(lo < 0xffffffff)


def test_anisotropy_scaling_vs_ppc_formula(lea):
    """testScalarFields.py: relative deviation <= 1e-5 against the ppc formula."""
    g = gold("ppc_formulas.json")
    v = np.array(g["unit_vectors"], dtype=np.float32)
    got = lea[1].eval_scalar_field(1, v)
    want = np.array(g["DimasAbsLenScalingFactor"])
    assert np.abs((got - want) / want).max() <= 1e-5


def test_direction_transforms_vs_ppc(lea):
    """testVectorTransforms.py tolerance 1e-4 (device float vs host double)."""
    g = gold("ppc_formulas.json")
    v = np.array(g["unit_vectors"], dtype=np.float32)
    pre = lea[1].eval_vector_transform(0, v)
    post = lea[1].eval_vector_transform(1, v)
    assert np.abs(pre - np.array(g["PPCPre"])).max() <= 1e-4
    assert np.abs(post - np.array(g["PPCPost"])).max() <= 1e-4
    # pre and post undo each other
    back = lea[1].eval_vector_transform(1, pre)
    assert np.abs(back - v).max() < 1e-5


def _tilt_host(t, x, y, z):
    """Host twin: I3CLSimScalarFieldIceTiltZShift::GetValue (…IceTiltZShift.cxx:113-143)."""
    zc, dist, corr = t["zCoordinates"], t["distancesFromOriginAlongTilt"], t["zCorrections"]
    spacing = (zc[-1] - zc[0]) / (len(zc) - 1)
    lnx, lny = math.cos(t["directionOfTiltAzimuth"]), math.sin(t["directionOfTiltAzimuth"])
    zr = (z - zc[0]) / spacing
    k = int(min(max(math.floor(zr), 0.0), len(zc) - 2))
    fa, fb = zr - k, (k + 1) - zr
    nr = lnx * x + lny * y
    for j in range(1, len(dist)):
        if nr < dist[j] or j == len(dist) - 1:
            w = dist[j] - dist[j - 1]
            lo = (dist[j] - nr) / w
            hi = (nr - dist[j - 1]) / w
            v_lo = corr[j - 1][k + 1] * fa + corr[j - 1][k] * fb
            v_hi = corr[j][k + 1] * fa + corr[j][k] * fb
            return v_hi * hi + v_lo * lo
    return 0.0


def test_tilt_vs_host_double(lea):
    """testScalarFieldIceTiltZShift.py: |device - host| <= 10 cm over +-1200 m."""
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1200.0, 1200.0, size=(20000, 3)).astype(np.float32)
    got = lea[1].eval_scalar_field(0, pts)
    want = np.array([_tilt_host(lea[0].medium.tilt, *map(float, p)) for p in pts[:4000]])
    assert np.abs(got[:4000] - want).max() <= 0.1
    assert np.abs(got[:4000] - want).max() <= 2e-2   # in fact tighter than the reference's 10 cm


def test_medium_functions_vs_host_double(mie):
    sc, osc = mie
    m = sc.medium
    rng = np.random.default_rng(6)
    wl = rng.uniform(265e-9, 675e-9, 5000)
    layers = rng.integers(0, m.layersNum, 5000)
    n = osc.eval_wlen_function(0, layers, wl)
    np.testing.assert_allclose(n, [m.GetPhaseRefractiveIndex(w) for w in wl.astype(np.float32).astype(float)], rtol=3e-6)
    vg = osc.eval_wlen_function(1, layers, wl)
    np.testing.assert_allclose(vg, [0.299792458 / m.GetGroupRefractiveIndex(w) for w in wl.astype(np.float32).astype(float)], rtol=3e-6)
    sl = osc.eval_wlen_function(2, layers, wl)
    np.testing.assert_allclose(sl, [m.GetScatteringLength(int(l), w) for l, w in zip(layers, wl.astype(np.float32).astype(float))], rtol=2e-5)
    al = osc.eval_wlen_function(3, layers, wl)
    np.testing.assert_allclose(al, [m.GetAbsorptionLength(int(l), w) for l, w in zip(layers, wl.astype(np.float32).astype(float))], rtol=2e-5)
    bias = osc.eval_wlen_function(4, layers, wl)
    np.testing.assert_allclose(bias, [sc.bias.GetValue(w) for w in wl.astype(np.float32).astype(float)], rtol=2e-4)
    # physical sanity: SpiceMie scattering lengths of metres, absorption of tens to hundreds of metres at 400 nm
    s400 = osc.eval_wlen_function(2, np.arange(m.layersNum), np.full(m.layersNum, 400e-9))
    a400 = osc.eval_wlen_function(3, np.arange(m.layersNum), np.full(m.layersNum, 400e-9))
    # (layer 0 is the table's bedrock row: a_dust = 999)
    assert 0.3 < s400.min() and s400.max() < 60.0 and 5.0 < a400[1:].min() and a400.max() < 400.0 and a400[0] < 0.1


def test_scatter_direction_vs_host_double():
    rng = np.random.default_rng(7)
    n = 5000
    d = rng.normal(size=(n, 3))
    d /= np.sqrt((d ** 2).sum(1))[:, None]
    d[:10] = [0, 0, 1]
    d[10:20] = [0, 0, -1]
    cosa = rng.uniform(-1, 1, n)
    rnd = rng.uniform(0, 1, n)
    sina = np.sqrt(1 - cosa ** 2)
    got = pyoracle.scatter_direction(np.stack([cosa, sina, d[:, 0], d[:, 1], d[:, 2], rnd], axis=-1))
    # the angle between old and new direction is the requested one, and the result is a unit vector
    np.testing.assert_allclose((got * d).sum(1), cosa, atol=5e-5)  # fp32 cancellation for near-vertical directions
    np.testing.assert_allclose((got ** 2).sum(1), 1.0, atol=1e-6)
    from clsim_b200.steps import _rotate_by_angle
    want = _rotate_by_angle(d.astype(np.float32).astype(float), cosa.astype(np.float32).astype(float), rnd.astype(np.float32).astype(float))
    assert np.abs(got - want).max() < 1e-4 and np.median(np.abs(got - want)) < 2e-7


def test_scattering_angle_sampler(mie):
    """Mixed SL/HG: <cos theta> = g for both parts; fraction below/above the split."""
    sc, osc = mie
    a, x = rng_streams(1)
    c, _ = osc.sample(0, x[0], a[0], 400000)
    g, f = sc.medium.meanCosine, sc.medium.fractionOfFirstDistribution
    assert abs(c.mean() - g) < 2e-3
    assert c.min() >= -1.0 and c.max() <= 1.0
    # analytic CDFs: SL  F(c) = ((1+c)/2)^(1/beta);  HG  F(c) = (1-g^2)/(2g) * (1/sqrt(1+g^2-2gc) - 1/(1+g))
    beta = (1 - g) / (1 + g)
    def cdf(v):
        v = np.asarray(v, dtype=float)
        sl = ((1 + v) / 2) ** (1 / beta)
        hg = (1 - g * g) / (2 * g) * (1 / np.sqrt(1 + g * g - 2 * g * v) - 1 / (1 + g))
        return f * sl + (1 - f) * hg
    assert sps.kstest(c[:100000].astype(float), cdf).pvalue > 1e-3


def test_wavelength_sampler(mie):
    sc, osc = mie
    a, x = rng_streams(2)
    w, _ = osc.sample(1, x[1], a[1], 300000)
    assert w.min() >= 260e-9 and w.max() <= 680e-9
    gen = sc.generators[0]
    grid = gen.x0 + gen.dx * np.arange(len(gen.y))
    cum = np.concatenate([[0.0], np.cumsum(0.5 * (gen.y[1:] + gen.y[:-1]) * gen.dx)])
    cum /= cum[-1]
    def cdf(v):
        # piecewise-linear density -> piecewise-quadratic CDF
        v = np.asarray(v, dtype=float)
        k = np.clip(((v - gen.x0) / gen.dx).astype(int), 0, len(gen.y) - 2)
        t = v - grid[k]
        slope = (gen.y[k + 1] - gen.y[k]) / gen.dx
        return cum[k] + (gen.y[k] * t + 0.5 * slope * t * t) / (0.5 * (gen.y[1:] + gen.y[:-1]) * gen.dx).sum()
    assert sps.kstest(w[:100000].astype(float), cdf).pvalue > 1e-3
    # the no-dispersion generator: 1/lambda uniform
    flat = ice.WlenBias(constant=1.0)
    g2 = ice.makeCherenkovWavelengthGenerator(flat, True, sc.medium)
    osc2 = pyoracle.Scene(sc.medium, sc.geo, [g2], flat, sc.options())
    w2, _ = osc2.sample(1, x[0], a[0], 100000)
    inv = 1.0 / w2.astype(float)
    lo, hi = 1 / 675e-9, 1 / 265e-9
    assert sps.kstest((inv - lo) / (hi - lo), "uniform").pvalue > 1e-3


def test_homogeneous_medium_random_walk_statistics():
    """Physics sanity of the whole loop without detector effects (save-all, one layer):
    absorbed after exp(1)-distributed absorption lengths; segments/photon = 1 + lambda_a/lambda_s;
    nothing else is saved."""
    sc = make_scene("homogeneous")
    from clsim_b200 import steps
    opt = sc.options(stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=1.0)
    osc = pyoracle.Scene(sc.medium, None, sc.generators, sc.bias, opt)
    bunch = steps.point_source_steps(200, 100, seed=3)
    a, x = rng_streams(len(bunch))
    ph, cnt, st, _, _ = osc.propagate(bunch, x, a, cap=200 * 100, num_threads=os.cpu_count() or 1)
    assert cnt == 200 * 100 == st["photons"]
    assert sps.kstest(ph["dist_in_abs_lens"].astype(float), "expon").pvalue > 1e-3
    m = sc.medium
    wl = ph["wavelength"].astype(float)
    ratio = np.array([m.GetAbsorptionLength(0, w) / m.GetScatteringLength(0, w) for w in wl])
    # E[scatters | wlen] = lambda_a/lambda_s
    assert abs(ph["num_scatters"].mean() / ratio.mean() - 1.0) < 0.03
    assert st["segments"] == int(ph["num_scatters"].sum()) + st["photons"]
    # time = start time + path / v_g
    np.testing.assert_allclose(ph["t"] - ph["start_t"], ph["cherenkov_dist"] / ph["group_velocity"], rtol=2e-4)
    # draw accounting (quirk 1): 4 per photon + 1 per segment + 2 per scatter + 1 prescale draw per photon
    assert st["draws"] == 4 * st["photons"] + st["segments"] + 2 * int(ph["num_scatters"].sum()) + st["photons"]


def test_dom_hits_are_on_the_pancaked_surface(mie):
    sc, osc = mie
    from clsim_b200 import steps
    bunch = steps.muon_track_steps(1500, seed=4)
    a, x = rng_streams(len(bunch))
    ph, cnt, st, _, _ = osc.propagate(bunch, x, a, num_threads=os.cpu_count() or 1)
    assert len(ph) > 100
    # undoing the pancake maps every hit onto the true DOM sphere of radius 0.1651 m
    r = np.sqrt(ph["x"].astype(np.float64) ** 2 + ph["y"].astype(np.float64) ** 2 + ph["z"].astype(np.float64) ** 2)
    assert np.all(np.abs(r - 0.16510) < 1e-3)
    # ids are real IDs of the geometry
    ids = set(zip(sc.geo.stringIDs.tolist(), sc.geo.domIDs.tolist()))
    assert all((int(s), int(d)) in ids for s, d in zip(ph["string_id"], ph["om_id"]))
    assert np.all(ph["weight"] > 0)


@pytest.mark.skipif(not pyoracle.ref_rng_available(), reason="oracle/_ref not built (no /root/reference at build time)")
def test_seeding_against_the_references_init_mwc_rng(tmp_path):
    """R2, the seeding half: init_MWC_RNG (private/opencl/mwcrng_init.h, compiled unmodified into oracle/_ref/libclsim_ref_rng.so)
    reads the multipliers from a table in the reference's text format and draws the start states; with a random service that
    hands out splitmix64's values 32 bits at a time it makes the multipliers and the states the product (clsimcu_seed_rng_states,
    clsimcu_safeprime_multipliers) and the oracle make.  Multipliers chosen so that the rejection rule fires: the table's own,
    and small ones for which half of the candidates are refused."""
    from clsim_b200 import capi

    def splitmix64(seed, n):
        out, s, m = [], seed, (1 << 64) - 1
        for _ in range(n):
            s = (s + 0x9E3779B97F4A7C15) & m
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
            out.append(z ^ (z >> 31))
        return np.array(out, dtype=np.uint64)

    n = 3000
    table = capi.safeprime_multipliers(0, n)
    # the reference's text format (resources/.../rnd.txt): the multiplier first, the rest of the line is skipped
    path = tmp_path / "safeprimes_base32.txt"
    path.write_text("".join("%d %d %d\n" % (a, a * (1 << 32) - 1, a * (1 << 31) - 1) for a in map(int, table)))
    for seed in (0, 5, 123456789):
        x, a, used = pyoracle.ref_init_mwc_rng(n, str(path), splitmix64(seed, n + 64))
        assert np.array_equal(a, table)
        assert np.array_equal(x, capi.seed_rng_states(seed, table))          # the product's rule
        assert np.array_equal(x, pyoracle.seed_states(seed, table))          # the oracle's
        assert used >= n
    # multipliers near 2^31: every second candidate has an upper half >= a - 1 and is refused
    small = (np.arange(500, dtype=np.uint64) * 2 + (1 << 31)).astype(np.uint32)
    path.write_text("".join("%d\n" % a for a in small))
    x, a, used = pyoracle.ref_init_mwc_rng(len(small), str(path), splitmix64(77, 4 * len(small)))
    assert np.array_equal(a, small) and used > 1.7 * len(small)
    assert np.array_equal(x, capi.seed_rng_states(77, small)) and np.array_equal(x, pyoracle.seed_states(77, small))
    # the reference's own shipped table: same first rows as the product generates
    ref_table = "/root/reference/resources/scripts/compareToPPCredux/test_ice_models/lea/rnd.txt"
    import os
    if os.path.isfile(ref_table):
        x, a, _ = pyoracle.ref_init_mwc_rng(2000, ref_table, splitmix64(1, 2100))
        assert np.array_equal(a, capi.safeprime_multipliers(0, 2000)) and np.array_equal(x, capi.seed_rng_states(1, a))
    # a value that does not fit 32 bits is refused by the reference's range check (:97-100)
    path.write_text("4294967296\n")
    with pytest.raises(RuntimeError):
        pyoracle.ref_init_mwc_rng(1, str(path), splitmix64(1, 8))
