"""Parity proper, part 2: the fast persistent kernel (the product path).

(a) per photon: every photon recorded in save-all mode carries the RNG states it was created
    and propagated from; the CPU oracle replays it and must land on the same record within the
    stated fp32 tolerance (approximate MUFU math on the GPU vs libm: 1 cm on the end point over
    paths of O(100 m), 1e-3 relative on path length / absorption lengths, equal scatter count);
(b) statistically: >= 1e8 photons through the fast kernel and through the reference-order
    kernel (itself tied to the oracle photon by photon in test_gpu_reference_kernel.py):
    hit fraction, per-string counts (chi2), arrival time / zenith / wavelength / scatter-count
    distributions (KS) with p > 0.01 / (number of tests) each;
(c) size-independent properties at full bench size: photon conservation, dummy steps, ragged
    photon counts, hits only on existing DOM IDs.
All calls go through the C ABI."""
import os

import numpy as np
import pytest
from scipy import stats as sps

from clsim_b200 import capi, steps
from clsim_b200.description import KERNEL_FAST, KERNEL_REFERENCE, STEP_DTYPE
from oracle import pyoracle
from tests.scenes import add_flasher_generator, dom_near, make_scene

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


@pytest.mark.parametrize("name", ["spice_mie", "spice_lea", "homogeneous"])
def test_replay_parity_per_photon(name):
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
    assert res["photons"] == 256 * 40 == len(photons) == res["hits"]
    osc = pyoracle.Scene(sc.medium, None, sc.generators, sc.bias, opt)
    good = 0
    worst = 0.0
    for i, p in enumerate(photons):
        s = int(p["identifier"])  # identifier == step index in this test
        saved, q = osc.single_photon_split(bunch[s], tags_x[i, 0], tags_a[i, 0], tags_x[i, 1], tags_a[i, 1])
        assert saved
        # creation is reproduced (same stream, same draws): start point and wavelength agree tightly
        assert abs(q["start_x"] - p["start_x"]) < 1e-3 and abs(q["start_z"] - p["start_z"]) < 1e-3
        assert abs(q["wavelength"] - p["wavelength"]) < 1e-6 * q["wavelength"]
        dev = max(abs(float(q[k]) - float(p[k])) for k in ("x", "y", "z"))
        same = (q["num_scatters"] == p["num_scatters"] and dev < 1e-2
                and abs(q["cherenkov_dist"] - p["cherenkov_dist"]) <= 1e-3 * max(1.0, q["cherenkov_dist"])
                and abs(q["dist_in_abs_lens"] - p["dist_in_abs_lens"]) <= 1e-3 * max(1.0, q["dist_in_abs_lens"])
                and abs(q["t"] - p["t"]) <= 0.05 + 1e-4 * abs(q["t"]))
        if same:
            good += 1
            worst = max(worst, dev)
    # photons whose fate flips on a rounding difference (absorbed one segment earlier/later) are allowed at the 1 % level
    assert good >= 0.99 * len(photons), (good, len(photons))
    assert worst < 1e-2


def _run_resident(sc, bunch, mode, seed, repeat=1, per_item=2, **more_options):
    opt = sc.options(kernel_mode=mode, max_num_workitems=len(bunch), rng_seed=seed, output_photons_per_workitem=per_item, **more_options)
    hits = []
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        eng.upload_resident(bunch)
        tot = {"photons": 0, "hits": 0, "segments": 0}
        for _ in range(repeat):
            r = eng.run_resident(1)
            for k in tot:
                tot[k] += r[k]
            hits.append(eng.download_resident())
    return np.concatenate(hits), tot


def _compare_distributions(a, b, tot_a, tot_b, n_tests_extra=0):
    """a, b: hit lists; returns dict of p-values."""
    p = {}
    # hit fraction: two-sample binomial z test
    fa, fb = len(a) / tot_a["photons"], len(b) / tot_b["photons"]
    pool = (len(a) + len(b)) / float(tot_a["photons"] + tot_b["photons"])
    z = (fa - fb) / np.sqrt(pool * (1 - pool) * (1.0 / tot_a["photons"] + 1.0 / tot_b["photons"]))
    p["hit_fraction"] = 2 * sps.norm.sf(abs(z))
    # segments per photon (mean of a sum of ~30 per photon: compare within 0.2 %)
    p["_seg_ratio"] = (tot_a["segments"] / tot_a["photons"]) / (tot_b["segments"] / tot_b["photons"])
    # per-string counts: chi2 contingency over strings with enough entries
    sa = np.bincount(a["string_id"].astype(int), minlength=100)
    sb = np.bincount(b["string_id"].astype(int), minlength=100)
    keep = (sa + sb) >= 20
    p["per_string_chi2"] = sps.chi2_contingency(np.stack([sa[keep], sb[keep]]))[1]
    # per-DOM-layer (om id) counts
    da = np.bincount(a["om_id"].astype(int), minlength=70)
    db = np.bincount(b["om_id"].astype(int), minlength=70)
    keep = (da + db) >= 20
    p["per_om_chi2"] = sps.chi2_contingency(np.stack([da[keep], db[keep]]))[1]
    # per-DOM counts: chi2 contingency over every (string, OM) cell with >= 20 hits in the two arms together
    cell_a = a["string_id"].astype(np.int64) * 100 + a["om_id"].astype(np.int64)
    cell_b = b["string_id"].astype(np.int64) * 100 + b["om_id"].astype(np.int64)
    ca = np.bincount(cell_a, minlength=10000)
    cb = np.bincount(cell_b, minlength=10000)
    keep = (ca + cb) >= 20
    if keep.sum() >= 2:
        p["per_dom_chi2"] = sps.chi2_contingency(np.stack([ca[keep], cb[keep]]))[1]
        p["_per_dom_cells"] = int(keep.sum())
    # per-DOM arrival times: KS on each of the 50 busiest DOMs, the smallest p-value Bonferroni-corrected for the 50
    busiest = [c for c in np.argsort(-(np.minimum(ca, cb)))[:50] if min(ca[c], cb[c]) >= 30]
    if busiest:
        ta, tb = a["t"] - a["start_t"], b["t"] - b["start_t"]
        pv = [sps.ks_2samp(ta[cell_a == c], tb[cell_b == c]).pvalue for c in busiest]
        p["per_dom_time_ks"] = min(1.0, min(pv) * len(pv))
    p["arrival_time_ks"] = sps.ks_2samp(a["t"] - a["start_t"], b["t"] - b["start_t"]).pvalue
    p["zenith_ks"] = sps.ks_2samp(a["theta"], b["theta"]).pvalue
    p["wavelength_ks"] = sps.ks_2samp(a["wavelength"], b["wavelength"]).pvalue
    p["num_scatters_ks"] = sps.ks_2samp(a["num_scatters"] + np.random.default_rng(0).uniform(0, 1, len(a)),
                                        b["num_scatters"] + np.random.default_rng(1).uniform(0, 1, len(b))).pvalue
    p["abs_lens_ks"] = sps.ks_2samp(a["dist_in_abs_lens"], b["dist_in_abs_lens"]).pvalue
    # impact angle on the DOM: cosine between the photon direction and the outward normal at the hit point.
    # (The impact RADIUS carries no information: un-pancaking maps every hit onto the true DOM sphere.)
    def cos_eta(h):
        d = np.stack([np.sin(h["theta"]) * np.cos(h["phi"]), np.sin(h["theta"]) * np.sin(h["phi"]), np.cos(h["theta"])], axis=-1)
        r = np.stack([h["x"], h["y"], h["z"]], axis=-1).astype(np.float64)
        return (d * r).sum(1) / np.sqrt((r ** 2).sum(1))
    p["impact_angle_ks"] = sps.ks_2samp(cos_eta(a), cos_eta(b)).pvalue
    p["azimuth_ks"] = sps.ks_2samp(a["phi"], b["phi"]).pvalue
    return p


def _assert_same_distributions(attempt, seg_tol):
    """Two-stage test.  `attempt(k)` runs both arms with the k-th set of seeds and returns the p-values.  Each of
    the m statistics is tested at 0.01/m (Bonferroni: p > 0.01 for the family, the bar BASELINE.json states); a
    statistic that falls below it is tested once more on independent samples and must pass then.  A true
    difference fails both times; chance alone fails the family about once in 10^4 runs instead of once in 10^2,
    which matters for a suite that is run at every round."""
    p = attempt(0)
    print({k: float("%.3g" % v) for k, v in p.items()})
    assert abs(p.pop("_seg_ratio") - 1.0) < seg_tol
    p.pop("_per_dom_cells", None)
    m = len(p)
    low = [k for k, v in p.items() if not v > 0.01 / m]
    if low:
        p2 = attempt(1)
        print("second sample for", low, {k: float("%.3g" % v) for k, v in p2.items()})
        assert abs(p2.pop("_seg_ratio") - 1.0) < seg_tol
        p2.pop("_per_dom_cells", None)
        for k in low:
            assert p2[k] > 0.01 / m, (k, p[k], p2[k])


@pytest.mark.parametrize("name,n_steps", [("spice_mie", 1 << 19), ("spice_lea", 1 << 19)])
def test_statistical_parity_1e8_photons(name, n_steps):
    """>= 1e8 photons per arm (2^19 steps x 200)."""
    sc = make_scene(name)
    bunch = steps.muon_track_steps(n_steps, seed=31) if name == "spice_mie" else steps.muon_bundle_steps(n_steps, num_muons=50, seed=32)
    kept = {}

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=101 + 1000 * k)
        ref, tot_r = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=202 + 1000 * k)
        assert tot_f["photons"] == tot_r["photons"] == int(bunch["num_photons"].sum()) >= 1e8
        assert len(fast) > 5e4 and len(ref) > 5e4
        kept["fast"], kept["ref"] = fast, ref
        return _compare_distributions(fast, ref, tot_f, tot_r)

    _assert_same_distributions(attempt, 2e-3)
    fast, ref = kept["fast"], kept["ref"]
    # geometric facts: after the pancake is undone every hit sits on the true DOM sphere
    # (propagation_kernel.c.cl:340-355), IDs exist
    for h in (fast, ref):
        r = np.sqrt(h["x"].astype(np.float64) ** 2 + h["y"].astype(np.float64) ** 2 + h["z"].astype(np.float64) ** 2)
        assert np.all(np.abs(r - 0.16510) < 1e-3)
    assert fast["string_id"].min() >= 1 and fast["string_id"].max() <= 86
    assert fast["om_id"].min() >= 1 and fast["om_id"].max() <= 60


@pytest.mark.parametrize("name,n_steps", [("spice_mie", 1 << 19), ("spice_lea", 1 << 19)])
def test_north_star_a_fast_kernel_against_the_reference_1e8_photons(name, n_steps):
    """North-star (a), directly: per-DOM hit counts and arrival-time and angular distributions of the FAST kernel
    against the reference on the same inputs, chi2 / KS p > 0.01 (family-wise), >= 1e8 photons per arm, on BASELINE
    config 2 (SpiceMie, muon track) and config 3 (SpiceLea + tilt + anisotropy, muon bundle).  The CPU arm is the
    reference's OWN kernel text compiled for the host (oracle/_ref) where that library exists, else the oracle
    restatement that is bit-identical to it (tests/test_ref_kernel.py); no CUDA kernel stands in between."""
    sc = make_scene(name)
    bunch = steps.muon_track_steps(n_steps, seed=131) if name == "spice_mie" else steps.muon_bundle_steps(n_steps, num_muons=50, seed=132)
    opt = sc.options(max_num_workitems=len(bunch))
    a = capi.safeprime_multipliers(0, len(bunch))
    if pyoracle.ref_available():
        cpu = pyoracle.RefScene(sc.medium, sc.geo, sc.generators, sc.bias, opt)
        assert cpu.variant() is not None
    else:
        cpu = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, opt)
    total = int(bunch["num_photons"].sum())
    assert total >= 1e8

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=301 + 1000 * k)
        x = pyoracle.seed_states(9000 + k, a)
        out = cpu.propagate(bunch, x, a, cap=len(bunch), num_threads=THREADS)
        want, counted = out[0], out[1]
        assert counted == len(want) > 5e4 and tot_f["hits"] == len(fast)
        # (the reference's kernel text does not count segments: the segment ratio is taken from the restatement's
        # counter on a sample of the same bunch)
        osc = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, opt)
        sub = slice(0, None, 32)   # every 32nd step: all muons of a bundle take part
        _, _, st, _, _ = osc.propagate(bunch[sub], x[sub], a[sub], num_threads=THREADS)
        tot_c = {"photons": total, "hits": counted, "segments": st["segments"] * (total / float(st["photons"]))}
        return _compare_distributions(fast, want, tot_f, tot_c)

    _assert_same_distributions(attempt, 5e-3)


def test_fast_kernel_against_oracle_small_sample():
    """Direct fast-kernel vs CPU-oracle comparison at the size the oracle finishes in seconds."""
    sc = make_scene("homogeneous")
    src = dom_near(sc.geo, (0.0, 0.0, 0.0)) + np.array([10.0, 5.0, 3.0])
    bunch = steps.point_source_steps(5000, 200, pos=tuple(src), seed=41)  # config 1: 1e6 photons
    a, _, _ = pyoracle.safeprimes(0, len(bunch))
    osc = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, sc.options())

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=7 + 1000 * k, repeat=4)
        wants, tot_o = [], {"photons": 0, "hits": 0, "segments": 0}
        for rep in range(2):
            x = pyoracle.seed_states(900 + rep + 1000 * k, a)
            w, cnt, st, _, _ = osc.propagate(bunch, x, a, num_threads=THREADS)
            wants.append(w)
            tot_o["photons"] += st["photons"]
            tot_o["segments"] += st["segments"]
        want = np.concatenate(wants)
        assert len(want) > 3000
        return _compare_distributions(fast, want, tot_f, tot_o)

    _assert_same_distributions(attempt, 1e-2)


def test_dense_strings_take_the_cell_walk(monkeypatch):
    """Strings closer together than the pixels of the collision map are wide: the map has no range there, every
    leg is parked and takes the reference's cell walk in the slow phase (sparse_collision_kernel.c.cl:194-460).
    A cluster of three strings 6-7 m apart in its own subdetector next to the 24-DOM ring, a map forced coarse
    (CLSIMCU_PIXEL_BUDGET), a point source inside the cluster: the fast kernel must see what the reference-order
    kernel sees."""
    from clsim_b200 import geometry
    from clsim_b200.description import SimpleGeometry
    sc = make_scene("homogeneous", geo_kind="ring")
    ring = sc.geo
    sid, did, xs, ys, zs = list(ring.stringIDs), list(ring.domIDs), list(ring.posX), list(ring.posY), list(ring.posZ)
    sub = ["Ring"] * len(sid)
    for k, (cx, cy) in enumerate(((0.0, 0.0), (6.0, 0.5), (-0.5, 7.0))):
        for d in range(6):
            sid.append(20 + k); did.append(d + 1); xs.append(cx + 0.1 * d); ys.append(cy - 0.05 * d); zs.append(25.0 - 10.0 * d)
            sub.append("Cluster")
    sc.geo = SimpleGeometry(sid, did, xs, ys, zs, ring.OMRadius, subdetectors=sub)
    monkeypatch.setenv("CLSIMCU_PIXEL_BUDGET", "512")   # pixels of ~12 m over the 240 m ring
    bunch = steps.point_source_steps(1 << 14, 200, pos=(3.0, 3.0, 2.0), seed=77)

    def attempt(k):
        # (room for every photon: a full output buffer keeps the first hits written, a biased sample)
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=11 + 1000 * k, per_item=200)
        ref, tot_r = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=12 + 1000 * k, per_item=200)
        assert tot_f["photons"] == tot_r["photons"] == int(bunch["num_photons"].sum())
        assert tot_f["hits"] == len(fast) and tot_r["hits"] == len(ref)
        assert len(ref) > 2e4 and set(np.unique(ref["string_id"])) >= {20, 21, 22}
        return _compare_distributions(fast, ref, tot_f, tot_r)

    _assert_same_distributions(attempt, 1e-2)


@pytest.mark.parametrize("scat_kind", [1, 2])
def test_single_scattering_angle_distributions(scat_kind):
    """The ice models mix two scattering-angle samplers (R9), and that mix is compiled into the fast kernel; a
    medium with Henyey-Greenstein (1) or simplified-Liu (2) alone takes the kernel's generic sampler code."""
    sc = make_scene("homogeneous")
    sc.medium.scat_kind = scat_kind
    src = dom_near(sc.geo, (0.0, 0.0, 0.0)) + np.array([10.0, 5.0, 3.0])
    bunch = steps.point_source_steps(1 << 16, 200, pos=tuple(src), seed=61)

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=31 + 1000 * k)
        ref, tot_r = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=32 + 1000 * k)
        assert tot_f["photons"] == tot_r["photons"] == int(bunch["num_photons"].sum())
        assert tot_f["hits"] == len(fast) and tot_r["hits"] == len(ref) and len(ref) > 1e4
        return _compare_distributions(fast, ref, tot_f, tot_r)

    _assert_same_distributions(attempt, 2e-3)


@pytest.mark.parametrize("kind", ["unequal", "no_dispersion", "constant"])
def test_other_wavelength_generators(kind):
    """The wavelength generators besides the equally spaced table of the default setup (R3a): a table on unequal
    abscissae (the staged path of the fast kernel with its abscissae in global memory), the analytic 1/lambda^2
    spectrum (I3CLSimRandomValueWlenCherenkovNoDispersion) and a single wavelength (I3CLSimRandomValueConstant)."""
    from clsim_b200.description import WlenGenerator
    sc = make_scene("spice_mie")
    g0 = sc.generators[0]
    if kind == "unequal":
        x = g0.x0 + g0.dx * np.arange(len(g0.y))
        keep = np.ones(len(x), dtype=bool)
        keep[3::4] = False                      # drop every fourth node: unequal spacing, a different spectrum for both kernels
        keep[0] = keep[-1] = True
        sc.generators = [WlenGenerator.interpolated_unequal(x[keep], g0.y[keep])]
    elif kind == "no_dispersion":
        sc.generators = [WlenGenerator.cherenkov_no_dispersion(sc.medium.GetMinWavelength(), sc.medium.GetMaxWavelength())]
    else:
        sc.generators = [WlenGenerator.constant(405e-9)]
    bunch = steps.muon_track_steps(1 << 16, seed=71)

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=41 + 1000 * k)
        ref, tot_r = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=42 + 1000 * k)
        assert tot_f["photons"] == tot_r["photons"] == int(bunch["num_photons"].sum())
        assert tot_f["hits"] == len(fast) and tot_r["hits"] == len(ref) and len(ref) > 5e3
        if kind == "constant":
            assert np.all(fast["wavelength"] == np.float32(405e-9)) and np.all(ref["wavelength"] == np.float32(405e-9))
        return _compare_distributions(fast, ref, tot_f, tot_r)

    _assert_same_distributions(attempt, 2e-3)


def test_fixed_number_of_absorption_lengths():
    """FixedNumberOfAbsorptionLengths (propagation_kernel.c.cl:582-585): every photon lives exactly that many
    absorption lengths unless a DOM stops it.  Fast kernel against the reference-order kernel on SpiceLea with tilt
    and anisotropy (the budget is rescaled at every scatter there), same statistics as the other tests."""
    sc = make_scene("spice_lea")
    bunch = steps.muon_track_steps(1 << 17, seed=55)

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=21 + 1000 * k, fixed_number_of_absorption_lengths=3.0)
        ref, tot_r = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=22 + 1000 * k, fixed_number_of_absorption_lengths=3.0)
        assert tot_f["photons"] == tot_r["photons"] == int(bunch["num_photons"].sum())
        assert tot_f["hits"] == len(fast) and tot_r["hits"] == len(ref) and len(ref) > 2e4
        for h in (fast, ref):
            assert h["dist_in_abs_lens"].max() <= 3.0 + 1e-3
        return _compare_distributions(fast, ref, tot_f, tot_r)

    _assert_same_distributions(attempt, 2e-3)


def test_flasher_mode_statistics():
    sc = add_flasher_generator(make_scene("spice_lea", oversize=1.0))
    dom = dom_near(sc.geo, (0.0, 0.0, -200.0))
    bunch = steps.flasher_steps(1 << 17, dom, seed=51)
    kept = {}

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=11 + 1000 * k)
        ref, tot_r = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=12 + 1000 * k)
        assert len(ref) > 2000
        kept["fast"] = fast
        return _compare_distributions(fast, ref, tot_f, tot_r)

    _assert_same_distributions(attempt, 5e-3)
    fast = kept["fast"]
    # oversize 1: hits on the true DOM surface
    # (the entry point comes from along - sqrt(along^2 - |r|^2 + R^2) in fp32, like the reference's: on a grazing hit at the
    # end of a 20 m leg the cancellation leaves a few mm, so the bound is on the bulk and, loosely, on the worst hit)
    r = np.sqrt(fast["x"] ** 2 + fast["y"] ** 2 + fast["z"] ** 2)
    off = np.abs(r - 0.16510)
    assert np.mean(off < 1e-3) > 0.999 and off.max() < 2e-2


def test_conservation_and_ragged_inputs():
    """Every photon of every step is created exactly once: ragged photon counts, dummy steps,
    a single huge step, bunch sizes that are not multiples of the warp size."""
    sc = make_scene("spice_mie", geo_kind="ring")
    rng = np.random.default_rng(61)
    bunch = steps.muon_track_steps(1237, seed=62)
    bunch["num_photons"] = rng.integers(0, 400, len(bunch))
    bunch["num_photons"][::7] = 0          # dummy steps
    bunch["num_photons"][5] = 50000        # one very long step
    opt = sc.options(kernel_mode=KERNEL_FAST, max_num_workitems=2048, rng_seed=3)
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        eng.upload_resident(bunch)
        r = eng.run_resident(3)
        assert r["photons"] == 3 * int(bunch["num_photons"].sum())
        assert r["segments"] > r["photons"]
        one = steps.muon_track_steps(1, seed=63)
        eng.upload_resident(one)
        r1 = eng.run_resident(1)
        assert r1["photons"] == 200
        allzero = bunch.copy()
        allzero["num_photons"] = 0
        eng.upload_resident(allzero)
        r0 = eng.run_resident(1)
        assert r0["photons"] == 0 and r0["hits"] == 0 and r0["segments"] == 0
    # the same counters, with the statistics build of the kernel checked against the step list
    opt = sc.options(kernel_mode=KERNEL_FAST, stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=1.0,
                     max_num_workitems=2048, output_photons_per_workitem=200, rng_seed=4)
    with capi.Engine(sc.medium, None, sc.generators, sc.bias, opt) as eng:
        eng.upload_resident(bunch)
        r = eng.run_resident(1)
        saved = eng.download_resident()
    assert len(saved) == int(bunch["num_photons"].sum()) == r["hits"]
    # every step contributed exactly num_photons records (steps are distinguishable by identifier here)
    bunch2 = bunch.copy()
    bunch2["identifier"] = np.arange(len(bunch2))
    with capi.Engine(sc.medium, None, sc.generators, sc.bias, opt) as eng:
        eng.upload_resident(bunch2)
        eng.run_resident(1)
        saved = eng.download_resident()
    counts = np.bincount(saved["identifier"].astype(int), minlength=len(bunch2))
    assert np.array_equal(counts, bunch2["num_photons"])


def test_non_stop_detection_on_the_fast_kernel():
    """StopDetectedPhotons = false (sparse_collision_kernel.c.cl:166-187): every DOM a leg crosses is recorded and the
    photon flies on.  The fast kernel against the reference-order kernel (statistics) and against the oracle directly
    (counts); more hits than with stopping, some photons seen by more than one DOM."""
    sc = make_scene("spice_mie")
    bunch = steps.muon_track_steps(1 << 17, seed=81)
    kept = {}

    def attempt(k):
        fast, tot_f = _run_resident(sc, bunch, KERNEL_FAST, seed=51 + 1000 * k, per_item=4, stop_detected_photons=False)
        ref, tot_r = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=52 + 1000 * k, per_item=4, stop_detected_photons=False)
        assert tot_f["photons"] == tot_r["photons"] == int(bunch["num_photons"].sum())
        assert tot_f["hits"] == len(fast) and tot_r["hits"] == len(ref) and len(ref) > 2e4
        kept["fast"], kept["tot"] = fast, tot_f
        return _compare_distributions(fast, ref, tot_f, tot_r)

    _assert_same_distributions(attempt, 2e-3)
    stopped, tot_s = _run_resident(sc, bunch, KERNEL_FAST, seed=53)
    # (a photon that flies on can be seen again: about 2 % more hits than with stopping, 28 480 against 27 996 when measured)
    assert len(kept["fast"]) > 1.005 * len(stopped) * (kept["tot"]["photons"] / float(tot_s["photons"]))
    # the oracle (intent of the reference's de-duplication, see tests/test_ref_kernel.py) on a sample
    small = bunch[:1 << 13]
    a = capi.safeprime_multipliers(0, len(small))
    osc = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, sc.options(stop_detected_photons=False))
    want, counted, st, _, _ = osc.propagate(small, pyoracle.seed_states(5, a), a, cap=4 * len(small), num_threads=THREADS)
    got, tot = _run_resident(sc, small, KERNEL_FAST, seed=54, per_item=4, stop_detected_photons=False, repeat=4)
    z = (len(got) / 4.0 - counted) / np.sqrt(counted * (1 + 0.25))
    assert abs(z) < 4.5, (len(got) / 4.0, counted)


def test_photon_history_on_the_fast_kernel():
    """PhotonHistoryEntries (propagation_kernel.c.cl:833-837): the last N scatter points of every recorded photon.
    Save-all mode, every photon recorded with the RNG states it was made from: the oracle replays each photon with its
    trajectory, and the fast kernel's history must be the oracle's scatter points (1 cm, like the end point; absorption
    lengths used: 1e-3) -- for photons with at most N scatters the whole path, else the last N points."""
    n_hist = 6
    sc = make_scene("spice_lea")
    bunch = steps.muon_track_steps(128, photons_per_step=40, seed=91)
    bunch["identifier"] = np.arange(len(bunch))
    opt = sc.options(kernel_mode=KERNEL_FAST, stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=1.0,
                     max_num_workitems=len(bunch), output_photons_per_workitem=40, rng_seed=6, photon_history_entries=n_hist)
    with capi.Engine(sc.medium, None, sc.generators, sc.bias, opt) as eng:
        eng.upload_resident(bunch)
        res = eng.run_resident(1)
        photons = eng.download_resident()
        history = eng.download_resident_history(len(photons), n_hist)
        tags_x, tags_a = eng.download_resident_rng_tags(len(photons))
    assert res["photons"] == 128 * 40 == len(photons)
    osc = pyoracle.Scene(sc.medium, None, sc.generators, sc.bias, opt)
    good = checked = 0
    for i, p in enumerate(photons):
        saved, q, traj = osc.single_photon_split(bunch[int(p["identifier"])], tags_x[i, 0], tags_a[i, 0], tags_x[i, 1], tags_a[i, 1], max_points=400)
        s = int(p["num_scatters"])
        if not saved or int(q["num_scatters"]) != s or len(traj) < s + 2:
            continue   # fate flipped on a rounding difference (1 % level, see test_replay_parity_per_photon)
        checked += 1
        rec = min(s, n_hist)
        assert np.all(np.isnan(history[i, rec:])) and not np.any(np.isnan(history[i, :rec]))
        if rec == 0:
            good += 1
            continue
        # trajectory row k (k >= 1) is the state after segment k: the k-th scatter point for k <= s
        pts = traj[1 + s - rec:1 + s]
        used = traj[0, 7] - pts[:, 7]
        ok = (np.abs(history[i, :rec, :3] - pts[:, :3]).max() < 1e-2) and np.all(np.abs(history[i, :rec, 3] - used) <= 1e-3 * np.maximum(1.0, used))
        good += bool(ok)
    assert checked >= 0.98 * len(photons) and good >= 0.99 * checked, (good, checked, len(photons))


def test_photon_history_of_detected_photons_through_the_converter():
    """History through EnqueueSteps / GetConversionResult with stopping detection on the fast kernel: structure and
    geometry of what comes back (the last scatter point of a hit lies one straight flight away from the hit DOM)."""
    from clsim_b200.converter import initializeCUDA
    n_hist = 5
    sc = make_scene("spice_mie")
    conv = initializeCUDA({"ordinal": 0, "approximateNumberOfWorkItems": 1 << 16}, 9, sc.geo, sc.medium, sc.bias, sc.generators,
                          stopDetectedPhotons=True, pancakeFactor=5.0, photonHistoryEntries=n_hist, kernelMode=KERNEL_FAST)
    bunch = steps.muon_track_steps(1 << 16, seed=92)
    conv.EnqueueSteps(bunch, 3)
    res = conv.GetConversionResult()
    conv.Close()
    ph, hist = res.photons, res.photonHistories
    assert len(ph) > 5000 and hist.shape == (len(ph), n_hist, 4)
    s = ph["num_scatters"].astype(int)
    rec = np.minimum(s, n_hist)
    for j in range(n_hist):
        assert np.all(np.isnan(hist[rec <= j, j])) and not np.any(np.isnan(hist[rec > j, j]))
    pos = {(int(a), int(b)): (x, y, z) for a, b, x, y, z in zip(sc.geo.stringIDs, sc.geo.domIDs, sc.geo.posX, sc.geo.posY, sc.geo.posZ)}
    some = np.flatnonzero(rec > 0)
    dom = np.array([pos[(int(ph["string_id"][i]), int(ph["om_id"][i]))] for i in some])
    last = hist[some, rec[some] - 1, :3]
    to_dom = np.sqrt(((last - dom) ** 2).sum(1))
    # the flight from the last scatter point to the hit is what is left of the path; whole paths for photons with few scatters
    few = some[s[some] <= n_hist]
    start = np.stack([ph["start_x"][few], ph["start_y"][few], ph["start_z"][few]], axis=-1)
    chain = np.concatenate([start[:, None, :], hist[few, :, :3]], axis=1)
    flown = np.array([np.sqrt((np.diff(chain[i, :rec[f] + 1], axis=0) ** 2).sum(1)).sum() for i, f in enumerate(few)])
    left = ph["cherenkov_dist"][few] - flown
    d_few = to_dom[np.isin(some, few)]
    assert np.all(left > -1e-2) and np.all(np.abs(left - d_few) < 0.8255 + 0.05)
    # absorption lengths used: increasing along the history, not beyond the photon's total
    w = hist[some, :, 3]
    assert np.all(np.nan_to_num(np.diff(w, axis=1), nan=1.0) > -1e-4)
    assert np.all(np.nanmax(w, axis=1) <= ph["dist_in_abs_lens"][some] + 1e-3)


def test_generic_direction_transforms_agree_with_the_block_form(monkeypatch):
    """The ppc anisotropy matrices take the five-product block form (kVarBlockTransforms); the nine-product form that any
    other matrix takes must describe the same physics: the usual family of statistics, block form against generic form
    (CLSIMCU_GENERIC_TRANSFORMS is the kernel launcher's test hook)."""
    sc = make_scene("spice_lea")
    bunch = steps.muon_track_steps(1 << 17, seed=71)

    def attempt(k):
        monkeypatch.delenv("CLSIMCU_GENERIC_TRANSFORMS", raising=False)
        block, tot_b = _run_resident(sc, bunch, KERNEL_FAST, seed=77 + 1000 * k)
        monkeypatch.setenv("CLSIMCU_GENERIC_TRANSFORMS", "1")
        generic, tot_g = _run_resident(sc, bunch, KERNEL_FAST, seed=78 + 1000 * k)
        monkeypatch.delenv("CLSIMCU_GENERIC_TRANSFORMS")
        assert tot_b["photons"] == tot_g["photons"] and len(block) > 1e4
        return _compare_distributions(block, generic, tot_b, tot_g)

    _assert_same_distributions(attempt, 2e-3)
