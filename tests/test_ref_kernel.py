"""The oracle's pin: oracle/clsim_oracle.cpp (the hand-written CPU restatement every GPU parity test checks
against) versus the REFERENCE'S OWN KERNEL TEXT -- resources/kernels/{mwcrng_kernel, propagation_kernel.h/.c,
sparse_collision_kernel.h/.c}.cl compiled for the host by oracle/Makefile under oracle/ref_shim/ (a C++ shim for
OpenCL C; one syntax rewrite, vector literals, 13 lines, logged in oracle/_ref/translate.log).

Bar: BIT-IDENTICAL hit lists (all 80 bytes of every record, in emission order), hit counters, photon histories and
final RNG states, on the five BASELINE configurations and on every option of the path.  The one exception is the
reference's non-stop mode, whose "already checked" bit masks are a latent bug (`1 << ulong` is a
32-bit shift, and the DOM mask is indexed by string number, out of bounds for strings >= 64): the oracle restates
the intent, the reference's text drops hits; asserted here as "reference hits are a subset of the oracle's, same
trajectories (RNG states)", and identical on a detector small enough that no two indices alias.

The library exists wherever /root/reference does (this container); the built .so travels to the GPU box.  These
tests run on the CPU."""
import os

import numpy as np
import pytest

from clsim_b200 import steps
from oracle import pyoracle
from tests.scenes import add_flasher_generator, dom_near, make_scene, rng_streams

pytestmark = pytest.mark.skipif(not pyoracle.ref_available(), reason="oracle/_ref not built (no /root/reference at build time)")
THREADS = min(8, os.cpu_count() or 1)


def both(sc, bunch, seed=1234, threads=THREADS, cap=None, **opts):
    a, x = rng_streams(len(bunch), seed)
    opt = sc.options(max_num_workitems=len(bunch), **opts)
    geo = None if opts.get("save_all_photons") else sc.geo
    rs = pyoracle.RefScene(sc.medium, geo, sc.generators, sc.bias, opt)
    assert rs.variant() is not None, "kernel option combination not compiled into oracle/_ref"
    ref = rs.propagate(bunch, x, a, cap=cap, num_threads=threads)
    ora = rs.oracle_propagate(bunch, x, a, cap=cap, num_threads=threads)
    # the restatement inside the _ref library is the same source as libclsim_oracle.so: same output, bit for bit
    osc = pyoracle.Scene(sc.medium, geo, sc.generators, sc.bias, opt)
    want, counted, _, x_cpu, hist = osc.propagate(bunch, x, a, cap=cap, num_threads=threads)
    assert counted == ora[1] and want.tobytes() == ora[0].tobytes() and np.array_equal(x_cpu, ora[2])
    return rs, ref, ora, x


def assert_identical(ref, ora):
    r, r_count, r_x, r_hist = ref
    o, o_count, o_x, o_hist = ora
    assert r_count == o_count
    assert len(r) == len(o)
    if r.tobytes() != o.tobytes():
        bad = {f: int((r[f] != o[f]).sum()) for f in r.dtype.names if (r[f] != o[f]).any()}
        raise AssertionError("hit records differ: %r" % bad)
    assert np.array_equal(r_x, o_x)
    if r_hist is not None:
        assert r_hist.tobytes() == o_hist.tobytes()


@pytest.mark.parametrize("name,variant,maker", [
    ("homogeneous", "ref_stop_pancake_tiltconst", lambda: steps.point_source_steps(5000, 200, seed=1)),              # config 1: 1e6 photons
    ("spice_mie", "ref_stop_pancake_tiltconst", lambda: steps.muon_track_steps(6000, seed=2)),                      # config 2
    ("spice_lea", "ref_stop_pancake_tilt", lambda: steps.muon_bundle_steps(6000, num_muons=20, seed=3)),            # config 3
    ("spice_lea", "ref_stop_pancake_tilt", lambda: steps.cascade_steps(4000, seed=4)),                              # config 4 (shape)
    ("spice_mie_tilt", "ref_stop_pancake_tilt", lambda: steps.cascade_steps(3000, seed=4)),
    ("spice_lea_notilt", "ref_stop_pancake_tiltconst", lambda: steps.muon_track_steps(3000, seed=12)),
])
def test_configs_bit_identical(name, variant, maker):
    sc = make_scene(name)
    bunch = maker()
    rs, ref, ora, x0 = both(sc, bunch)
    assert rs.variant() == variant
    assert ref[1] > 50
    assert_identical(ref, ora)
    assert not np.array_equal(ref[2], x0)


def test_config5_flasher_oversize_one():
    """Config 5: -DNO_FLASHER absent, no PANCAKE_FACTOR, photons start inside a DOM (quirk 9)."""
    sc = add_flasher_generator(make_scene("spice_lea", oversize=1.0))
    bunch = steps.flasher_steps(6000, dom_near(sc.geo, (0.0, 0.0, -200.0)), seed=5)
    rs, ref, ora, _ = both(sc, bunch)
    assert rs.variant() == "ref_stop_flasher_tilt"
    assert ref[1] > 30
    assert_identical(ref, ora)
    # ... and Cherenkov steps through the flasher-capable kernel, with the pancake
    sc = add_flasher_generator(make_scene("spice_mie"))
    rs, ref, ora, _ = both(sc, steps.muon_track_steps(2000, seed=6))
    assert rs.variant() == "ref_stop_pancake_flasher_tiltconst"
    assert_identical(ref, ora)


@pytest.mark.parametrize("name", ["spice_mie", "spice_lea"])
def test_photon_history(name):
    sc = make_scene(name)
    rs, ref, ora, _ = both(sc, steps.muon_track_steps(3000, seed=7), photon_history_entries=5)
    assert "history" in rs.variant() and ref[3] is not None and len(ref[3]) == len(ref[0]) > 50
    assert_identical(ref, ora)
    assert np.abs(ref[3]).sum() > 0


def test_fixed_number_of_absorption_lengths():
    sc = make_scene("spice_mie")
    rs, ref, ora, _ = both(sc, steps.muon_track_steps(2000, seed=8), fixed_number_of_absorption_lengths=3.0)
    assert rs.variant() == "ref_stop_pancake_fixedabs_tiltconst"
    assert_identical(ref, ora)
    assert ref[0]["dist_in_abs_lens"].max() <= 3.0 + 1e-4


@pytest.mark.parametrize("name", ["spice_mie", "spice_lea"])
def test_save_all_photons_with_prescale(name):
    sc = make_scene(name)
    bunch = steps.muon_track_steps(512, photons_per_step=100, seed=9)
    rs, ref, ora, _ = both(sc, bunch, stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=0.25,
                           cap=len(bunch) * 100)
    assert rs.variant().startswith("ref_saveall")
    assert 0.2 * 51200 < ref[1] < 0.3 * 51200
    assert_identical(ref, ora)


def test_output_overflow_and_dummy_steps():
    """Quirk 10 (the counter runs past the capacity, records beyond it are dropped) and quirk 11 (dummy steps draw nothing)."""
    sc = make_scene("homogeneous")
    src = tuple(dom_near(sc.geo, (0, 0, 0)) + np.array([3.0, 0, 0]))
    bunch = steps.pad_to_granularity(steps.point_source_steps(200, 200, pos=src, seed=10), 64)
    rs, ref, ora, x0 = both(sc, bunch, threads=1, cap=100)
    assert ref[1] > 1000 and len(ref[0]) == 100
    assert_identical(ref, ora)
    assert np.array_equal(ref[2][200:], x0[200:]) and not np.array_equal(ref[2][:200], x0[:200])


def test_one_thread_equals_many():
    sc = make_scene("spice_mie")
    bunch = steps.muon_track_steps(1500, seed=11)
    _, ref1, ora1, _ = both(sc, bunch, threads=1)
    _, refn, oran, _ = both(sc, bunch, threads=THREADS)
    assert_identical(ref1, refn)
    assert_identical(ref1, ora1)


def _as_multiset(p):
    out = {}
    for rec in p:
        k = rec.tobytes()
        out[k] = out.get(k, 0) + 1
    return out


def test_non_stop_mode_reference_masks_only_drop_hits():
    sc = make_scene("spice_mie")
    rs, ref, ora, _ = both(sc, steps.muon_track_steps(3000, seed=13), stop_detected_photons=False)
    assert rs.variant() == "ref_nonstop_pancake_tiltconst"
    assert np.array_equal(ref[2], ora[2])                     # same trajectories: detection does not touch the photon
    have = _as_multiset(ora[0])
    for k, n in _as_multiset(ref[0]).items():
        assert have.get(k, 0) >= n                            # every hit of the reference's text is one of the oracle's
    assert 0.5 * ora[1] < ref[1] <= ora[1]
    # a detector without aliasing indices (24 DOMs on 24 strings, all < 32): identical
    ring = make_scene("spice_mie", geo_kind="ring")
    src = tuple(dom_near(ring.geo, (0.0, 0.0, 0.0)) + np.array([4.0, 1.0, 2.0]))
    b = steps.point_source_steps(2000, 200, pos=src, seed=14)
    rs, ref, ora, _ = both(ring, b, stop_detected_photons=False)
    assert ref[1] > 20
    assert_identical(ref, ora)


def test_translation_is_the_vector_literal_rewrite_only():
    """The build log of oracle/_ref lists every line of the kernel text that was rewritten."""
    log = os.path.join(os.path.dirname(pyoracle._REF_LIB), "translate.log")
    with open(log) as f:
        text = f.read()
    minus = [l.strip()[2:] for l in text.split("\n") if l.startswith("  - ")]
    plus = [l.strip()[2:] for l in text.split("\n") if l.startswith("  + ")]
    assert len(minus) == len(plus) == 13
    import re
    lit = re.compile(r"\(\s*(?:const\s+)?(floating4_t|float4|float2|double4)\s*\)")
    for a, b in zip(minus, plus):
        assert lit.sub(lambda m: m.group(1), a).replace(" ", "") == b.replace(" ", "")

