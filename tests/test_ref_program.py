"""The oracle against the reference's WHOLE program, generated parts included.

tests/test_ref_kernel.py pins the oracle to the reference's static kernel text but feeds that text the oracle's own
restatements of what the reference GENERATES at run time (medium functions, wavelength generators, bias, geometry).
Here nothing is restated: the reference's 33 generator sources (private/clsim/function/*.cxx, random_value/*.cxx,
I3CLSimMediumProperties.cxx, private/opencl/I3CLSimHelperGenerateMediumPropertiesSource{,_Optimizers}.cxx) are compiled
unmodified into oracle/_ref/libclsim_ref_medium.so, its geometry generator into libclsim_ref_geometry.so; the text they
write is joined with resources/kernels/*.cl in the order I3CLSimStepToPhotonConverterOpenCL.cxx:655-667 joins it and
compiled for the host (oracle/ref_shim/ref_program.cpp; one syntax rewrite, vector literals).  That is the program the
reference would hand to its OpenCL compiler for the scene.

Bar: every generated function returns the same float as the oracle's restatement of it, bit for bit, and the whole
program's hit lists (80 bytes per record, emission order), hit counters, histories and final RNG states equal the
oracle's on the five BASELINE configurations and every option of the path.

CPU only; needs /root/reference (the kernel text is read where it lies) -- skipped elsewhere."""
import numpy as np
import pytest

from clsim_b200 import steps
from oracle import pyoracle
from tests.scenes import add_flasher_generator, dom_near, make_scene, rng_streams

pytestmark = pytest.mark.skipif(not pyoracle.ref_program_available(), reason="needs /root/reference and oracle/_ref (this container)")

_programs = {}


def program(name, flasher=False, geo_kind="ic86", oversize=5.0, **opts):
    key = (name, flasher, geo_kind, oversize, tuple(sorted(opts.items())))
    if key not in _programs:
        sc = make_scene(name, oversize=oversize, geo_kind=geo_kind)
        if flasher:
            sc = add_flasher_generator(sc)
        opt = sc.options(max_num_workitems=1024, **opts)
        geo = None if opts.get("save_all_photons") else sc.geo
        prog = pyoracle.RefProgram(sc.medium, geo, sc.generators, sc.bias, opt, save_all_dom_stub=bool(opts.get("save_all_photons")))
        ora = pyoracle.Scene(sc.medium, geo, sc.generators, sc.bias, opt)
        _programs[key] = (sc, prog, ora)
    return _programs[key]


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


MEDIA = ["homogeneous", "spice_mie", "spice_mie_tilt", "spice_lea", "spice_lea_notilt"]


# ----------------------------------------------------------------------------------------------------------------
# the generated functions, one by one
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MEDIA)
def test_generated_wavelength_functions_equal_the_oracles(name):
    """getPhaseRefIndex, getGroupVelocity, getScatteringLength, getAbsorptionLength (every layer, with the generator's
    per-layer folding of …_Optimizers.cxx) and getWavelengthBias: same float out of the same float in."""
    sc, prog, ora = program(name)
    L = sc.medium.layersNum
    rng = np.random.default_rng(5)
    wl = np.concatenate([np.linspace(265e-9, 675e-9, 83), rng.uniform(260e-9, 680e-9, 400)]).astype(np.float32)
    for which in (0, 1, 2, 3, 4):
        layers = np.repeat(np.arange(L), len(wl)) if which in (2, 3) else np.zeros(len(wl))
        w = np.tile(wl, L) if which in (2, 3) else wl
        got, want = prog.eval_wlen_function(which, layers, w), ora.eval_wlen_function(which, layers, w)
        assert np.isfinite(got).all() and (got > 0).all()
        assert same_bits(got, want), "function %d: %d of %d values differ" % (which, (got != want).sum(), len(got))


@pytest.mark.parametrize("name", MEDIA)
def test_generated_fields_and_transforms_equal_the_oracles(name):
    """getTiltZShift, getDirectionalAbsLenCorrFactor, transformDirectionPre/PostScatter."""
    sc, prog, ora = program(name)
    rng = np.random.default_rng(6)
    pos = rng.uniform(-700, 700, (4000, 3)).astype(np.float32)
    pos[:50, 2] = rng.uniform(-3000, 3000, 50)          # far outside the tilt table: the clamped rows
    d = rng.normal(size=(4000, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[:3] = np.eye(3, dtype=np.float32)
    if sc.medium.tilt is not None:
        assert same_bits(prog.eval_scalar_field(0, pos), ora.eval_scalar_field(0, pos))
        assert np.abs(prog.eval_scalar_field(0, pos)).max() > 1.0
    else:
        assert "getTiltZShift_IS_CONSTANT" in prog.generated.medium      # the #define the kernel text tests for
        assert not prog.eval_scalar_field(0, pos).any() and not ora.eval_scalar_field(0, pos).any()
    assert same_bits(prog.eval_scalar_field(1, d), ora.eval_scalar_field(1, d))
    for which in (0, 1):
        got = prog.eval_vector_transform(which, d)
        assert same_bits(got, ora.eval_vector_transform(which, d))
        if sc.medium.anisotropy is not None:
            assert np.abs(got - d).max() > 1e-3
            assert np.abs(np.linalg.norm(got.astype(np.float64), axis=1) - 1).max() < 1e-6


@pytest.mark.parametrize("name,flasher", [("homogeneous", False), ("spice_mie", True), ("spice_lea", True)])
def test_generated_samplers_draw_the_same_streams(name, flasher):
    """makeScatteringCosAngle (the Liu / Henyey-Greenstein mix) and generateWavelength (Cherenkov with dispersion on the
    biased spectrum; the LED spectrum on unequal bins): the same values from the same (x, a), the streams left in the same state."""
    sc, prog, ora = program(name, flasher=flasher)
    a, x = rng_streams(24, 99)
    for which in range(0, 2 + (1 if flasher else 0)):
        for i in range(24):
            got, gx = prog.sample(which, x[i], a[i], 2000)
            want, wx = ora.sample(which, x[i], a[i], 2000)
            assert gx == wx and same_bits(got, want), "sampler %d, stream %d" % (which, i)
    cos = prog.sample(0, x[0], a[0], 20000)[0]
    assert abs(cos.mean() - sc.medium.meanCosine) < 0.01 and cos.min() >= -1 and cos.max() <= 1
    wl = prog.sample(1, x[1], a[1], 20000)[0]
    assert 260e-9 <= wl.min() and wl.max() <= 680e-9          # the bias table's range (I3CLSimModuleHelper.cxx:224-)


def test_host_twins_of_the_generated_functions():
    """Every reference class carries the function twice: as device text and as a double-precision host method (GetValue).
    The two agree to float accuracy -- which also says that the objects were put together with the arguments in the right places."""
    sc, prog, _ = program("spice_lea")
    g = prog.generated
    wl = np.linspace(270e-9, 670e-9, 41)
    for layer in (0, 17, 85, 170):
        lay = np.full(len(wl), layer)
        np.testing.assert_allclose(prog.eval_wlen_function(0, lay, wl), [g.host_value(0, layer, w) for w in wl], rtol=3e-6)
        np.testing.assert_allclose(prog.eval_wlen_function(2, lay, wl), [g.host_value(2, layer, w) for w in wl], rtol=2e-5)
        np.testing.assert_allclose(prog.eval_wlen_function(3, lay, wl), [g.host_value(3, layer, w) for w in wl], rtol=2e-5)
        # group velocity = c / group index
        np.testing.assert_allclose(prog.eval_wlen_function(1, lay, wl), [0.299792458 / g.host_value(1, layer, w) for w in wl], rtol=3e-6)
    rng = np.random.default_rng(8)
    pos = rng.uniform(-500, 500, (300, 3))
    np.testing.assert_allclose(prog.eval_scalar_field(0, pos), [g.host_value(4, 0, *p) for p in pos], rtol=2e-4, atol=2e-4)
    d = rng.normal(size=(300, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    np.testing.assert_allclose(prog.eval_scalar_field(1, d), [g.host_value(5, 0, *v) for v in d], rtol=1e-5)
    assert g.host_value(6) == pytest.approx(265e-9) and g.host_value(7) == pytest.approx(675e-9)


def test_generated_text_is_what_the_golden_tables_say():
    """The numbers inside the generated medium text are the layer tables of tests/golden/medium_*.json (the reference's own
    Python loader under recording stubs), as floats."""
    sc, prog, _ = program("spice_lea")
    defines, arrays = pyoracle.parse_generated_source(prog.generated.medium)
    m = sc.medium
    assert defines["MEDIUM_LAYERS"] == m.layersNum == len(arrays["getScatteringLength_b400"])
    assert np.float32(defines["MEDIUM_LAYER_BOTTOM_POS"]) == np.float32(m.layersZStart)
    assert np.float32(defines["MEDIUM_LAYER_THICKNESS"]) == np.float32(m.layersHeight)
    # ToFloatString writes ten digits after the point: the float nearest to the text is the float nearest to the double
    assert np.array_equal(np.array(arrays["getScatteringLength_b400"], np.float32), np.asarray(m.b400, np.float32))
    assert np.array_equal(np.array(arrays["getAbsorptionLength_aDust400"], np.float32), np.asarray(m.aDust400, np.float32))
    assert np.array_equal(np.array(arrays["getAbsorptionLength_deltaTau"], np.float32), np.asarray(m.deltaTau, np.float32))
    assert np.array_equal(np.array(arrays["getTiltZShift_data_zCorrections"], np.float32),
                          np.asarray(m.tilt["zCorrections"], np.float32).ravel())


# ----------------------------------------------------------------------------------------------------------------
# the whole program
# ----------------------------------------------------------------------------------------------------------------
def run_both(prog, ora, bunch, seed=1234, cap=None):
    a, x = rng_streams(len(bunch), seed)
    got = prog.propagate(bunch, x, a, cap=cap)
    want, counted, _, x_cpu, hist = ora.propagate(bunch, x, a, cap=cap, num_threads=4)
    return got, (want, counted, x_cpu, hist), x


def assert_identical(got, want):
    g, g_count, g_x, g_hist = got
    w, w_count, w_x, w_hist = want
    assert g_count == w_count and len(g) == len(w)
    if g.tobytes() != w.tobytes():
        raise AssertionError("hit records differ: %r" % {f: int((g[f] != w[f]).sum()) for f in g.dtype.names if (g[f] != w[f]).any()})
    assert np.array_equal(g_x, w_x)
    if g_hist is not None:
        assert g_hist.tobytes() == w_hist.tobytes()


@pytest.mark.parametrize("name,maker", [
    ("homogeneous", lambda: steps.point_source_steps(1500, 200, seed=1)),                  # config 1
    ("spice_mie", lambda: steps.muon_track_steps(2500, seed=2)),                          # config 2
    ("spice_lea", lambda: steps.muon_bundle_steps(2500, num_muons=20, seed=3)),           # config 3
    ("spice_lea", lambda: steps.cascade_steps(1500, seed=4)),                             # config 4 (shape)
    ("spice_mie_tilt", lambda: steps.cascade_steps(1200, seed=4)),
    ("spice_lea_notilt", lambda: steps.muon_track_steps(1500, seed=12)),
])
def test_whole_program_bit_identical(name, maker):
    sc, prog, ora = program(name)
    got, want, x0 = run_both(prog, ora, maker())
    assert got[1] > 30
    assert_identical(got, want)
    assert not np.array_equal(got[2], x0)


def test_whole_program_config5_flasher_oversize_one():
    """Config 5: two wavelength generators (no -DNO_FLASHER), no PANCAKE_FACTOR, photons starting inside a DOM."""
    sc, prog, ora = program("spice_lea", flasher=True, oversize=1.0)
    assert "#define PANCAKE_FACTOR" not in prog.text and "#define STOP_PHOTONS_ON_DETECTION" in prog.text
    bunch = steps.flasher_steps(3000, dom_near(sc.geo, (0.0, 0.0, -200.0)), seed=5)
    got, want, _ = run_both(prog, ora, bunch)
    assert got[1] > 15
    assert_identical(got, want)


def test_whole_program_photon_history():
    sc, prog, ora = program("spice_lea", photon_history_entries=5)
    got, want, _ = run_both(prog, ora, steps.muon_track_steps(1500, seed=7))
    assert got[3] is not None and len(got[3]) == len(got[0]) > 20 and np.abs(got[3]).sum() > 0
    assert_identical(got, want)


def test_whole_program_fixed_number_of_absorption_lengths():
    sc, prog, ora = program("spice_mie", fixed_number_of_absorption_lengths=3.0)
    assert "#define PROPAGATE_FOR_FIXED_NUMBER_OF_ABSORPTION_LENGTHS 3.0000000000e+00f" in prog.text
    got, want, _ = run_both(prog, ora, steps.muon_track_steps(1200, seed=8))
    assert_identical(got, want)
    assert got[0]["dist_in_abs_lens"].max() <= 3.0 + 1e-4


def test_whole_program_output_overflow():
    sc, prog, ora = program("homogeneous")
    src = tuple(dom_near(sc.geo, (0, 0, 0)) + np.array([3.0, 0, 0]))
    bunch = steps.pad_to_granularity(steps.point_source_steps(200, 200, pos=src, seed=10), 64)
    got, want, x0 = run_both(prog, ora, bunch, cap=100)
    assert got[1] > 1000 and len(got[0]) == 100
    assert_identical(got, want)
    assert np.array_equal(got[2][200:], x0[200:])


def test_whole_program_non_stop_small_detector():
    """StopDetectedPhotons = false on a detector whose indices do not alias in the reference's 32-bit masks (quirk 8)."""
    sc, prog, ora = program("spice_mie", geo_kind="ring", stop_detected_photons=False)
    src = tuple(dom_near(sc.geo, (0.0, 0.0, 0.0)) + np.array([4.0, 1.0, 2.0]))
    got, want, _ = run_both(prog, ora, steps.point_source_steps(1500, 200, pos=src, seed=14))
    assert got[1] > 15
    assert_identical(got, want)


def test_save_all_mode_of_the_reference_does_not_compile():
    """SAVE_ALL_PHOTONS: the reference joins no geometry source (…ConverterOpenCL.cxx:664), yet saveHit calls
    geometryGetDomPosition (propagation_kernel.c.cl:338).  Its program for this mode is ill-formed at this revision; the
    oracle (and the product) read the call as "a DOM at the origin" (deviations)."""
    sc = make_scene("spice_mie")
    opt = sc.options(max_num_workitems=64, stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=0.25)
    with pytest.raises(RuntimeError, match="geometryGetDomPosition"):
        pyoracle.RefProgram(sc.medium, None, sc.generators, sc.bias, opt)


@pytest.mark.parametrize("name", ["spice_mie", "spice_lea"])
def test_whole_program_save_all_with_the_one_line_stand_in(name):
    """... with that one function supplied, the rest of the reference's program in this mode equals the oracle."""
    # (pancake factor 1: with a DOM "at the origin" the pancake correction of saveHit has no meaning, and the oracle leaves it out)
    sc, prog, ora = program(name, stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=0.25, pancake_factor=1.0)
    assert "#define SAVE_ALL_PHOTONS_PRESCALE 2.5000000000e-01f" in prog.text
    bunch = steps.muon_track_steps(256, photons_per_step=100, seed=9)
    got, want, _ = run_both(prog, ora, bunch, cap=len(bunch) * 100)
    assert 0.2 * 25600 < got[1] < 0.3 * 25600
    assert_identical(got, want)


# ----------------------------------------------------------------------------------------------------------------
# the factories that make the wavelength generators (row R3a)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("medium_name", ["spice_mie", "spice_lea"])
def test_wavelength_generator_factories_against_the_references(medium_name):
    """clsim_b200/ice.py restates I3CLSimModuleHelper::makeCherenkovWavelengthGenerator / makeWavelengthGenerator
    (private/clsim/I3CLSimModuleHelper.cxx:75-300, compiled unmodified into libclsim_ref_medium.so).  The text the reference
    generates for the object ITS factory makes equals, character for character, the text it generates for the object built from
    the arrays ice.py computes -- i.e. every table entry agrees to the ten digits the text carries -- on every branch of the
    factories: tabulated bias with and without dispersion, no bias without dispersion (the closed-form sampler), a constant
    bias other than 1 (the made-up 10 nm binning), and a tabulated LED spectrum on unequal bins."""
    from clsim_b200 import ice
    from clsim_b200.description import WlenBias
    sc = make_scene(medium_name)
    medium, table_bias = sc.medium, sc.bias
    cases = [
        ("tabulated bias, dispersion", table_bias, False, None),
        ("tabulated bias, no dispersion", table_bias, True, None),
        ("no bias, no dispersion", WlenBias(constant=1.0), True, None),
        ("no bias, dispersion", WlenBias(constant=1.0), False, None),
        ("constant bias 0.3, dispersion", WlenBias(constant=0.3), False, None),
        ("constant bias 0.3, no dispersion", WlenBias(constant=0.3), True, None),
        ("LED spectrum, tabulated bias", table_bias, False, ice.GetFlasherLED405Spectrum()),
        ("LED spectrum, constant bias", WlenBias(constant=0.5), False, ice.GetFlasherLED405Spectrum()),
    ]
    kinds = set()
    for what, bias, nodisp, spectrum in cases:
        if spectrum is None:
            ours = ice.makeCherenkovWavelengthGenerator(bias, nodisp, medium)
        else:
            ours = ice.makeWavelengthGenerator(spectrum[0], spectrum[1], bias, medium)
        kinds.add(ours.kind)
        text_ours = pyoracle.RefGeneratedSource(medium, [ours], bias).wlen_generators
        text_ref = pyoracle.ref_made_wlen_generator_source(medium, bias, nodisp, spectrum)
        assert text_ours == text_ref, what
        assert "generateWavelength_0" in text_ref
    assert len(kinds) == 3        # equally spaced table, unequally spaced table, closed form


def test_product_tables_hold_the_numbers_of_the_generated_text():
    """The PRODUCT's host-side flattening (csrc/tables.cpp, through clsimcu_describe_tables_from_config, no GPU) against the
    numbers inside the reference's generated text: per-layer medium tables and the wavelength generators' density and
    cumulative tables (the reference normalises and accumulates them in I3CLSimRandomValueInterpolatedDistribution.cxx), as floats."""
    from clsim_b200 import capi
    sc = add_flasher_generator(make_scene("spice_lea"))
    t = capi.describe_tables(sc.medium, sc.geo, sc.generators, sc.bias, sc.options())
    g = pyoracle.RefGeneratedSource(sc.medium, sc.generators, sc.bias)
    _, med = pyoracle.parse_generated_source(g.medium)
    f32 = lambda v: np.asarray(v, dtype=np.float32)   # noqa: E731
    assert np.array_equal(f32(t["medium"]["b400"]), f32(med["getScatteringLength_b400"]))
    assert np.array_equal(f32(t["medium"]["a_dust400"]), f32(med["getAbsorptionLength_aDust400"]))
    assert np.array_equal(f32(t["medium"]["delta_tau"]), f32(med["getAbsorptionLength_deltaTau"]))
    _, gen = pyoracle.parse_generated_source(g.wlen_generators)
    for i, tab in enumerate(t["wlen_generators"]):
        assert np.array_equal(f32(tab["beta"]), f32(gen["_generateWavelength_%ddistYValues" % i])), i
        assert np.array_equal(f32(tab["acu"]), f32(gen["_generateWavelength_%ddistYCumulativeValues" % i])), i
        if tab["xs"]:
            assert np.array_equal(f32(tab["xs"]), f32(gen["_generateWavelength_%ddistXValues" % i])), i


def test_python_host_twins_against_the_references_host_methods():
    """clsim_b200.description.MediumProperties carries double-precision host twins of the medium functions (used by ice.py's
    generator factories and by the table-maker's normalisation): the same doubles as GetValue of the reference's classes."""
    sc, prog, _ = program("spice_lea")
    g, m = prog.generated, sc.medium
    for w in np.linspace(265e-9, 675e-9, 83):
        assert m.GetPhaseRefractiveIndex(w) == pytest.approx(g.host_value(0, 0, w), rel=1e-15)
        assert m.GetGroupRefractiveIndex(w) == pytest.approx(g.host_value(1, 0, w), rel=1e-15)
        for layer in (0, 40, 170):
            assert m.GetScatteringLength(layer, w) == pytest.approx(g.host_value(2, layer, w), rel=1e-13)
            assert m.GetAbsorptionLength(layer, w) == pytest.approx(g.host_value(3, layer, w), rel=1e-13)
    assert m.GetMinWavelength() == g.host_value(6) and m.GetMaxWavelength() == g.host_value(7)


# ----------------------------------------------------------------------------------------------------------------
# the reference's own validation fixture (resources/scripts/compareToPPCredux): 24 DOMs on a ring of 120 m around
# (257, 212, -399), a cascade at the centre, SpiceLea with tilt and anisotropy switched on and off -- its four ice
# directories read where they lie.  The reference compares arrival-time histograms with ppc by eye (a PDF); here its own
# program for each variant is held against the oracle, byte for byte.
# ----------------------------------------------------------------------------------------------------------------
PPC_FIXTURE = "/root/reference/resources/scripts/compareToPPCredux/test_ice_models"


@pytest.mark.parametrize("variant", ["lea", "lea_notilt", "lea_noanisotropy", "lea_notilt_noanisotropy"])
def test_whole_program_on_the_references_ppc_comparison_fixture(variant):
    import os
    from clsim_b200 import geometry, ice
    from tests.scenes import Scene
    directory = os.path.join(PPC_FIXTURE, variant)
    if not os.path.isdir(directory):
        pytest.skip("the reference's test ice models are not here")
    medium = ice.MakeIceCubeMediumProperties(iceDataDirectory=directory, useTiltIfAvailable=True)
    assert (medium.tilt is not None) == ("notilt" not in variant) and (medium.anisotropy is not None) == ("noanisotropy" not in variant)
    centre = (257.0, 212.0, -399.0)
    oversize = 5.0                                                                              # cfg.txt of the fixture: "over-R" 5
    geo = geometry.make_ring_geometry(oversize=oversize, radius=120.0, center=centre)      # generateTestingGeometry.py
    bias = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * oversize)
    sc = Scene(medium, geo, [ice.makeCherenkovWavelengthGenerator(bias, False, medium)], bias, oversize, oversize)
    opt = sc.options(max_num_workitems=1024)
    prog = pyoracle.RefProgram(sc.medium, sc.geo, sc.generators, sc.bias, opt)
    ora = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, opt)
    bunch = steps.cascade_steps(6000, pos=centre, seed=21)       # (the fixture's source: an electron cascade at the centre of the ring)
    got, want, _ = run_both(prog, ora, bunch)
    assert got[1] > 5
    assert_identical(got, want)
