"""Step generation without a GPU: the oracle's samplers against their analytic distributions,
the record layout, the host-side queue logic of the converter mirror, and the C ABI's argument checks."""
import ctypes as C
import math

import numpy as np
import pytest
from scipy import stats

from clsim_b200 import capi, ice, stepgen
from clsim_b200.description import STEP_DTYPE
from oracle import stepgen_oracle as so


def streams(n, seed=3):
    a = capi.safeprime_multipliers(500, n)
    rng = np.random.default_rng(seed)
    x = (rng.integers(1, 2 ** 31, n).astype(np.uint64) << np.uint64(20)) + rng.integers(1, 2 ** 20, n).astype(np.uint64)
    return x, a


def test_layouts_agree():
    assert so.SOURCE_DTYPE == stepgen.SOURCE_DTYPE and so.SOURCE_DTYPE.itemsize == 104
    assert C.sizeof(stepgen.StepGeneratorConfigStruct) == 40
    assert so.SOURCE_DTYPE.fields["num_steps"][1] == 80 and so.SOURCE_DTYPE.fields["kind"][1] == 100


@pytest.mark.parametrize("shape", [0.5, 2.03, 6.5])
def test_gamma_sampler_is_gamma_distributed(shape):
    x, a = streams(1)
    rng = so.Mwc(x[0], a[0])
    v = np.array([so.gamma_distributed(shape, rng) for _ in range(20000)])
    assert stats.kstest(v, "gamma", args=(shape,)).pvalue > 0.01
    assert abs(v.mean() - shape) < 5 * math.sqrt(shape / len(v))


def test_angular_smearing_and_rotation():
    # cos = max(1 - (-ln(1 - U I)/b)^(1/a), -1): its CDF is (1 - exp(-b (1-cos)^a)) / I   (…PPC.cxx:752-754)
    a_, b_ = 0.39, 2.61
    big_i = 1.0 - math.exp(-b_ * 2.0 ** a_)
    src = np.zeros(1, dtype=so.SOURCE_DTYPE)
    src["dir_z"] = 1.0
    src["length"], src["num_steps"], src["photons_per_step"], src["kind"] = 100.0, 20000, 200, so.TRACK_CASCADE_LIKE
    x, a = streams(64)
    steps, x_after = so.make_steps(src, x, a, STEP_DTYPE)
    assert len(steps) == 20000 and np.all(steps["num_photons"] == 200) and np.all(steps["length"] == np.float32(0.001))
    cos_t = np.cos(steps["theta"].astype(np.float64))   # axis = +z, so theta is the smearing angle
    cdf = lambda c: 1.0 - (1.0 - np.exp(-b_ * np.power(np.maximum(1.0 - c, 0.0), a_))) / big_i
    assert stats.kstest(cos_t, cdf).pvalue > 0.01
    assert stats.kstest(steps["phi"] / (2 * math.pi), "uniform").pvalue > 0.01
    # positions uniform along the track, time = distance / c
    assert stats.kstest(steps["z"] / 100.0, "uniform").pvalue > 0.01
    assert np.allclose(steps["t"], steps["z"] / so.C_LIGHT, rtol=1e-6)
    assert np.all(steps["x"] == 0) and np.all(steps["y"] == 0)
    assert not np.array_equal(x_after, x)


def test_step_layout_of_queue_entries():
    src = np.zeros(3, dtype=so.SOURCE_DTYPE)
    src["dir_x"] = 1.0
    src["kind"] = [so.TRACK_MUON_LIKE, so.CASCADE, so.TRACK_CASCADE_LIKE]
    src["length"] = [800.0, 0.0, 50.0]
    src["pa"], src["pb"] = [0, 4.5, 0], [0, 0.6, 0]
    src["num_steps"], src["photons_per_step"], src["photons_in_last_step"] = [3, 5, 0], [200, 200, 200], [17, 0, 9]
    src["identifier"] = [7, 8, 9]
    src["x"], src["t"] = [10.0, 20.0, 30.0], [1.0, 2.0, 3.0]
    x, a = streams(4)
    steps, _ = so.make_steps(src, x, a, STEP_DTYPE)
    assert list(steps["identifier"]) == [7] * 4 + [8] * 5 + [9]
    assert list(steps["num_photons"]) == [200, 200, 200, 17] + [200] * 5 + [9]
    # muon-like: the whole track from the vertex, direction untouched
    assert np.all(steps["length"][:4] == 800.0) and np.all(steps["x"][:4] == 10.0) and np.all(steps["t"][:4] == 1.0)
    assert np.allclose(steps["theta"][:4], math.pi / 2) and np.allclose(steps["phi"][:4], 0.0, atol=1e-6)
    # cascade: downstream of the vertex by pb * Gamma(pa), in time with the speed of light
    along = steps["x"][4:9] - 20.0
    assert np.all(along > 0) and np.allclose(steps["t"][4:9] - 2.0, along / so.C_LIGHT, rtol=1e-4)
    assert 30.0 <= steps["x"][9] < 80.0


def test_converter_mirror_queue_logic(has_gpu):
    if not has_gpu:
        # the host logic needs a generator object only at Initialize: without a device it must fail loudly
        conv = stepgen.I3CLSimLightSourceToStepConverterPPC()
        medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_mie", useTiltIfAvailable=False)
        conv.SetMediumProperties(medium)
        conv.SetWlenBias(ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * 5))
        with pytest.raises(stepgen.I3CLSimLightSourceToStepConverter_exception, match="RandomService not set"):
            conv.Initialize()
        conv.SetRandomService(1)
        with pytest.raises(capi.ClsimCudaError, match="no CPU fallback"):
            conv.Initialize()
    with pytest.raises(stepgen.I3CLSimLightSourceToStepConverter_exception, match="may not be <= 0"):
        stepgen.I3CLSimLightSourceToStepConverterPPC(photonsPerStep=0)
    conv = stepgen.I3CLSimLightSourceToStepConverterPPC()
    with pytest.raises(stepgen.I3CLSimLightSourceToStepConverter_exception, match="!= 1 is currently not supported"):
        conv.SetBunchSizeGranularity(2)
    with pytest.raises(stepgen.I3CLSimLightSourceToStepConverter_exception, match="not initialized"):
        conv.EnqueueLightSource(stepgen.Particle("MuMinus", 1e3, (0, 0, 0), (0, 0, 1), length=100.0), 0)


def test_photon_yield_per_metre():
    # unbiased Frank-Tamm yield between 265 and 675 nm in ice (n ~ 1.32-1.36) is about 4.5e4 photons per metre;
    # with the DOM acceptance as bias it drops by the mean acceptance
    medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_mie", useTiltIfAvailable=False)
    from clsim_b200.description import WlenBias
    flat = stepgen.NumberOfPhotonsPerMeter(medium, WlenBias(constant=1.0), 265e-9, 675e-9)
    n = medium.GetPhaseRefractiveIndex(400e-9)
    approx = 2 * math.pi / 137.0 * (1 - 1 / n ** 2) * (1 / 265e-9 - 1 / 675e-9)
    assert abs(flat / approx - 1) < 0.03 and 4e4 < flat < 5e4
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651 * 5)
    biased = stepgen.NumberOfPhotonsPerMeter(medium, acc, 265e-9, 675e-9)
    assert 0 < biased < flat * np.max(acc.values)


def test_shower_parameters_shape():
    a, b, scale, sigma = stepgen.ShowerParameters("EMinus", 1e3)
    assert scale == 1.0 and sigma == 0.0 and 5 < a < 8 and 0.4 < b < 0.8
    a_h, b_h, scale_h, sigma_h = stepgen.ShowerParameters("Hadrons", 1e3)
    assert 0.5 < scale_h < 1.0 and 0 < sigma_h < 0.2 and a_h < a and b_h > b


def test_abi_argument_checks():
    lib = stepgen._lib()
    h = C.c_void_p()
    cfg = stepgen.StepGeneratorConfigStruct()
    cfg.struct_size = 8
    assert lib.clsimcu_stepgen_create(C.byref(cfg), C.byref(h)) == -1 and b"struct_size" in lib.clsimcu_last_error()
    cfg.struct_size = C.sizeof(cfg)
    assert lib.clsimcu_stepgen_create(C.byref(cfg), C.byref(h)) == -1 and b"angular" in lib.clsimcu_last_error()
    n = C.c_size_t(0)
    assert lib.clsimcu_stepgen_generate(None, None, 0, None, 0, C.byref(n)) == -4
    assert lib.clsimcu_enqueue_sources(None, None, None, 0, 0) == -4


# ---------------------------------------------------------------------------------------------------------------------
# The samplers against the reference's own inline functions (private/clsim/I3CLSimLightSourceToStepConverterUtils.h),
# compiled unmodified into oracle/_ref/libclsim_ref_stepgen.so.  Bit for bit: same libm, same float temporaries.
# ---------------------------------------------------------------------------------------------------------------------
from oracle import pyoracle  # noqa: E402

needs_ref = pytest.mark.skipif(not pyoracle.ref_stepgen_available(), reason="oracle/_ref/libclsim_ref_stepgen.so not built (no /root/reference at build time)")


@needs_ref
def test_mwc_draws_equal_the_reference():
    L = pyoracle.ref_stepgen_lib()
    x, a = streams(32, seed=5)
    for i in range(32):
        rng = so.Mwc(x[i], a[i])
        state = C.c_uint64(int(x[i]))
        for k in range(200):
            if k % 2:
                assert rng.co() == L.ref_mwc_co(C.byref(state), int(a[i]))
            else:
                assert rng.oc() == L.ref_mwc_oc(C.byref(state), int(a[i]))
            assert rng.x == state.value


@needs_ref
@pytest.mark.parametrize("shape", [0.05, 0.3, 0.5, 0.99, 1.0, 1.5, 2.03, 4.7, 6.5, 17.2, 250.0])
def test_gamma_sampler_equals_the_reference_bit_for_bit(shape):
    """Weibull branch below shape 1, Cheng's above, float temporaries and all (…Utils.h:78-111): the same value from the
    same stream, and the same number of draws consumed (the rejection loops leave the stream in the same state)."""
    L = pyoracle.ref_stepgen_lib()
    x, a = streams(16, seed=int(shape * 100) + 1)
    for i in range(16):
        rng = so.Mwc(x[i], a[i])
        state = C.c_uint64(int(x[i]))
        for _ in range(500):
            ours = so.gamma_distributed(shape, rng)
            theirs = L.ref_gamma_distributed(shape, C.byref(state), int(a[i]))
            assert ours == theirs and rng.x == state.value


@needs_ref
def test_rotation_equals_the_reference_bit_for_bit():
    L = pyoracle.ref_stepgen_lib()
    rng = np.random.default_rng(8)
    cases = []
    for _ in range(3000):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        cosa = rng.uniform(-1.0, 1.0)
        cases.append((cosa, math.sqrt(1.0 - cosa * cosa), d[0], d[1], d[2], rng.uniform()))
    # vertical directions take the other branch (sin(theta) == 0), both signs of z, and the degenerate angles
    cases += [(0.3, math.sqrt(1 - 0.09), 0.0, 0.0, 1.0, 0.123), (0.3, math.sqrt(1 - 0.09), 0.0, 0.0, -1.0, 0.9), (1.0, 0.0, 0.6, 0.0, 0.8, 0.5),
              (-1.0, 0.0, 0.0, 0.6, -0.8, 0.25), (0.0, 1.0, 1.0, 0.0, 0.0, 0.0)]
    for cosa, sina, x, y, z, r in cases:
        xyz = (C.c_double * 3)(x, y, z)
        L.ref_scatter_direction_by_angle(cosa, sina, xyz, r)
        assert so.scatter_direction(cosa, sina, x, y, z, r) == (xyz[0], xyz[1], xyz[2])


@needs_ref
def test_stream_seeding_rule_equals_the_reference():
    """mwcRngInitState (…Utils.h:49-61): x != 0, high word < a - 1, low word < 2^32 - 1, drawn as (high, low) until it
    fits.  The repo draws the two words from one splitmix64 value (tables.cpp seed_rng_states, restated in the oracle);
    handed the same words, the reference's function accepts the same candidate."""
    L = pyoracle.ref_stepgen_lib()

    def splitmix(seed, n):
        out, state, mask = [], seed, (1 << 64) - 1
        for _ in range(n):
            state = (state + 0x9E3779B97F4A7C15) & mask
            z = state
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
            out.append(z ^ (z >> 31))
        return out

    rejected = 0
    for seed in range(40):
        # real multipliers (a rejection in ~3 % of the draws) and small ones (255 of 256, 15 of 16 high words are too large)
        a = np.concatenate([capi.safeprime_multipliers(100 * seed, 24), np.array([0x01000000, 0x10000000], dtype=np.uint32)]).astype(np.uint32)
        ours = pyoracle.seed_states(seed, a)
        words = splitmix(seed, 20000)
        flat = np.array([w for v in words for w in (v >> 32, v & 0xFFFFFFFF)], dtype=np.uint32)
        at = 0
        for i in range(len(a)):
            used = C.c_size_t(0)
            theirs = L.ref_mwc_init_state(flat[at:].ctypes.data, len(flat) - at, int(a[i]), C.byref(used))
            assert used.value >= 2 and used.value % 2 == 0 and at + used.value < len(flat)
            rejected += used.value // 2 - 1
            at += used.value
            assert theirs == int(ours[i])
    assert rejected > 40


@needs_ref
def test_the_draw_that_puts_a_step_at_infinity():
    """ry = 1 - co() is exactly 1 when the draw's low word is 0 (once in 2^32): the reference's Cheng branch divides by
    zero, carries +inf through exp() and a NaN through its rejection test, and returns +inf -- the position of the step
    along the shower axis.  This is the step that held rank 6 of the 8-GPU run (DESIGN section 5,
    tools/replay_config4_stepgen.py); the oracle restates it, the kernels end such a step on the spot."""
    L = pyoracle.ref_stepgen_lib()
    a = int(capi.safeprime_multipliers(7, 1)[0])
    # a state whose SECOND draw (ry) has a zero low word: s1 = (h1, l1) with l1 a + h1 = 0 mod 2^32, s0 its predecessor
    l1 = 0x12345678
    h1 = (-(l1 * a)) & 0xFFFFFFFF
    s1 = (h1 << 32) | l1
    s0 = ((s1 % a) << 32) | (s1 // a)
    assert s1 // a < (1 << 32)
    check = so.Mwc(s0, a)
    check.oc()
    assert check.x == s1 and check.oc() == 1.0
    state = C.c_uint64(s0)
    theirs = L.ref_gamma_distributed(3.2, C.byref(state), a)
    rng = so.Mwc(s0, a)
    ours = so.gamma_distributed(3.2, rng)
    assert math.isinf(theirs) and theirs > 0 and ours == theirs and rng.x == state.value
