"""Step generation on the device against the oracle: every step of every stream replayed;
the reference-shaped converter; bunches generated and propagated without visiting the host."""
import math

import numpy as np
import pytest

from clsim_b200 import capi, stepgen, steps as steplib
from clsim_b200.description import KERNEL_FAST, STEP_DTYPE
from oracle import stepgen_oracle as so
from clsim_b200.sharding import mcpe_row_offset, stepgen_row_offset
from tests.scenes import make_scene

pytestmark = pytest.mark.gpu


def sources_mixed(seed=1):
    rng = np.random.default_rng(seed)
    src = np.zeros(40, dtype=stepgen.SOURCE_DTYPE)
    d = rng.normal(size=(40, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    d[0] = (0, 0, 1)
    d[1] = (0, 0, -1)       # vertical axes take the other branch of scatterDirectionByAngle
    src["dir_x"], src["dir_y"], src["dir_z"] = d.T
    src["x"], src["y"], src["z"] = rng.uniform(-400, 400, (3, 40))
    src["t"] = rng.uniform(0, 1000, 40)
    src["kind"] = rng.integers(0, 3, 40)
    src["kind"][:2] = [so.CASCADE, so.TRACK_CASCADE_LIKE]
    src["length"] = rng.uniform(1, 900, 40)
    src["pa"] = np.where(rng.uniform(size=40) < 0.2, rng.uniform(0.3, 0.95, 40), rng.uniform(1.5, 9.0, 40))  # both gamma algorithms
    src["pb"] = rng.uniform(0.3, 0.8, 40)
    src["num_steps"] = rng.integers(0, 900, 40)
    src["photons_per_step"] = 200
    src["photons_in_last_step"] = np.where(rng.uniform(size=40) < 0.5, 0, rng.integers(1, 200, 40))
    src["identifier"] = np.arange(40)
    return src


def close_steps(got, want):
    """Records equal up to the last bits of the transcendental functions (CUDA's and glibc's libm differ by <= 1-2 ulp
    in double; after rounding to float almost every field is identical)."""
    assert len(got) == len(want)
    for k in ("num_photons", "identifier", "source_type", "weight", "beta", "length"):
        assert np.array_equal(got[k], want[k]), k
    ok = np.ones(len(got), bool)
    for k, tol in (("x", 2e-4), ("y", 2e-4), ("z", 2e-4), ("t", 1e-3)):
        ok &= np.abs(got[k].astype(np.float64) - want[k]) <= tol * (1 + np.abs(want[k]) * 1e-3)
    dphi = np.abs(got["phi"].astype(np.float64) - want["phi"])
    dphi = np.minimum(dphi, 2 * math.pi - dphi)
    ok &= (np.abs(got["theta"].astype(np.float64) - want["theta"]) <= 2e-6) & ((dphi <= 2e-5) | (np.sin(want["theta"]) < 1e-3))
    return ok


def test_every_step_of_every_stream_replayed():
    gen = stepgen.StepGenerator(rng_seed=99, rng_first_multiplier=stepgen_row_offset(1))
    x0, a = gen.rng_state()
    assert np.array_equal(a, capi.safeprime_multipliers(stepgen_row_offset(1), len(a)))
    src = sources_mixed()
    # make some entries long enough that streams are used more than once
    src["num_steps"][5] = 2 * len(a) + 17
    got = gen.generate(src)
    x1, _ = gen.rng_state()
    want, x_want = so.make_steps(src, x0, a, STEP_DTYPE)
    ok = close_steps(got, want)
    assert ok.mean() > 0.9995, (~ok).nonzero()[0][:10]
    assert np.mean(x1 == x_want) > 0.9995     # a flipped rejection in a gamma draw shifts one stream
    exact = np.mean([got[k].tobytes() == want[k].tobytes() for k in range(0, len(got), 7)])
    assert exact > 0.98
    # second call continues the streams
    got2 = gen.generate(src[:8])
    want2, _ = so.make_steps(src[:8], x_want, a, STEP_DTYPE)
    assert close_steps(got2, want2).mean() > 0.999
    assert len(gen.generate(src[:0])) == 0
    gen.close()


def test_argument_errors():
    gen = stepgen.StepGenerator(rng_seed=1)
    src = sources_mixed()[:3]
    bad = src.copy(); bad["kind"][1] = 7
    with pytest.raises(capi.ClsimCudaError, match="unknown step source kind"):
        gen.generate(bad)
    bad = src.copy(); bad["dir_x"][0] = 3.0
    with pytest.raises(capi.ClsimCudaError, match="not a unit vector"):
        gen.generate(bad)
    bad = src.copy(); bad["kind"][2] = so.TRACK_CASCADE_LIKE; bad["length"][2] = 0.0
    with pytest.raises(capi.ClsimCudaError, match="cascade segment with length"):
        gen.generate(bad)
    gen.close()


def test_reference_shaped_converter_yields_and_barrier():
    sc = make_scene("spice_mie")
    conv = stepgen.I3CLSimLightSourceToStepConverterPPC(photonsPerStep=200)
    conv.SetMediumProperties(sc.medium)
    conv.SetWlenBias(sc.bias)
    conv.SetRandomService(5)
    conv.SetMaxBunchSize(4096)
    conv.Initialize()
    with pytest.raises(stepgen.I3CLSimLightSourceToStepConverter_exception, match="already initialized"):
        conv.SetMaxBunchSize(1)
    with pytest.raises(stepgen.I3CLSimLightSourceToStepConverter_exception, match="no particle is enqueued"):
        conv.GetConversionResult()
    mu = stepgen.Particle("MuMinus", 1e4, (0, 0, -300), (0.3, 0.1, 0.9), time=5.0, length=600.0)
    conv.EnqueueLightSource(mu, 3)
    conv.EnqueueLightSource(stepgen.Particle("EMinus", 50.0, (10, 20, 30), (0, 1, 0), time=7.0), 4)
    conv.EnqueueBarrier()
    with pytest.raises(stepgen.I3CLSimLightSourceToStepConverter_exception, match="barrier is enqueued"):
        conv.EnqueueLightSource(mu, 5)
    series, reset = [], False
    while conv.MoreStepsAvailable():
        s, reset = conv.GetConversionResultWithBarrierInfo()
        assert len(s) <= 4096 + 1
        series.append(s)
    assert reset and not conv.BarrierActive()
    allsteps = np.concatenate(series)
    per_m = conv.meanPhotonsPerMeter
    extr = 1 + max(0.0, 0.1880 + 0.0206 * math.log(1e4))
    mu_steps = allsteps[allsteps["identifier"] == 3]
    long_steps = mu_steps[mu_steps["length"] > 1.0]
    assert np.all(long_steps["length"] == 600.0) and np.all(long_steps["z"] == -300.0)
    n_mu = long_steps["num_photons"].sum()
    assert abs(n_mu - per_m * 600.0) < 6 * math.sqrt(per_m * 600.0)
    n_ca = mu_steps[mu_steps["length"] < 1.0]["num_photons"].sum()
    assert abs(n_ca - per_m * 600.0 * (extr - 1)) < 0.02 * per_m * 600.0 * (extr - 1) + 400
    # cascade-like steps sit on the track
    c = mu_steps[mu_steps["length"] < 1.0]
    d = np.array(mu.dir)
    along = (np.stack([c["x"], c["y"], c["z"] + 300.0], 1) * d).sum(1)
    assert along.min() >= -1e-3 and along.max() <= 600.0 + 1e-3
    off = np.linalg.norm(np.stack([c["x"], c["y"], c["z"] + 300.0], 1) - along[:, None] * d, axis=1)
    assert off.max() < 1e-3
    em = allsteps[allsteps["identifier"] == 4]
    want = per_m * 5.21 * 0.924 / 0.9216 * 50.0
    assert abs(em["num_photons"].sum() - want) < 6 * math.sqrt(want)
    assert np.all(em["y"] >= 20.0) and 0.5 < (em["y"] - 20.0).mean() < 5.0     # shower maximum a few metres downstream


def test_bunches_generated_and_propagated_on_the_device():
    sc = make_scene("spice_mie")
    n_steps = 1 << 15
    opt = sc.options(kernel_mode=KERNEL_FAST, max_num_workitems=n_steps, rng_seed=5)
    gen = stepgen.StepGenerator(rng_seed=8, rng_first_multiplier=stepgen_row_offset(2))
    # the muon-track workload of the benchmark as two queue entries per muon
    track = dict(x=-300.0, y=50.0, z=-350.0, t=0.0, dir_x=math.sqrt(0.5), dir_y=0.0, dir_z=math.sqrt(0.5), length=1000.0, photons_per_step=200)
    src = np.zeros(2, dtype=stepgen.SOURCE_DTYPE)
    for k, v in track.items():
        src[k] = v
    src["kind"] = [so.TRACK_MUON_LIKE, so.TRACK_CASCADE_LIKE]
    src["num_steps"] = [n_steps * 3 // 4, n_steps // 4]
    src["identifier"] = [1, 2]
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        x0, a = gen.rng_state()
        gen.enqueue_into(eng, src, 21)
        res = eng.get_result()
        assert res.identifier == 21 and res.num_photons_generated == n_steps * 200
        frac_dev = len(res.photons) / float(res.num_photons_generated)
        # the same bunch made by the oracle and sent from the host
        want, _ = so.make_steps(src, x0, a, STEP_DTYPE)
        eng.enqueue(want, 22)
        res_host = eng.get_result()
        frac_host = len(res_host.photons) / float(res_host.num_photons_generated)
        assert len(res.photons) > 2000
        assert abs(frac_dev - frac_host) < 6 * math.sqrt(frac_host / res.num_photons_generated) * 1.5
        assert set(np.unique(res.photons["identifier"])) <= {1, 2}
        # preconditions of EnqueueSteps hold for generated bunches too
        big = src.copy(); big["num_steps"][0] = n_steps
        with pytest.raises(capi.ClsimCudaError, match="greater than maximum number of work items"):
            gen.enqueue_into(eng, big, 23)
        empty = src.copy(); empty["num_steps"] = 0
        with pytest.raises(capi.ClsimCudaError, match="Steps are empty"):
            gen.enqueue_into(eng, empty, 24)
    gen.close()


def test_converter_feeds_the_engine_without_host_steps():
    sc = make_scene("spice_mie")
    opt = sc.options(kernel_mode=KERNEL_FAST, max_num_workitems=1 << 15, rng_seed=6, enable_double_buffering=True)
    conv = stepgen.I3CLSimLightSourceToStepConverterPPC(photonsPerStep=200)
    conv.SetMediumProperties(sc.medium)
    conv.SetWlenBias(sc.bias)
    conv.SetRandomService(17)
    conv.Initialize(rngFirstMultiplierRow=stepgen_row_offset(3))
    for i in range(6):
        conv.EnqueueLightSource(stepgen.Particle("MuMinus", 1e3, (-200 + 50 * i, 10, -300), (0.5, 0.1, 0.86), length=700.0), i)
    conv.EnqueueBarrier()
    total_steps, bunches = 0, 0
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        while conv.MoreStepsAvailable():
            n = conv.EnqueueInto(eng, 100 + bunches)
            if n == 0:
                break
            total_steps += n
            bunches += 1
        assert bunches >= 1 and not conv.BarrierActive()
        hits = 0
        generated = 0
        for _ in range(bunches):
            r = eng.get_result()
            hits += len(r.photons)
            generated += r.num_photons_generated
            assert set(np.unique(r.photons["identifier"])) <= set(range(6))
    per_m = conv.meanPhotonsPerMeter
    extr = 1 + max(0.0, 0.1880 + 0.0206 * math.log(1e3))
    assert abs(generated - 6 * per_m * 700.0 * extr) < 0.02 * 6 * per_m * 700.0 * extr
    assert hits > 100
