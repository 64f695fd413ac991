"""Steps the fast kernel cannot give an ice layer to: position, time, direction, length or beta not a finite number, or
beyond 2^63.  The reference's own step generator makes one such step per 2^32 draws of its gamma sampler
(I3CLSimLightSourceToStepConverterUtils.h:100-108: log(ry / (1 - ry)) at ry == 1), i.e. once in some fifty runs of
1e10 photons; the reference's kernel flies those photons through its outermost layer until they are absorbed -- they
hit nothing.  The fast kernel counts them as created and ends them on the spot (kernel_fast.cu, fill_queue); what must
never happen is that one of them holds its warp, and with it the launch, for ever.

(The file sorts last on purpose: should a launch hang after all, pytest-timeout ends the process here, behind every
other GPU test.)"""
import numpy as np
import pytest

from clsim_b200 import capi, steps
from clsim_b200.description import KERNEL_FAST, KERNEL_REFERENCE
from tests.scenes import make_scene, sized

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]


def _poisoned(bunch):
    bad = bunch.copy()
    where = {}
    # what the cascade generator makes at ry == 1: `along` = +inf, so every coordinate is +-inf (and the time)
    bad["x"][7], bad["y"][7], bad["z"][7], bad["t"][7] = np.inf, np.inf, -np.inf, np.inf
    where[7] = "cascade step at infinity"
    bad["z"][100] = np.nan
    where[100] = "z is not a number"
    bad["z"][1000] = -3e30
    where[1000] = "below the lower end of the layer table"
    bad["z"][1001] = 3e30
    where[1001] = "above the upper end of the layer table"
    bad["x"][2000] = 1e25
    where[2000] = "x beyond 2^63"
    bad["t"][2048] = np.inf
    where[2048] = "time is infinite"
    bad["theta"][3000] = np.nan
    where[3000] = "direction is not a number"
    bad["beta"][len(bad) - 1] = np.nan
    where[len(bad) - 1] = "beta is not a number"
    return bad, sorted(where)


@pytest.mark.parametrize("mode", [KERNEL_FAST, KERNEL_REFERENCE])   # (the reference-order kernel takes the same way out)
@pytest.mark.parametrize("name", ["spice_mie", "spice_lea"])     # plain layers; tilt + anisotropy (the ice of configs 3-5)
def test_steps_at_infinity_end_at_once(name, mode):
    sc = make_scene(name)
    bunch = steps.muon_track_steps(sized(1 << 15, 1 << 12), seed=91)   # ~7000 hits (oracle, both ice models)
    bunch["identifier"] = np.arange(len(bunch))
    bad, where = _poisoned(bunch)
    opt = sc.options(kernel_mode=mode, max_num_workitems=len(bunch), rng_seed=17, output_photons_per_workitem=4)
    with capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt) as eng:
        eng.upload_resident(bunch)
        clean = eng.run_resident(1)
        clean_hits = eng.download_resident()
        eng.upload_resident(bad)
        r = eng.run_resident(1)          # must come back
        hits = eng.download_resident()
        # ... and the same through the queueing interface, host buffers in, hit list out
        eng.enqueue(bad, 5)
        res = eng.get_result()
    total = int(bunch["num_photons"].sum())
    assert clean["photons"] == total and r["photons"] == total          # counted as created
    assert len(clean_hits) > sized(3000, 300)
    for h in (hits, res.photons):
        assert not np.isin(h["identifier"], where).any()                # they hit nothing
        assert np.all(np.isfinite(h["x"])) and np.all(np.isfinite(h["t"])) and np.all(np.isfinite(h["cherenkov_dist"]))
        # the other steps are untouched: as many hits as before, within the fluctuation of a count
        expected = len(clean_hits) * (1.0 - len(where) / len(bunch))
        assert abs(len(h) - expected) < 6.0 * np.sqrt(expected) + 0.01 * expected
