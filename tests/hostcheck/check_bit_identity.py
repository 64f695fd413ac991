"""Run by tests/test_hostcheck.py in a subprocess with CLSIMCU_LIB = the host check build (not collected on its own: the name
does not match test_*.py).  The reference-order CUDA kernel's source, compiled for the host and driven through the engine,
against the oracle: the same bytes -- hit records in the oracle's order after the engine's own ordering is undone (the engine
returns hits in the order the kernel's atomic counter handed out slots: work-item order on one host thread), final RNG
states, histories -- on the five BASELINE configurations and the options of the path.  On the GPU the same comparison holds
to the last bits of CUDA's libm (tests/test_gpu_reference_kernel.py); here both sides run on glibc."""
import os

import numpy as np
import pytest

from clsim_b200 import capi, steps
from clsim_b200.description import KERNEL_REFERENCE
from oracle import pyoracle
from tests.scenes import add_flasher_generator, dom_near, make_scene, rng_streams

pytestmark = pytest.mark.gpu
assert "hostcheck" in capi.LIB_PATH, "this file is for the host check build only"


def both(sc, bunch, seed=1234, **opts):
    a, x = rng_streams(len(bunch), seed)
    opt = sc.options(kernel_mode=KERNEL_REFERENCE, max_num_workitems=len(bunch), rng_n=len(bunch), rng_a=a, rng_x=x, output_photons_per_workitem=8, **opts)
    geo = None if opts.get("save_all_photons") else sc.geo
    with capi.Engine(sc.medium, geo, sc.generators, sc.bias, opt) as eng:
        eng.enqueue(bunch, 7)
        got = eng.get_result()
        x_dev, _ = eng.rng_get(len(bunch))
    osc = pyoracle.Scene(sc.medium, geo, sc.generators, sc.bias, opt)
    want, counted, _, x_cpu, hist = osc.propagate(bunch, x, a, cap=8 * len(bunch), num_threads=min(4, os.cpu_count() or 1))
    return got, want, counted, x_dev, x_cpu, hist


def identical(got, want, counted, x_dev, x_cpu):
    assert got.num_hits_counted == counted == len(want) and len(got.photons) == len(want)
    assert got.photons.tobytes() == want.tobytes(), {f: int((got.photons[f] != want[f]).sum()) for f in want.dtype.names if (got.photons[f] != want[f]).any()}
    assert np.array_equal(x_dev, x_cpu)


@pytest.mark.parametrize("name,maker", [
    ("homogeneous", lambda: steps.point_source_steps(1024, 200, seed=1)),
    ("spice_mie", lambda: steps.muon_track_steps(2048, seed=2)),
    ("spice_lea", lambda: steps.muon_bundle_steps(2048, num_muons=20, seed=3)),
    ("spice_lea", lambda: steps.cascade_steps(1024, seed=4)),
    ("spice_mie_tilt", lambda: steps.cascade_steps(1024, seed=4)),
])
def test_configs(name, maker):
    got, want, counted, x_dev, x_cpu, _ = both(make_scene(name), maker())
    assert counted > 30
    identical(got, want, counted, x_dev, x_cpu)


def test_flasher_oversize_one():
    sc = add_flasher_generator(make_scene("spice_lea", oversize=1.0))
    bunch = steps.flasher_steps(2048, dom_near(sc.geo, (0.0, 0.0, -200.0)), seed=5)
    got, want, counted, x_dev, x_cpu, _ = both(sc, bunch)
    assert counted > 10
    identical(got, want, counted, x_dev, x_cpu)


def test_history_fixed_absorption_lengths_and_save_all():
    sc = make_scene("spice_lea")
    got, want, counted, x_dev, x_cpu, hist = both(sc, steps.muon_track_steps(1024, seed=7), photon_history_entries=5)
    identical(got, want, counted, x_dev, x_cpu)
    h = got.history
    assert h is not None and len(h) == len(want)
    # the ABI hands the rows over in forward order (the oldest first; the reference's host code unrolls the kernel's ring
    # buffer, ...ConverterOpenCL.cxx:940-989), the oracle returns the ring as the kernel leaves it: the same rows, rotated.
    # Unused rows are NaN on the ABI's side, zero in the oracle's buffer.
    n = np.minimum(want["num_scatters"], 5)
    for i in range(len(want)):
        mine = np.asarray(h[i][:n[i]], np.float32)
        ring = np.asarray(hist[i], np.float32)
        theirs = ring if want["num_scatters"][i] <= 5 else np.roll(ring, -(int(want["num_scatters"][i]) % 5), axis=0)
        assert np.array_equal(mine, theirs[:n[i]]), i
    sc = make_scene("spice_mie")
    got, want, counted, x_dev, x_cpu, _ = both(sc, steps.muon_track_steps(1024, seed=8), fixed_number_of_absorption_lengths=3.0)
    identical(got, want, counted, x_dev, x_cpu)
    got, want, counted, x_dev, x_cpu, _ = both(sc, steps.muon_track_steps(128, photons_per_step=20, seed=9), stop_detected_photons=False,
                                               save_all_photons=True, save_all_photons_prescale=0.25, pancake_factor=1.0)
    assert 0.2 * 2560 < counted < 0.3 * 2560      # (below the 8 records per work-item the engine was given room for)
    identical(got, want, counted, x_dev, x_cpu)


def test_non_stop_detection_on_a_small_detector():
    sc = make_scene("spice_mie", geo_kind="ring")
    src = tuple(dom_near(sc.geo, (0.0, 0.0, 0.0)) + np.array([4.0, 1.0, 2.0]))
    got, want, counted, x_dev, x_cpu, _ = both(sc, steps.point_source_steps(1024, 200, pos=src, seed=14), stop_detected_photons=False)
    assert counted > 10
    identical(got, want, counted, x_dev, x_cpu)
