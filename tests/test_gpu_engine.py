"""The drop-in boundary on a real device: converter life cycle, error behaviour, threading
contract (any-order results to any caller, I3CLSimServer's five workers), statistics keys,
double buffering.  Written against the same behaviours the reference's interface documents
(public/clsim/I3CLSimStepToPhotonConverter.h, …ConverterOpenCL.cxx:1324-1640,
resources/tests/testCLSimServer.py)."""
import threading

import numpy as np
import pytest

from clsim_b200 import capi, steps
from clsim_b200.converter import (I3CLSimStepToPhotonConverter_exception, I3CLSimStepToPhotonConverterCUDA, configureCUDADevices,
                                  initializeCUDA)
from clsim_b200.description import KERNEL_FAST, KERNEL_REFERENCE, STEP_DTYPE
from clsim_b200.sharding import merge_results
from tests.scenes import HOSTCHECK, make_scene, sized

pytestmark = pytest.mark.gpu


def make_converter(sc, **kw):
    if HOSTCHECK:
        kw.setdefault("kernelMode", KERNEL_REFERENCE)   # (the host check build holds no fast kernel, tests/scenes.py)
    dev = configureCUDADevices(UseGPUs=True, UseOnlyDeviceNumber=0, OverrideApproximateNumberOfWorkItems=kw.pop("work_items", 8192))[0]
    return initializeCUDA(dev, 1, sc.geo, sc.medium, sc.bias, sc.generators, stopDetectedPhotons=True, pancakeFactor=sc.pancake, **kw)


def test_life_cycle_and_errors():
    sc = make_scene("spice_mie", geo_kind="ring")
    conv = I3CLSimStepToPhotonConverterCUDA(1, useNativeMath=not HOSTCHECK)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="not initialized"):
        conv.EnqueueSteps(steps.muon_track_steps(8), 0)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="WlenGenerators not set"):
        conv.Initialize()
    conv.SetWlenGenerators(sc.generators)
    conv.SetWlenBias(sc.bias)
    conv.SetMediumProperties(sc.medium)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="Geometry not set"):
        conv.Compile()
    conv.SetGeometry(sc.geo)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="Device not selected"):
        conv.Compile()
    conv.SetDevice(0)
    conv.SetStopDetectedPhotons(True)
    conv.SetDOMPancakeFactor(5.0)
    conv.SetWorkgroupSize(64)
    conv.SetMaxNumWorkitems(1024)
    assert not conv.IsInitialized()
    conv.Initialize()
    assert conv.IsInitialized()
    assert conv.GetWorkgroupSize() == 64 and conv.GetMaxNumWorkitems() == 1024
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="already initialized"):
        conv.Initialize()
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="already initialized"):
        conv.SetStopDetectedPhotons(False)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="Steps are empty"):
        conv.EnqueueSteps(np.zeros(0, dtype=STEP_DTYPE), 0)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="greater than maximum"):
        conv.EnqueueSteps(steps.muon_track_steps(2048), 0)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception, match="multiple of the workgroup size"):
        conv.EnqueueSteps(steps.muon_track_steps(100), 0)
    assert conv.QueueSize() == 0 and not conv.MorePhotonsAvailable()
    conv.EnqueueSteps(steps.pad_to_granularity(steps.muon_track_steps(100), 64), 42)
    res = conv.GetConversionResult()
    assert res.identifier == 42
    assert res.photons is not None and res.photonHistories is None   # photons never NULL (I3CLSimClientModule.cxx:589)
    st = conv.GetStatistics()
    assert sorted(st) == sorted(["TotalDeviceTime", "TotalHostTime", "NumKernelCalls", "TotalNumPhotonsGenerated", "TotalNumPhotonsAtDOMs",
                                 "AverageDeviceTimePerPhoton", "AverageHostTimePerPhoton", "DeviceUtilization"])
    assert st["NumKernelCalls"] == 1 and st["TotalNumPhotonsGenerated"] == 100 * 200
    assert 0 < st["DeviceUtilization"] <= 1.0 + 1e-9
    conv.Close()
    with pytest.raises(RuntimeError):
        configureCUDADevices(UseGPUs=False, UseCPUs=True)
    with pytest.raises(I3CLSimStepToPhotonConverter_exception):
        I3CLSimStepToPhotonConverterCUDA(1).SetDoublePrecision(True)


def test_bunch_with_no_hits_returns_empty_not_null():
    sc = make_scene("spice_mie", geo_kind="ring")
    conv = make_converter(sc)
    far = steps.point_source_steps(64, 10, pos=(5000.0, 5000.0, 0.0), seed=3)  # 5 km away from the 24 DOMs
    conv.EnqueueSteps(far, 9)
    res = conv.GetConversionResult()
    assert res.identifier == 9 and res.photons is not None and len(res.photons) == 0
    conv.Close()


@pytest.mark.parametrize("double_buffering", [False, True])
def test_five_workers_any_order_results(double_buffering):
    """I3CLSimServer runs queueDepth=5 worker threads per converter, each doing EnqueueSteps then
    GetConversionResult and matching by identifier (I3CLSimServer.cxx:126-135, 310-343)."""
    sc = make_scene("homogeneous")
    conv = make_converter(sc, enableDoubleBuffering=double_buffering, work_items=4096)
    n_workers, per_worker = 5, 6
    got, errors = [], []
    lock = threading.Lock()

    def worker(w):
        try:
            for k in range(per_worker):
                ident = w * 100 + k
                bunch = steps.point_source_steps(512 + 64 * w, 50, seed=ident)
                bunch["identifier"] = ident
                conv.EnqueueSteps(bunch, ident)
                res = conv.GetConversionResult()   # may be some other worker's bunch
                with lock:
                    got.append((res.identifier, res.photons))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(w,)) for w in range(n_workers)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errors and all(not t.is_alive() for t in ts)
    assert sorted(i for i, _ in got) == sorted(w * 100 + k for w in range(n_workers) for k in range(per_worker))
    for ident, photons in got:
        assert np.all(photons["identifier"] == ident)       # results are never mixed between bunches
    merged = merge_results(got)
    assert len(merged) == n_workers * per_worker
    st = conv.GetStatistics()
    assert st["NumKernelCalls"] == n_workers * per_worker
    assert st["TotalNumPhotonsAtDOMs"] == sum(len(p) for _, p in got)
    conv.Close()


def test_destroy_with_pending_work_does_not_hang():
    sc = make_scene("spice_mie", geo_kind="ring")
    conv = make_converter(sc, work_items=sized(1 << 16, 1 << 12))
    for i in range(4):
        conv.EnqueueSteps(steps.muon_track_steps(sized(1 << 15, 1 << 11), seed=i), i)
    done = threading.Event()

    def closer():
        conv.Close()
        done.set()

    t = threading.Thread(target=closer)
    t.start()
    t.join(timeout=120)
    assert done.is_set()


def test_two_converters_disjoint_rng_slices_give_independent_results():
    sc = make_scene("homogeneous")
    dev = configureCUDADevices(UseOnlyDeviceNumber=0, OverrideApproximateNumberOfWorkItems=8192)[0]
    bunch = steps.point_source_steps(8192, 200, seed=5)  # ~440 hits expected
    outs = []
    for row in (0, 163840, 0):
        conv = initializeCUDA(dev, 17, sc.geo, sc.medium, sc.bias, sc.generators, pancakeFactor=5.0, rngFirstMultiplierRow=row)
        conv.EnqueueSteps(bunch, 1)
        outs.append(conv.GetConversionResult().photons)
        conv.Close()
    n = [len(o) for o in outs]
    assert min(n) > 300
    # same seed, different multiplier slice: different photons, same physics
    assert abs(n[0] - n[1]) < 6 * np.sqrt(n[0])
    assert not np.array_equal(np.sort(outs[0]["wavelength"])[:50], np.sort(outs[1]["wavelength"])[:50])


def test_reference_mode_converter_history():
    sc = make_scene("spice_mie", geo_kind="ring")
    conv = make_converter(sc, kernelMode=KERNEL_REFERENCE, photonHistoryEntries=4, work_items=2048)
    src = (sc.geo.posX[0] + 8.0, sc.geo.posY[0], sc.geo.posZ[0])
    conv.EnqueueSteps(steps.point_source_steps(2048, 100, pos=src, seed=8), 3)
    res = conv.GetConversionResult()
    assert len(res.photons) > 50
    assert res.photonHistories.shape == (len(res.photons), 4, 4)
    scat = res.photons["num_scatters"]
    filled = (~np.isnan(res.photonHistories[:, :, 0])).sum(1)
    assert np.array_equal(filled, np.minimum(scat, 4))
    # history rows are in forward order: absorption-length column increases
    multi = np.where(scat >= 2)[0]
    assert len(multi) > 0
    for i in multi[:200]:
        col = res.photonHistories[i, :min(int(scat[i]), 4), 3]
        assert np.all(np.diff(col) >= 0)
    conv.Close()


def test_fast_kernel_converter_history_and_its_limit():
    # the same option on the fast kernel (the default of the converter class)
    sc = make_scene("spice_mie", geo_kind="ring")
    src = (sc.geo.posX[0] + 8.0, sc.geo.posY[0], sc.geo.posZ[0])
    conv = make_converter(sc, kernelMode=KERNEL_FAST, photonHistoryEntries=4, work_items=2048)
    conv.EnqueueSteps(steps.point_source_steps(2048, 100, pos=src, seed=9), 4)
    res = conv.GetConversionResult()
    assert res.photonHistories is not None and res.photonHistories.shape == (len(res.photons), 4, 4)
    filled = (~np.isnan(res.photonHistories[:, :, 0])).sum(1)
    assert np.array_equal(filled, np.minimum(res.photons["num_scatters"], 4)) and len(res.photons) > 50
    conv.Close()
    # more history than the device keeps is refused at creation, with a message (C ABI)
    opt = sc.options(kernel_mode=KERNEL_FAST, photon_history_entries=33, max_num_workitems=2048)
    with pytest.raises(capi.ClsimCudaError, match="at most 32"):
        capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, opt)
