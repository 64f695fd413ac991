"""The C++ host class I3CLSimStepToPhotonConverterCUDA (clsim_b200/host/), the drop-in for the reference's
I3CLSimStepToPhotonConverterOpenCL (public/clsim/I3CLSimStepToPhotonConverterOpenCL.h:78-258).

CPU: the contract checks of test_converter_cuda (error behaviour and messages of …OpenCL.cxx:492-508,
1324-1544; Initialize must THROW without a device -- no CPU fallback) and that the C++ flattening of the
polymorphic description objects produces byte-for-byte the device tables the Python flattening produces
for the same model.  GPU: the same binary drives a real device with 5 producer threads."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from clsim_b200 import capi, geometry
from clsim_b200.description import ConverterOptions, MediumProperties, WlenBias, WlenGenerator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "clsim_b200", "host", "build", "test_converter_cuda")


@pytest.fixture(scope="module")
def binary():
    entry.build_product()
    entry.build_host_class()
    assert os.path.isfile(BINARY)
    return BINARY


def python_twin_of_cxx_model(aniso):
    """The model test_converter_cuda.cxx builds in make_medium / make_bias / make_generator / make_ring_geometry."""
    nm, deg = 1e-9, math.pi / 180.0
    m = MediumProperties()
    m.layersNum, m.layersZStart, m.layersHeight = 12, -60.0, 10.0
    m.kappa, m.A, m.B, m.alpha = 1.08410680294, 6954.09033203, 6617.75439453, 0.898608505726
    m.D, m.E = 400.0 ** m.kappa, 0.0
    layer = np.arange(12)
    m.b400 = (0.020 + 0.004 * np.sin(0.9 * layer)) / (1.0 - 0.9)
    m.aDust400 = 0.006 + 0.002 * np.cos(0.7 * layer)
    m.deltaTau = 5.0 + 0.3 * layer
    m.scat_kind, m.fractionOfFirstDistribution, m.meanCosine = 0, 0.45, 0.9
    if aniso:
        dist = np.array([-500.0, -100.0, 0.0, 150.0, 400.0])
        zc = -70.0 + 7.5 * np.arange(20)
        corr = 0.004 * dist[:, None] * np.cos(0.2 * np.arange(20))[None, :]
        m.tilt = {"distancesFromOriginAlongTilt": dist, "zCoordinates": zc, "zCorrections": corr, "directionOfTiltAzimuth": 225.0 * deg}
        m.anisotropy = {"anisotropyDirAzimuth": 216.0 * deg, "magnitudeAlongDir": 0.04, "magnitudePerpToDir": -0.08}
        k1, k2 = math.exp(0.04), math.exp(-0.08)
        ca, sa = math.cos(216.0 * deg), math.sin(216.0 * deg)
        T = np.array([[ca, sa, 0], [-sa, ca, 0], [0, 0, 1.0]])
        m.preMatrix = T.T @ np.diag([k1, k2, 1.0 / (k1 * k2)]) @ T
        m.postMatrix = T.T @ np.diag([1 / k1, 1 / k2, k1 * k2]) @ T
    i = np.arange(43)
    bias = WlenBias(values=0.02 + 0.11 * np.exp(-0.5 * ((i - 14) / 7.0) ** 2), start_wlen=260 * nm, wlen_step=10 * nm)
    w = 260 * nm + 10 * nm * i
    n = np.array([m.GetPhaseRefractiveIndex(x) for x in w])
    gen = WlenGenerator.interpolated(260 * nm, 10 * nm, bias.values * (2.0 * math.pi / 137.0) / (w * w) * (1.0 - 1.0 / (n * n)))
    geo = geometry.make_ring_geometry(oversize=5.0)
    opt = ConverterOptions(stop_detected_photons=True, pancake_factor=5.0, rng_seed=12345)
    return m, geo, [gen], bias, opt


def test_contract_without_a_device(binary, has_gpu):
    if has_gpu:
        pytest.skip("the no-device contract (Initialize must throw) is checked on machines without a GPU")
    res = subprocess.run([binary, "--no-gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    print(res.stdout)
    assert res.returncode == 0, res.stdout
    assert "no CPU fallback" in res.stdout
    assert "0 failed" in res.stdout


@pytest.mark.parametrize("aniso", [False, True])
def test_cxx_flattening_equals_python_flattening(binary, aniso):
    res = subprocess.run([binary, "--describe"] + (["aniso"] if aniso else []), stdout=subprocess.PIPE, text=True, timeout=120)
    assert res.returncode == 0
    got = json.loads(res.stdout)
    m, geo, gens, bias, opt = python_twin_of_cxx_model(aniso)
    want = capi.describe_tables(m, geo, gens, bias, opt)
    assert sorted(got.keys()) == sorted(want.keys())

    def same(a, b, path):
        if isinstance(a, dict):
            assert sorted(a) == sorted(b), path
            for k in a:
                same(a[k], b[k], path + "/" + k)
        elif isinstance(a, list):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                same(x, y, "%s[%d]" % (path, i))
        elif isinstance(a, float) or isinstance(b, float):
            # the same doubles go through the same float conversion; libm pow/exp/sin in the two model builders
            # may differ in the last bit of the double
            assert a == pytest.approx(b, rel=2e-7, abs=1e-30), path
        else:
            assert a == b, path
    same(got, want, "")


@pytest.mark.gpu
def test_cxx_converter_on_the_device(binary):
    res = subprocess.run([binary, "--gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(res.stdout)
    assert res.returncode == 0, res.stdout
    assert "0 failed" in res.stdout


SERVER_BINARY = os.path.join(ROOT, "clsim_b200", "host", "build", "test_server_inprocess")


def test_server_seam_with_dummy_converters(binary):
    """resources/tests/testCLSimServer.py restated in C++ (DummyConverter, three clients, results matched by
    identifier), plus the bunch-size harmonisation of I3CLSimServer.cxx:95-113.  No device needed."""
    res = subprocess.run([SERVER_BINARY], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    print(res.stdout)
    assert res.returncode == 0, res.stdout
    assert "0 failed" in res.stdout


@pytest.mark.gpu
def test_server_seam_on_devices(binary):
    """Two CUDA converters behind one in-process server (one per device; on a 1-GPU box both on device 0), three
    clients; photon conservation through the summed statistics."""
    res = subprocess.run([SERVER_BINARY, "--gpu", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(res.stdout)
    assert res.returncode == 0, res.stdout
    assert "0 failed" in res.stdout


NEIGHBOURS_BINARY = os.path.join(ROOT, "clsim_b200", "host", "build", "test_neighbours_cuda")


def test_cxx_neighbour_classes_contract(binary, has_gpu):
    """I3CLSimPhotonToMCPEConverterCUDA / I3CLSimStepGeneratorCUDA (C++): argument errors with the reference's messages,
    and no construction without a device."""
    res = subprocess.run([NEIGHBOURS_BINARY, "--gpu-args-only" if has_gpu else "--no-gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=120)
    print(res.stdout)
    assert res.returncode == 0, res.stdout
    assert "0 failed" in res.stdout


@pytest.mark.gpu
def test_cxx_neighbour_classes_on_the_device(binary):
    """Photon -> MCPE against a C++ restatement of the reference's Convert (explicit uniforms: every survivor), the
    conversion attached to a converter, steps made on the device, bunches generated in place."""
    res = subprocess.run([NEIGHBOURS_BINARY, "--gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(res.stdout)
    assert res.returncode == 0, res.stdout
    assert "0 failed" in res.stdout
