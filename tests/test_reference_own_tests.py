"""The reference's own tests of the generated functions, run for real.

resources/tests/{testScalarFields, testVectorTransforms, testSpiceLeaTransforms, testScalarFieldIceTiltZShift}.py are the
reference's known-answer tests for the path's medium functions: each evaluates a description object three ways -- the OpenCL
text it generates (through an OpenCL device), its C++ host method ("reference function"), and a formula written out in Python
"stolen from ppc" -- on 1e5 (1e4) random arguments and demands agreement within a stated tolerance.  They cannot run as
shipped (no OpenCL device, no IceTray).  Here the same three evaluations are made with the reference's own code compiled for
the host (oracle/_ref/libclsim_ref_medium.so: the classes; their generated text compiled under the OpenCL-C shim,
pyoracle.RefProgram), the same numbers of trials, the same argument distributions, the same tolerances -- and a fourth column,
the ORACLE's restatement (what the GPU parity tests check against), is held to the same bar.  The Python formulas below are
the reference scripts' (ppc's), retyped for numpy arrays.

CPU only; needs /root/reference -- skipped elsewhere."""
import math

import numpy as np
import pytest

from clsim_b200 import ice
from oracle import pyoracle
from tests.scenes import make_scene

pytestmark = pytest.mark.skipif(not pyoracle.ref_program_available(), reason="needs /root/reference and oracle/_ref (this container)")

DEG = math.pi / 180.0


@pytest.fixture(scope="module")
def lea():
    sc = make_scene("spice_lea")
    opt = sc.options(max_num_workitems=1024)
    prog = pyoracle.RefProgram(sc.medium, sc.geo, sc.generators, sc.bias, opt)
    ora = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, opt)
    return sc, prog, ora


def unit_vectors(n, seed):
    rng = np.random.default_rng(seed)
    zen = np.arccos(rng.uniform(0.0, 1.0, n) * 2.0 - 1.0)
    azi = rng.uniform(0.0, 2.0 * math.pi, n)
    return np.stack([np.sin(zen) * np.cos(azi), np.sin(zen) * np.sin(azi), np.cos(zen)], axis=1)


def test_scalar_fields_anisotropy_abs_len_scaling(lea):
    """testScalarFields.py: I3CLSimScalarFieldAnisotropyAbsLenScaling(216 deg, 0.04, -0.08), 100 000 directions, relative
    deviation of the C++ host method and of the generated text from the ppc formula <= 1e-5."""
    sc, prog, ora = lea
    an = sc.medium.anisotropy
    thx, logk1, logk2 = an["anisotropyDirAzimuth"], an["magnitudeAlongDir"], an["magnitudePerpToDir"]
    assert thx == pytest.approx(216.0 * DEG) and (logk1, logk2) == (0.04, -0.08)          # the script's parameters
    v = unit_vectors(100000, 1)
    azx, azy = math.cos(thx), math.sin(thx)
    k1, k2 = math.exp(logk1), math.exp(logk2)
    kz = 1.0 / (k1 * k2)
    n1, n2, n3 = azx * v[:, 0] + azy * v[:, 1], -azy * v[:, 0] + azx * v[:, 1], v[:, 2]
    s1, s2, s3, l1, l2, l3 = n1 * n1, n2 * n2, n3 * n3, k1 * k1, k2 * k2, kz * kz
    python = 1.0 / ((1 / l1 + 1 / l2 + 1 / l3 - (s1 / l1 + s2 / l2 + s3 / l3)) * (s1 * l1 + s2 * l2 + s3 * l3) / 2.0)
    host = pyoracle.ref_medium_host_values(prog.generated, 5, v)
    text = prog.eval_scalar_field(1, v)
    oracle = ora.eval_scalar_field(1, v)
    for name, got in (("C++ reference function", host), ("generated text", text), ("oracle", oracle)):
        assert np.abs((python - got) / python).max() <= 1e-5, name
    assert np.abs((python - host) / python).max() < 1e-13


def test_vector_transforms_matrix(lea):
    """testVectorTransforms.py: I3CLSimVectorTransformMatrix(matrix, renormalize), 100 000 directions, deviation of the C++
    host method and of the generated text from numpy.dot (+ renormalisation) <= 1e-4."""
    sc, prog, ora = lea
    v = unit_vectors(100000, 2)
    for which, matrix in ((0, sc.medium.preMatrix), (1, sc.medium.postMatrix)):
        m = np.asarray(matrix, dtype=float).reshape(3, 3)
        python = v @ m.T
        python /= np.linalg.norm(python, axis=1, keepdims=True)
        host = pyoracle.ref_medium_host_values(prog.generated, 8 + which, v)
        text = prog.eval_vector_transform(which, v)
        oracle = ora.eval_vector_transform(which, v)
        for name, got in (("C++ reference function", host), ("generated text", text), ("oracle", oracle)):
            assert np.abs(python - got).max() <= 1e-4, (which, name)
        assert np.abs(python - host).max() < 1e-14


def test_spice_lea_transforms_against_ppc():
    """testSpiceLeaTransforms.py: the matrices GetSpiceLeaAnisotropyTransforms makes (here: clsim_b200.ice's twin of the
    reference's Python), applied by the reference's I3CLSimVectorTransformMatrix::ApplyTransform, against ppc's pro.cu:621 / :636
    written out -- 10 000 directions, absolute deviation <= 1e-14."""
    thx, logk1, logk2 = 216.0, 0.04, -0.08
    _, pre, post = ice.GetSpiceLeaAnisotropyTransforms(anisotropyDirAzimuth=thx * DEG, magnitudeAlongDir=logk1, magnitudePerpToDir=logk2)
    medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=False)
    assert np.array_equal(np.asarray(medium.preMatrix, float).ravel(), np.asarray(pre, float).ravel())       # the medium carries exactly these
    assert np.array_equal(np.asarray(medium.postMatrix, float).ravel(), np.asarray(post, float).ravel())
    g = pyoracle.RefGeneratedSource(medium, [ice.makeCherenkovWavelengthGenerator(ice.GetIceCubeDOMAcceptance(), False, medium)],
                                    ice.GetIceCubeDOMAcceptance())
    v = unit_vectors(10000, 3)
    azx, azy = math.cos(thx * DEG), math.sin(thx * DEG)
    k1, k2 = math.exp(logk1), math.exp(logk2)
    kz = 1.0 / (k1 * k2)

    def ppc(vec, forward):
        n1 = (azx * vec[:, 0] + azy * vec[:, 1]) * (k1 if forward else 1.0 / k1)
        n2 = (-azy * vec[:, 0] + azx * vec[:, 1]) * (k2 if forward else 1.0 / k2)
        nx, ny = n1 * azx - n2 * azy, n1 * azy + n2 * azx
        nz = vec[:, 2] * (kz if forward else 1.0 / kz)
        r = 1.0 / np.sqrt(nx * nx + ny * ny + nz * nz)
        return np.stack([r * nx, r * ny, r * nz], axis=1)

    assert np.abs(pyoracle.ref_medium_host_values(g, 8, v) - ppc(v, True)).max() <= 1e-14
    assert np.abs(pyoracle.ref_medium_host_values(g, 9, v) - ppc(v, False)).max() <= 1e-14


def test_scalar_field_ice_tilt_z_shift(lea):
    """testScalarFieldIceTiltZShift.py: GetIceTiltZShift() of SpiceLea, 100 000 points in a (2.4 km)^3 cube, generated text
    against the C++ host method: deviation <= 10 cm."""
    sc, prog, ora = lea
    rng = np.random.default_rng(4)
    p = rng.uniform(-1200.0, 1200.0, (100000, 3))
    host = pyoracle.ref_medium_host_values(prog.generated, 4, p)
    text = prog.eval_scalar_field(0, p)
    oracle = ora.eval_scalar_field(0, p)
    assert np.abs(text - host).max() <= 0.10
    assert np.abs(oracle - host).max() <= 0.10
    assert np.abs(text - host).max() < 2e-2                # (what it is in fact: 1 cm, float against double on corrections of up to 90 m)
    assert np.abs(host).max() > 20.0
