"""The re-formulations the fast kernel relies on, restated in numpy and checked against the reference formulation they
replace (CPU only; the device code itself is checked on the GPU in test_gpu_fast_kernel.py):

* the SL + HG scattering mix from five constants folded on the host (engine.cu upload_tables, DevMedium::sl_off ...),
  with the 2^-32 scale of the draw folded in  ==  I3CLSimRandomValueMixed / SimplifiedLiu / HenyeyGreenstein
  (I3CLSimRandomValueMixed.cxx:117-146 and the two samplers);
* the rotation with ONE reciprocal root (kernel_fast.cu rotate_packed)  ==  scatterDirectionByAngle
  (propagation_kernel.c.cl:83-129);
* the bin of a wavelength draw found from a guide table and a forward scan (kernel_fast.cu draw_wavelength)  ==  the
  reference's linear scan (I3CLSimRandomValueInterpolatedDistribution.cxx:236-337)."""
import numpy as np
import pytest

from tests.scenes import make_scene


@pytest.fixture(scope="module")
def mie():
    return make_scene("spice_mie")


def test_folded_scattering_mix_equals_the_reference_samplers(mie):
    f = float(np.float32(mie.medium.fractionOfFirstDistribution))
    g = float(np.float32(mie.medium.meanCosine))
    beta = float(np.float32((1.0 - g) / (1.0 + g)))
    omf = float(np.float32(1.0 - f))
    # engine.cu: computed in double, rounded to float
    sl_off = np.float32(beta * (np.log2(1.0 / f) - 32.0) + 1.0)
    h0 = np.float32(1.0 + g * (2.0 / omf - 1.0))
    h1 = np.float32(-2.0 * g / omf / 4294967296.0)
    hc = np.float32((1.0 + g * g) / (2.0 * g))
    hw = np.float32((1.0 - g * g) ** 2 / (2.0 * g))
    split = np.float32(f) * np.float32(4294967296.0)
    rng = np.random.default_rng(5)
    draws = np.concatenate([rng.integers(1, 2 ** 32, 200000, dtype=np.uint64), [1, 2 ** 32 - 1, int(f * 2 ** 32) - 1, int(f * 2 ** 32) + 1]])
    # the device converts toward zero
    U = np.floor(draws.astype(np.float64)).astype(np.float32)
    U = np.where(U.astype(np.float64) > draws.astype(np.float64), np.nextafter(U, np.float32(0)), U)
    # fast kernel, float32 throughout
    cos_sl = np.exp2(np.float32(beta) * np.log2(U) + sl_off) - np.float32(1)
    r = np.float32(1) / (h1 * U + h0)
    cos_hg = hc - hw * (r * r)
    fast = np.where(U < split, cos_sl, cos_hg).astype(np.float64)
    # reference, double: one draw rr; rr < f -> SL with rr / f, else HG with (1 - rr) / (1 - f)
    rr = U.astype(np.float64) / 4294967296.0
    sl = 2.0 * (rr / f) ** beta - 1.0
    s = 2.0 * ((1.0 - rr) / omf) - 1.0
    ii = (1.0 - g * g) / (1.0 + g * s)
    hg = (1.0 + g * g - ii * ii) / (2.0 * g)
    ref = np.where(rr < f, sl, hg)
    assert np.all((U < split) == (rr < f))
    assert np.abs(fast - ref).max() < 2e-5            # float32 evaluation of exponents near 32 * beta, see DESIGN 2.1
    assert np.abs(fast - ref).mean() < 2e-7
    assert fast.min() >= -1.0 - 1e-6 and fast.max() <= 1.0 + 1e-6


def test_rotation_with_one_reciprocal_root_equals_the_reference_rotation():
    rng = np.random.default_rng(6)
    n = 100000
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    cosa = rng.uniform(-1, 1, n)
    sina = np.sqrt(1 - cosa ** 2)
    b = rng.uniform(0, 2 * np.pi, n)
    sinb, cosb = np.sin(b), np.cos(b)
    # reference (propagation_kernel.c.cl:83-129), general branch
    sinth = np.sqrt(1 - d[:, 2] ** 2)
    costh = d[:, 2]
    sinph, cosph = d[:, 1] / sinth, d[:, 0] / sinth
    # rotate (sina cosb, sina sinb, cosa) from the frame of d into the lab frame
    ref = np.stack([cosph * costh * sina * cosb - sinph * sina * sinb + d[:, 0] * cosa,
                    sinph * costh * sina * cosb + cosph * sina * sinb + d[:, 1] * cosa,
                    -sinth * sina * cosb + costh * cosa], axis=1)
    # fast kernel: sin^2(theta) from x and y, k and m from one reciprocal root
    s2 = d[:, 0] ** 2 + d[:, 1] ** 2
    sina2 = 1 - cosa ** 2
    x = sina2 * s2
    r = 1 / np.sqrt(np.maximum(x, 1e-36))
    k, m = sina2 * r, x * r
    u = cosb * k
    w = cosa - d[:, 2] * sinb * k
    fast = np.stack([d[:, 0] * w - d[:, 1] * u, d[:, 1] * w + d[:, 0] * u, d[:, 2] * cosa + m * sinb], axis=1)
    # the two parametrise the azimuth differently (the reference measures it from the meridian plane, the kernel's
    # formula from the perpendicular): same cone, azimuth shifted by a quarter turn -> compare the invariants ...
    assert np.abs(np.linalg.norm(fast, axis=1) - 1).max() < 1e-12
    assert np.abs((fast * d).sum(1) - cosa).max() < 1e-12          # scattering angle
    # ... and the azimuth itself after undoing the quarter turn: fast(b) == ref(b') with (cos b', sin b') = (-sin b, cos b)
    sinb2, cosb2 = cosb, -sinb
    ref2 = np.stack([cosph * costh * sina * cosb2 - sinph * sina * sinb2 + d[:, 0] * cosa,
                     sinph * costh * sina * cosb2 + cosph * sina * sinb2 + d[:, 1] * cosa,
                     -sinth * sina * cosb2 + costh * cosa], axis=1)
    assert min(np.abs(fast - ref).max(), np.abs(fast - ref2).max()) < 1e-9


def test_guide_table_and_forward_scan_find_the_bin_of_the_linear_scan(mie):
    gen = mie.generators[0]
    y = np.asarray(gen.y, dtype=np.float64)
    integral = np.concatenate([[0.0], np.cumsum(gen.dx * (y[1:] + y[:-1]) / 2.0)])
    cum = (integral / integral[-1]).astype(np.float32)          # tables.cpp make_generator
    n = len(cum)
    cells = 64
    def linear_scan(r, start=0):
        k = start
        while k < n - 2 and cum[k + 1] < r:
            k += 1
        return k
    guide = np.array([linear_scan(np.float32(c) / np.float32(cells)) for c in range(cells)], dtype=np.uint8)
    assert np.all(np.diff(guide.astype(int)) >= 0)
    rng = np.random.default_rng(7)
    u = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64)
    rs = np.concatenate([(np.float32(1) - (u.astype(np.float32) * np.float32(2.3283064365386963e-10))).astype(np.float32),
                         cum, np.nextafter(cum, np.float32(2)), np.float32([1.0, 1e-9])])
    rs = rs[(rs > 0) & (rs <= 1)]
    for r in rs:
        c = min(int(np.float32(r) * np.float32(cells)), cells - 1)
        assert linear_scan(r, int(guide[c])) == linear_scan(r) == int(np.searchsorted(cum[1:n - 1], r, side="left"))
