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


# ---- round 2 re-formulations ----------------------------------------------------------------------------------------

def _reference_tilt_shift(t, x, y, z):
    """I3CLSimScalarFieldIceTiltZShift.cxx:145-216 in double."""
    dist, zc, corr = t["distancesFromOriginAlongTilt"], t["zCoordinates"], t["zCorrections"]
    az = t["directionOfTiltAzimuth"]
    lnx, lny = np.cos(az), np.sin(az)
    z0, dz = zc[0], zc[1] - zc[0]
    zr = (z - z0) / dz
    k = np.clip(np.floor(zr).astype(int), 0, len(zc) - 2)
    above = zr - k
    nr = lnx * x + lny * y
    out = np.zeros_like(x)
    for i in range(len(x)):
        for j in range(1, len(dist)):
            if nr[i] < dist[j] or j == len(dist) - 1:
                w_lo = (dist[j] - nr[i]) / (dist[j] - dist[j - 1])
                v_lo = corr[j - 1][k[i] + 1] * above[i] + corr[j - 1][k[i]] * (1 - above[i])
                v_hi = corr[j][k[i] + 1] * above[i] + corr[j][k[i]] * (1 - above[i])
                out[i] = v_hi * (1 - w_lo) + v_lo * w_lo
                break
    return out, nr


def test_tilt_interval_grid_and_cells_equal_the_reference_interpolation():
    """kernel_fast.cu tilt_shift: the column interval from a uniform grid over the distance along the tilt direction (cells of
    at most one node, widened by a margin; host part in engine.cu, cell records in the kernel's staging code) and the z
    interpolation of both columns from one float4 cell record == the reference's scan and bilinear formula."""
    t = make_scene("spice_lea").medium.tilt
    dist = np.asarray(t["distancesFromOriginAlongTilt"], dtype=np.float32)
    corr = np.asarray(t["zCorrections"], dtype=np.float32)
    nd, nz = corr.shape
    # host (engine.cu): cell width below the smallest gap between interior nodes
    gap = min(float(dist[i]) - float(dist[i - 1]) for i in range(2, nd - 1))
    width = gap * 0.99
    cells = int(np.ceil((float(dist[nd - 2]) - float(dist[1])) / width)) + 2
    assert cells <= 64
    scale, offset = np.float32(1.0 / width), np.float32(1.0 - float(dist[1]) / width)
    # staging: (nodes below the cell, the node inside it)
    w, d1, margin = 1.0 / float(scale), float(dist[1]), 1e-3 / float(scale)
    lut = []
    for c in range(cells):
        lo, hi = d1 + (c - 1) * w - margin, d1 + c * w + margin
        below, inside = 1, np.inf
        for i in range(1, nd - 1):
            d = float(dist[i])
            if c > 0 and d < lo:
                below += 1
            elif c == cells - 1 or d <= hi:
                inside = min(inside, d)
        lut.append((below, inside))
    rng = np.random.default_rng(11)
    nr = np.concatenate([rng.uniform(-800, 800, 200000), np.repeat(dist.astype(np.float64), 3) + np.tile([-1e-4, 0.0, 1e-4], nd),
                         d1 + w * np.arange(-2, cells + 2), d1 + w * np.arange(-2, cells + 2) + 1e-5]).astype(np.float32)
    cell = np.clip(np.trunc(nr * scale + offset).astype(int), 0, cells - 1)
    j_fast = np.array([lut[c][0] + (0 if v < lut[c][1] else 1) for c, v in zip(cell, nr)])
    j_ref = 1 + (nr[:, None] >= dist[None, 1:nd - 1]).sum(1)           # first j in [1, nd-1] with nr < dist[j], else nd-1
    assert np.array_equal(j_fast, j_ref)
    # the interpolation from cell records against the reference formula, on points across the table
    n = 4000
    x, y = rng.uniform(-600, 600, n), rng.uniform(-600, 600, n)
    z = rng.uniform(t["zCoordinates"][0] - 20, t["zCoordinates"][-1] + 20, n)
    want, nr64 = _reference_tilt_shift(t, x, y, z)
    zc = t["zCoordinates"]
    inv_dz = np.float32(1.0 / (zc[1] - zc[0]))
    zr = np.float32(z) * inv_dz + np.float32(-zc[0] / (zc[1] - zc[0]))
    k = np.clip(np.floor(zr).astype(int), 0, nz - 2)
    above = (zr - k).astype(np.float32)
    nr32 = nr64.astype(np.float32)
    j = 1 + (nr32[:, None] >= dist[None, 1:nd - 1]).sum(1)
    here_x, here_y = dist[j], (1.0 / (dist[j] - dist[j - 1])).astype(np.float32)
    w_lo = (here_x - nr32) * here_y
    lo0, hi0 = corr[j - 1, k], corr[j, k]
    v_lo = lo0 + (corr[j - 1, k + 1] - lo0) * above
    v_hi = hi0 + (corr[j, k + 1] - hi0) * above
    got = v_hi + w_lo * (v_lo - v_hi)
    assert np.abs(got - want).max() < 2e-3 and np.abs(got - want).mean() < 1e-4       # metres, on shifts of up to +-60 m; fp32 of z / dz


def test_block_transform_equals_the_full_matrix_and_the_reference_renormalisation():
    """kernel_fast.cu apply_block_matrix (five products) == I3CLSimVectorTransformMatrix.cxx:101-133 for the ppc matrices."""
    m = make_scene("spice_lea").medium
    rng = np.random.default_rng(12)
    d = rng.normal(size=(20000, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    for M in (np.asarray(m.preMatrix, dtype=np.float64), np.asarray(m.postMatrix, dtype=np.float64)):
        assert M[0, 2] == 0 and M[1, 2] == 0 and M[2, 0] == 0 and M[2, 1] == 0           # what launch_variant checks
        want = d @ M.T
        want /= np.linalg.norm(want, axis=1)[:, None]
        M32, d32 = M.astype(np.float32), d.astype(np.float32)
        nx = M32[0, 1] * d32[:, 1] + M32[0, 0] * d32[:, 0]
        ny = M32[1, 1] * d32[:, 1] + M32[1, 0] * d32[:, 0]
        nzv = M32[2, 2] * d32[:, 2]
        inv = 1.0 / np.sqrt(nx * nx + ny * ny + nzv * nzv)
        got = np.stack([nx * inv, ny * inv, nzv * inv], axis=1)
        assert np.abs(got - want).max() < 5e-7


def test_table_maker_fast_arc_cosine_and_step_parts():
    """tabulate_device.h tab_acos (Abramowitz & Stegun 4.4.46) against acos; fill_queue's split of a step into parts."""
    x = np.linspace(-1, 1, 400001)
    a = np.minimum(np.abs(x), 1.0)
    p = -0.0012624911
    for c in (0.0066700901, -0.0170881256, 0.0308918810, -0.0501743046, 0.0889789874, -0.2145988016, 1.5707963050):
        p = p * a + c
    r = np.sqrt(1 - a) * p
    got = np.where(x < 0, 3.14159265359 - r, r)
    assert np.abs(got - np.arccos(x)).max() < 1e-7            # the azimuth bins are degrees wide
    # part k of C covers photons [n k / C, n (k + 1) / C): the parts of a step add up to its photon count, none is negative
    for n in (0, 1, 7, 199, 200, 333, 65539, 2 ** 32 - 1):
        for C_ in (1, 2, 4, 8):
            parts = [n * (k + 1) // C_ - n * k // C_ for k in range(C_)]
            assert sum(parts) == n and min(parts) >= 0 and max(parts) - min(parts) <= 1
