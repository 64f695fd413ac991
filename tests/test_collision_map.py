"""The xy collision map of the fast kernel (clsim_b200/csrc/tables.cpp build_collision_map; device_scene.h
DevGeometry::near_info), checked on the CPU for the guarantee the kernel's pruning rests on:

    a photon anywhere in a pixel (or outside the map, clamped to a border pixel) that flies less than the pixel's RANGE
    cannot come within string_max_radius of any string other than the one the pixel names.

The named string itself is always tested exactly by the kernel (2-D segment / cylinder test), legs are cut at the
range, and pixels without a range (strings denser than the pixels are wide) send every leg to the reference's cell
walk.  The reference has no such map: it walks its cell grids for every segment (sparse_collision_kernel.c.cl:194-460)."""
import numpy as np
import pytest

from clsim_b200 import capi
from clsim_b200.description import SimpleGeometry
from tests.scenes import make_scene


def decode(m):
    info = np.asarray(m["info"], dtype=np.uint32).reshape(m["ny"], m["nx"])
    who = (info & np.uint32(0xffff)) >> np.uint32(4)
    bits = info & np.uint32(0xffff0000)
    rng = bits.view(np.float32)
    return who.astype(int), rng, bits


def pixel_of(m, x, y):
    """The kernel's pixel assignment (plan_leg): float32 fma, conversion toward zero saturating at 0, clamp above."""
    fx = np.float32(x) * np.float32(m["inv_pixel"]) + np.float32(m["off_x"])
    fy = np.float32(y) * np.float32(m["inv_pixel"]) + np.float32(m["off_y"])
    px = np.clip(np.trunc(np.maximum(fx, 0)).astype(np.int64), 0, m["nx"] - 1)
    py = np.clip(np.trunc(np.maximum(fy, 0)).astype(np.int64), 0, m["ny"] - 1)
    return px, py


def check_guarantee(m, n_points, seed, outside=0.0):
    who, rng, bits = decode(m)
    sx, sy = np.asarray(m["string_pos_x"]), np.asarray(m["string_pos_y"])
    R = m["string_max_radius"]
    r = np.random.default_rng(seed)
    x = r.uniform(m["x0"] - outside, m["x0"] + m["nx"] * m["pixel"] + outside, n_points)
    y = r.uniform(m["y0"] - outside, m["y0"] + m["ny"] * m["pixel"] + outside, n_points)
    px, py = pixel_of(m, x, y)
    w, rg, b = who[py, px], rng[py, px].astype(np.float64), bits[py, px]
    walk = b == np.uint32(0x7f800000)
    assert np.all(w[walk] == m["num_strings"])                       # the NaN record behind the last string
    assert np.all(w[~walk] < m["num_strings"]) and np.all(rg[~walk] >= 1.0)
    d = np.hypot(x[:, None] - sx[None, :], y[:, None] - sy[None, :])  # to every string axis
    d_other = d.copy()
    d_other[np.arange(n_points)[~walk], w[~walk]] = np.inf
    slack = d_other.min(axis=1) - R - rg
    assert np.all(slack[~walk] >= 0.0), slack[~walk].min()
    return walk.mean(), float(np.median(rg[~walk])) if (~walk).any() else 0.0


@pytest.mark.parametrize("budget", [40000, 11000, 2000, 512])
def test_ic86_like_detector(budget):
    sc = make_scene("spice_mie")
    m = capi.describe_collision_map(sc.medium, sc.geo, sc.generators, sc.bias, sc.options(), pixel_budget=budget)
    assert m["nx"] * m["ny"] <= budget and m["num_strings"] == 86
    # every string lies inside the map, one pixel away from its border at least
    sx, sy = np.asarray(m["string_pos_x"]), np.asarray(m["string_pos_y"])
    assert sx.min() >= m["x0"] + 0.99 * m["pixel"] and sx.max() <= m["x0"] + (m["nx"] - 0.99) * m["pixel"]
    assert sy.min() >= m["y0"] + 0.99 * m["pixel"] and sy.max() <= m["y0"] + (m["ny"] - 0.99) * m["pixel"]
    walk_fraction, median_range = check_guarantee(m, 200000, seed=budget)
    check_guarantee(m, 50000, seed=budget + 1, outside=400.0)           # points outside the map clamp to the border pixels
    if budget >= 11000:
        assert walk_fraction == 0.0 and median_range > 40.0             # 125 m grid: the range never binds a typical leg
    who, rng, _ = decode(m)
    # the named string is the nearest one to the pixel centre
    cx = m["x0"] + (np.arange(m["nx"]) + 0.5) * m["pixel"]
    cy = m["y0"] + (np.arange(m["ny"]) + 0.5) * m["pixel"]
    d = np.hypot(cx[None, :, None] - sx[None, None, :], cy[:, None, None] - sy[None, None, :])
    named = who < m["num_strings"]
    assert np.all(d.argmin(axis=2)[named] == who[named])


def test_legs_flown_on_clearance_stay_clear_of_every_string():
    """kernel_fast.cu advance_photon: the leg that looks at the map leaves the lane a clearance,
    min(distance to the named string's axis - R, range) - 1 cm; the legs after it fly without map or test as long as
    their summed length stays below it.  Restated in numpy: random walks started anywhere, three legs per look."""
    sc = make_scene("spice_mie")
    m = capi.describe_collision_map(sc.medium, sc.geo, sc.generators, sc.bias, sc.options(), pixel_budget=11000)
    who, rng, bits = decode(m)
    sx, sy = np.asarray(m["string_pos_x"]), np.asarray(m["string_pos_y"])
    R = m["string_max_radius"]
    r = np.random.default_rng(11)
    n = 100000
    x = r.uniform(m["x0"], m["x0"] + m["nx"] * m["pixel"], n)
    y = r.uniform(m["y0"], m["y0"] + m["ny"] * m["pixel"], n)
    px, py = pixel_of(m, x, y)
    w, rg = who[py, px], rng[py, px].astype(np.float64)
    assert np.all(bits[py, px] != np.uint32(0x7f800000))
    clearance = np.minimum(np.hypot(sx[w] - x, sy[w] - y) - R, rg) - 0.01
    flown = np.zeros(n)
    flew = 0
    for leg in range(3):
        length = r.exponential(4.0, n)                       # metres; clear ice has legs this long
        phi = r.uniform(0, 2 * np.pi, n)
        sin_theta = np.sqrt(1 - r.uniform(-1, 1, n) ** 2)    # the xy projection of an isotropic direction
        ok = length < clearance - flown if leg > 0 else np.ones(n, dtype=bool)   # the first leg is tested by the kernel itself
        if leg > 0:
            # closest approach of the leg's xy projection to every string axis
            dx, dy = np.cos(phi) * sin_theta, np.sin(phi) * sin_theta
            t = np.clip(((sx[None, :] - x[:, None]) * dx[:, None] + (sy[None, :] - y[:, None]) * dy[:, None])
                        / np.maximum(dx * dx + dy * dy, 1e-30)[:, None], 0.0, length[:, None])
            dist = np.hypot(x[:, None] + t * dx[:, None] - sx[None, :], y[:, None] + t * dy[:, None] - sy[None, :])
            assert np.all(dist[ok].min(axis=1) > R)
            flew += int(ok.sum())
        else:
            dx, dy = np.cos(phi) * sin_theta, np.sin(phi) * sin_theta
        step = np.where(ok, length, 0.0)                      # a lane without clearance waits
        x, y, flown = x + step * dx, y + step * dy, flown + step
    assert flew > 1.5 * n                                     # most legs do fly on clearance


def test_dense_cluster_has_pixels_without_a_range():
    sc = make_scene("homogeneous", geo_kind="ring")
    ring = sc.geo
    sid, did, xs, ys, zs = list(ring.stringIDs), list(ring.domIDs), list(ring.posX), list(ring.posY), list(ring.posZ)
    sub = ["Ring"] * len(sid)
    for k, (cx, cy) in enumerate(((0.0, 0.0), (6.0, 0.5), (-0.5, 7.0))):
        for d in range(6):
            sid.append(20 + k); did.append(d + 1); xs.append(cx + 0.1 * d); ys.append(cy - 0.05 * d); zs.append(25.0 - 10.0 * d)
            sub.append("Cluster")
    geo = SimpleGeometry(sid, did, xs, ys, zs, ring.OMRadius, subdetectors=sub)
    m = capi.describe_collision_map(sc.medium, geo, sc.generators, sc.bias, sc.options(), pixel_budget=512)
    walk_fraction, _ = check_guarantee(m, 100000, seed=3)
    assert walk_fraction > 0.0                                            # the cluster is denser than the 12 m pixels
    who, rng, bits = decode(m)
    px, py = pixel_of(m, np.array([3.0]), np.array([3.0]))
    assert bits[py[0], px[0]] == np.uint32(0x7f800000)                    # where test_dense_strings_take_the_cell_walk puts its source
    # with the full budget the pixels are fine enough again
    m2 = capi.describe_collision_map(sc.medium, geo, sc.generators, sc.bias, sc.options(), pixel_budget=40000)
    check_guarantee(m2, 100000, seed=4)


def test_more_strings_than_the_map_can_name_are_refused_by_the_fast_kernel_only():
    # the map itself is host data; the 4094-string limit is the fast kernel's (fast_kernel_supports) and needs a device
    sc = make_scene("spice_mie", geo_kind="ring")
    m = capi.describe_collision_map(sc.medium, sc.geo, sc.generators, sc.bias, sc.options(), pixel_budget=4096)
    assert m["num_strings"] == 8 and max(m["info"]) & 0xffff <= 8 << 4
