"""Detector geometries for tests and benchmarks (flat DOM lists -> ``SimpleGeometry``).

* ``make_ring_geometry``: the 24-DOM ring the reference's benchmark builds in code for
  ``--minimal-gcd`` (resources/scripts/benchmark.py:63-114; same shape as
  compareToPPCredux/generateTestingGeometry.py).
* ``make_ic86_like_geometry``: a synthetic stand-in for the IC86 GCD file (which is not in
  the reference tree, benchmark.py:275): 78 strings on the 125 m triangular grid in the
  6-7-8-9-10-10-9-8-7-4 row pattern with 60 DOMs at 17 m spacing, plus 8 DeepCore-like
  strings (10 DOMs at 10 m above the dust layer, 50 DOMs at 7 m below).  Strings are
  slightly bent (< 0.4 m) so the int16 DOM-position quantisation of the geometry tables is
  exercised.  All DOMs are in one subdetector ("Unknown"), which is what the reference does
  when the frame has no Subdetectors object (I3CLSimSimpleGeometryFromI3Geometry.cxx:101-118).
* ``read_geometry_text_file``: rows ``string dom x y z`` (I3CLSimSimpleGeometryTextFile.cxx:58-100).
"""
import math

import numpy as np

from .description import SimpleGeometry

DOM_RADIUS = 0.16510  # metres (python/traysegments/I3CLSimMakePhotons.py default DOMRadius)


def make_ring_geometry(oversize=1.0, radius=120.0, center=(0.0, 0.0, 0.0)):
    dirs = np.array([[0, 1], [1, 1], [1, 0], [1, -1], [0, -1], [-1, -1], [-1, 0], [-1, 1]], dtype=float)
    dirs /= np.sqrt((dirs ** 2).sum(1))[:, None]
    sid, did, xs, ys, zs = [], [], [], [], []
    for s in range(8):
        for d, dz in enumerate((radius, 0.0, -radius)):
            sid.append(s + 1)
            did.append(d + 1)
            xs.append(center[0] + dirs[s, 0] * radius)
            ys.append(center[1] + dirs[s, 1] * radius)
            zs.append(center[2] + dz)
    return SimpleGeometry(sid, did, xs, ys, zs, DOM_RADIUS * oversize)


def make_ic86_like_geometry(oversize=5.0, bend=True):
    rows = [6, 7, 8, 9, 10, 10, 9, 8, 7, 4]
    spacing = 125.0
    row_dy = spacing * math.sin(math.radians(60.0))
    strings = []
    y0 = -0.5 * (len(rows) - 1) * row_dy
    for r, count in enumerate(rows):
        x_start = -0.5 * (count - 1) * spacing
        if count == 4:  # the short last row of the real detector sits on one side
            x_start = -0.5 * (7 - 1) * spacing - 0.5 * spacing + spacing
        for k in range(count):
            strings.append((x_start + k * spacing, y0 + r * row_dy))
    assert len(strings) == 78
    sid, did, xs, ys, zs = [], [], [], [], []
    for s, (sx, sy) in enumerate(strings):
        for d in range(60):
            z = 500.0 - 17.0 * d + 0.013 * ((s * 7) % 11)  # small per-string vertical offsets
            dx = dy = 0.0
            if bend:
                dx = 0.35 * math.sin(0.11 * d + 0.7 * s)
                dy = 0.25 * math.cos(0.07 * d + 1.3 * s)
            sid.append(s + 1)
            did.append(d + 1)
            xs.append(sx + dx)
            ys.append(sy + dy)
            zs.append(z)
    # DeepCore-like strings around the central string
    cx, cy = strings[35]
    dc = [(72.0, a) for a in (30.0, 90.0, 150.0, 210.0, 270.0, 330.0)] + [(41.0, 15.0), (41.0, 195.0)]
    for k, (rad, ang) in enumerate(dc):
        sx = cx + rad * math.cos(math.radians(ang)) + 20.0
        sy = cy + rad * math.sin(math.radians(ang)) + 31.0
        for d in range(60):
            if d < 10:
                z = 190.0 - 10.0 * d
            else:
                z = -157.0 - 7.0 * (d - 10)
            dx = dy = 0.0
            if bend:
                dx = 0.2 * math.sin(0.13 * d + k)
                dy = 0.2 * math.cos(0.09 * d + 2.0 * k)
            sid.append(79 + k)
            did.append(d + 1)
            xs.append(sx + dx)
            ys.append(sy + dy)
            zs.append(z)
    return SimpleGeometry(sid, did, xs, ys, zs, DOM_RADIUS * oversize)


def read_geometry_text_file(path, om_radius):
    data = np.loadtxt(path, ndmin=2)
    return SimpleGeometry(data[:, 0].astype(int), data[:, 1].astype(int), data[:, 2], data[:, 3], data[:, 4], om_radius)
