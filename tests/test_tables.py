"""The product's table flattening (clsim_b200/csrc/tables.cpp) against the oracle's independent
restatement of the reference generators, bit for bit, plus structural facts the reference's
generator guarantees.  CPU only (table building is host code)."""
import numpy as np
import pytest

from clsim_b200 import capi, geometry, ice
from clsim_b200.description import SimpleGeometry
from oracle import pyoracle
from tests.scenes import make_scene


def both(sc, geo=None, **opts):
    geo = geo or sc.geo
    t_p = capi.describe_tables(sc.medium, geo, sc.generators, sc.bias, sc.options(**opts))
    t_o = pyoracle.Scene(sc.medium, geo, sc.generators, sc.bias, sc.options(**opts)).tables()
    return t_p, t_o


@pytest.mark.parametrize("name,oversize,kind", [
    ("spice_mie", 5.0, "ic86"), ("spice_lea", 1.0, "ic86"), ("homogeneous", 5.0, "ring"), ("spice_1", 3.0, "ic86")])
def test_product_tables_equal_oracle_tables(name, oversize, kind):
    if name == "spice_1":
        sc = make_scene("spice_mie", oversize, kind)
        sc.medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_1")
    else:
        sc = make_scene(name, oversize, kind)
    t_p, t_o = both(sc)
    assert t_p.keys() == t_o.keys()
    for k in t_p:
        assert t_p[k] == t_o[k], k


def test_geometry_table_invariants():
    sc = make_scene("spice_mie")
    t, _ = both(sc)
    geo = sc.geo
    assert t["num_strings"] == 86 and t["max_dom_index"] == 60
    # at most one string per cell, every string reachable from its own position
    for c in t["cells"]:
        idx = np.array(c["index"]).reshape(c["num_y"], c["num_x"])
        sx, sy, wx, wy = c["start_width"]
        for s in range(86):
            cx = int((t["string_pos_x"][s] - sx) / wx)
            cy = int((t["string_pos_y"][s] - sy) / wy)
            assert idx[min(cy, c["num_y"] - 1), min(cx, c["num_x"] - 1)] == s
    # every DOM is found through its string set's z layering, and decodes to its position within 1 mm
    l2o = np.array(t["layer_to_om"])
    sets = t["string_in_set"]
    order = {}
    for i, (s, d) in enumerate(zip(geo.stringIDs, geo.domIDs)):
        order.setdefault(int(s), []).append(i)
    sid_sorted = sorted(order)
    assert t["string_index_to_id"] == sid_sorted
    for si, sid in enumerate(sid_sorted):
        st = sets[si]
        for di, i in enumerate(order[sid]):
            layer = int((geo.posZ[i] - t["layer_start_z"][st]) / t["layer_height"][st])
            assert l2o[st * t["max_layers"] + layer] == di
            at = t["string_tmpl_start"][si] + di
            x = t["tmpl_x"][at] * t["tmpl_mul"][0] + t["string_mean_x"][si]
            y = t["tmpl_y"][at] * t["tmpl_mul"][1] + t["string_mean_y"][si]
            assert abs(x - geo.posX[i]) < 1e-3 and abs(y - geo.posY[i]) < 1e-3 and abs(t["tmpl_z"][at] - geo.posZ[i]) < 1e-4
            assert t["dom_index_to_id"][si][di] == geo.domIDs[i]
    assert len(l2o) % 64 == 0
    assert t["string_max_radius"] >= t["om_radius"]


def test_edge_geometries():
    sc = make_scene("spice_mie")
    # a single string
    g1 = SimpleGeometry([7] * 10, list(range(1, 11)), [3.0] * 10, [4.0] * 10, [100.0 - 17.0 * i for i in range(10)], 0.8255)
    t_p, t_o = both(sc, g1)
    assert t_p == t_o and t_p["cells"][0]["num_x"] == 1 and t_p["num_sets"] == 1
    # a string with a missing DOM (double spacing must not inflate the mean spacing, …GeometrySource.cxx:821-832)
    z = [100.0 - 17.0 * i for i in range(12) if i != 5]
    g2 = SimpleGeometry([1] * 11 + [2] * 11, list(range(11)) * 2, [0.0] * 11 + [125.0] * 11, [0.0] * 22, z + z, 0.8255)
    t_p, t_o = both(sc, g2)
    assert t_p == t_o
    assert t_p["num_sets"] == 1      # identical layering is shared ("string set")
    # two subdetectors get separate cell grids, ordered by subdetector name
    g3 = SimpleGeometry([1] * 5 + [2] * 5, list(range(5)) * 2, [0.0] * 5 + [60.0] * 5, [0.0] * 10, [10.0 * i for i in range(5)] * 2,
                        0.5, subdetectors=["IceCube"] * 5 + ["DeepCore"] * 5)
    t_p, t_o = both(sc, g3)
    assert t_p == t_o and len(t_p["cells"]) == 2
    assert t_p["cells"][0]["index"] == [1]   # "DeepCore" < "IceCube": grid 0 holds string index 1
    # perfectly straight strings: quantisation scale 0, positions decode to the mean
    g4 = geometry.make_ic86_like_geometry(5.0, bend=False)
    t_p, t_o = both(sc, g4)
    assert t_p == t_o and max(t_p["tmpl_mul"]) < 1e-12  # rounding noise of the mean only
    # errors follow the reference
    with pytest.raises(capi.ClsimCudaError, match="Empty geometry"):
        capi.describe_tables(sc.medium, SimpleGeometry([], [], [], [], [], 0.5), sc.generators, sc.bias, sc.options())
    with pytest.raises(capi.ClsimCudaError, match="OM radius"):
        capi.describe_tables(sc.medium, SimpleGeometry([1, 1], [1, 2], [0, 0], [0, 0], [0, 10], -1.0), sc.generators, sc.bias, sc.options())


def test_float_literal_rounding_of_medium_tables():
    """Tables are float(double printed with 10 digits): ToFloatString semantics (quirk 5)."""
    sc = make_scene("spice_mie")
    t, _ = both(sc)
    want = [np.float32(float("%.10e" % v)) for v in sc.medium.b400]
    assert np.array_equal(np.array(t["medium"]["b400"], dtype=np.float32), np.array(want, dtype=np.float32))
    gen = t["wlen_generators"][0]
    assert gen["acu"][0] == 0.0 and gen["acu"][-1] == 1.0
    assert np.all(np.diff(gen["acu"]) >= 0)
