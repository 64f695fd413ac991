"""Row R7 pinned to the reference's own output: the geometry tables of the product (clsim_b200/csrc/tables.cpp) and of
the oracle against what the REFERENCE'S geometry source generator emits for the same detector.

private/opencl/I3CLSimHelperGenerateGeometrySource.cxx (1279 lines: string templates with 16-bit offsets, the x-y cell
division per subdetector, the z layering per "string set") is compiled UNMODIFIED from where it lies under
/root/reference into oracle/_ref/libclsim_ref_geometry.so (oracle/ref_shim/ref_geometry.cpp; IceTray logging,
boost::lexical_cast / BOOST_FOREACH and the OpenCL scalar typedefs are stand-in headers under oracle/ref_shim/host/).
It returns the OpenCL source text the reference would compile into its kernel, and the three side buffers it uploads.
The text is parsed the way an OpenCL compiler reads it (`1.5e+00f` is a float literal) and every number is held against
the product's table of the same meaning: equal as float32 / as integers, for the IceCube-like detector and for the edge
geometries of tests/test_tables.py.

The library exists wherever /root/reference did at build time (this container); it travels to the GPU box prebuilt."""
import numpy as np
import pytest

from clsim_b200 import capi, geometry
from clsim_b200.description import SimpleGeometry
from oracle import pyoracle
from tests.scenes import make_scene

pytestmark = pytest.mark.skipif(not pyoracle.ref_geometry_available(), reason="oracle/_ref/libclsim_ref_geometry.so not built (no /root/reference at build time)")


def f32(v):
    return np.asarray(v, dtype=np.float64).astype(np.float32)


def assert_tables_equal_reference(t, geo):
    text, layer_to_om, string_ids, dom_ids = pyoracle.ref_geometry_source(geo)
    d, a = pyoracle.parse_generated_source(text)

    def same_floats(ours, theirs, what):
        assert np.array_equal(f32(ours), f32(theirs)), what

    assert t["num_strings"] == d["NUM_STRINGS"] == d["GEO_DOM_POS_NUM_STRINGS"]
    same_floats(t["om_radius"], d["OM_RADIUS"], "OM_RADIUS")
    same_floats(t["string_max_radius"], d["GEO_STRING_MAX_RADIUS"], "GEO_STRING_MAX_RADIUS")
    for ours, theirs in (("string_pos_x", "geoStringPosX"), ("string_pos_y", "geoStringPosY"), ("string_min_z", "geoStringMinZ"),
                         ("string_max_z", "geoStringMaxZ"), ("layer_start_z", "geoLayerStartZ"), ("layer_height", "geoLayerHeight"),
                         ("tmpl_z", "geoDomPosTemplatePositionsZ_flat"), ("string_mean_x", "geoDomPosStringMeanPosX"),
                         ("string_mean_y", "geoDomPosStringMeanPosY")):
        same_floats(t[ours], a[theirs], theirs)
    # z layering: string -> set, layers per set, (set, layer) -> DOM index; the buffer is what the reference uploads
    assert t["string_in_set"] == a["geoStringInStringSet"]
    assert t["num_sets"] == d["GEO_LAYER_STRINGSET_NUM"] and t["max_layers"] == d["GEO_LAYER_STRINGSET_MAX_NUM_LAYERS"]
    assert t["layer_num"] == a["geoLayerNum"]
    assert len(t["layer_to_om"]) == d["GEO_geoLayerToOMNumIndexPerStringSet_BUFFER_SIZE"] == len(layer_to_om)
    assert np.array_equal(np.asarray(t["layer_to_om"], dtype=np.uint16), layer_to_om)
    # x-y cell division, one grid per subdetector
    assert len(t["cells"]) == d["GEO_CELL_NUM_SUBDETECTORS"]
    for i, c in enumerate(t["cells"]):
        assert (c["num_x"], c["num_y"]) == (d["GEO_CELL_NUM_X_%d" % i], d["GEO_CELL_NUM_Y_%d" % i])
        same_floats(c["start_width"], [d["GEO_CELL_START_X_%d" % i], d["GEO_CELL_START_Y_%d" % i], d["GEO_CELL_WIDTH_X_%d" % i],
                                       d["GEO_CELL_WIDTH_Y_%d" % i]], "cell grid %d" % i)
        assert np.array_equal(np.asarray(c["index"], dtype=np.uint16), np.asarray(a["geoCellIndex_%d" % i], dtype=np.uint16))
    # DOM positions: string templates, 16-bit offsets from the string's mean where the reference uses them
    assert t["max_dom_index"] == d["GEO_MAX_DOM_INDEX"]
    assert len(t["tmpl_z"]) == d["GEO_DOM_POS_NUM_FLAT_LIST_ENTRIES"]
    assert t["string_tmpl_start"] == a["geoDomPosStringStartIndexInTemplateDomList"]
    if "GEO_DOM_POS_MAX_ABS_X_MULTIPLIER_IN_TEMPLATE" in d:
        same_floats(t["tmpl_mul"], [d["GEO_DOM_POS_MAX_ABS_X_MULTIPLIER_IN_TEMPLATE"], d["GEO_DOM_POS_MAX_ABS_Y_MULTIPLIER_IN_TEMPLATE"]], "multipliers")
        assert t["tmpl_x"] == a["geoDomPosTemplatePositionsX_flat"] and t["tmpl_y"] == a["geoDomPosTemplatePositionsY_flat"]
    else:
        raise AssertionError("the reference wrote float templates here; the product's tables assume 16-bit offsets")
    # ... and decoded the way the generated geometryGetDomPosition() does, every DOM is where both say it is
    scale = f32(t["tmpl_mul"])
    x = f32(t["tmpl_x"]) * scale[0]
    y = f32(t["tmpl_y"]) * scale[1]
    assert np.array_equal(x, f32(a["geoDomPosTemplatePositionsX_flat"]) * f32(d["GEO_DOM_POS_MAX_ABS_X_MULTIPLIER_IN_TEMPLATE"]))
    assert np.array_equal(y, f32(a["geoDomPosTemplatePositionsY_flat"]) * f32(d["GEO_DOM_POS_MAX_ABS_Y_MULTIPLIER_IN_TEMPLATE"]))
    # the ID rewrite tables (I3CLSimStepToPhotonConverterOpenCL.cxx:1565-1619)
    assert t["string_index_to_id"] == string_ids
    assert t["dom_index_to_id"] == dom_ids


def tables(sc, geo):
    t_p = capi.describe_tables(sc.medium, geo, sc.generators, sc.bias, sc.options())
    t_o = pyoracle.Scene(sc.medium, geo, sc.generators, sc.bias, sc.options()).tables()
    return t_p, t_o


@pytest.mark.parametrize("oversize,kind", [(5.0, "ic86"), (1.0, "ic86"), (16.0, "ic86"), (5.0, "ring"), (1.0, "ring")])
def test_detector_tables_equal_the_reference_generators(oversize, kind):
    sc = make_scene("spice_mie", oversize, kind)
    t_p, t_o = tables(sc, sc.geo)
    assert_tables_equal_reference(t_p, sc.geo)
    assert_tables_equal_reference(t_o, sc.geo)


def test_edge_geometries_equal_the_reference_generators():
    sc = make_scene("spice_mie")
    z = [100.0 - 17.0 * i for i in range(12) if i != 5]
    cases = [
        # a single string
        SimpleGeometry([7] * 10, list(range(1, 11)), [3.0] * 10, [4.0] * 10, [100.0 - 17.0 * i for i in range(10)], 0.8255),
        # a string with a missing DOM (…GeometrySource.cxx:821-832)
        SimpleGeometry([1] * 11 + [2] * 11, list(range(11)) * 2, [0.0] * 11 + [125.0] * 11, [0.0] * 22, z + z, 0.8255),
        # two subdetectors: separate cell grids, ordered by name
        SimpleGeometry([1] * 5 + [2] * 5, list(range(5)) * 2, [0.0] * 5 + [60.0] * 5, [0.0] * 10, [10.0 * i for i in range(5)] * 2, 0.5,
                       subdetectors=["IceCube"] * 5 + ["DeepCore"] * 5),
        # string IDs that are not 1..N, negative ones, DOM lists given out of order
        SimpleGeometry([40, -3, 40, -3, 7, 7], [2, 9, 1, 8, 60, 61], [10.0, 90.0, 10.2, 90.1, -50.0, -50.0], [5.0, 5.0, 5.1, 5.0, 70.0, 70.3],
                       [-20.0, 33.0, 14.0, 50.0, 0.0, -17.0], 0.3),
        # perfectly straight strings: the 16-bit offsets are all zero, the multipliers rounding noise of the mean
        geometry.make_ic86_like_geometry(5.0, bend=False),
    ]
    for geo in cases:
        t_p, t_o = tables(sc, geo)
        assert_tables_equal_reference(t_p, geo)
        assert_tables_equal_reference(t_o, geo)


def test_reference_generator_errors_are_the_products():
    """What the reference refuses, the product refuses (…GeometrySource.cxx:728-735)."""
    sc = make_scene("spice_mie")
    for geo, pattern in ((SimpleGeometry([], [], [], [], [], 0.5), "Empty geometry"),
                         (SimpleGeometry([1, 1], [1, 2], [0, 0], [0, 0], [0, 10], -1.0), "OM radius")):
        with pytest.raises(RuntimeError):
            pyoracle.ref_geometry_source(geo)
        with pytest.raises(capi.ClsimCudaError, match=pattern):
            capi.describe_tables(sc.medium, geo, sc.generators, sc.bias, sc.options())
