"""The drop-in boundary against the reference's own headers.

clsim_b200/host/I3CLSimStepToPhotonConverterCUDA.{h,cxx} is what INTEGRATION.md tells a maintainer of the reference to add to
their tree.  Stand-alone it is built against clsim_compat.h (a stand-in for the reference's interface); here it is built the way
the maintainer would build it: -DCLSIM_CUDA_IN_ICETRAY, against the reference's own public/clsim headers -- its abstract
I3CLSimStepToPhotonConverter, its step / photon records, its medium / geometry / function / random-value classes -- patched with
exactly the getters INTEGRATION.md section 2 lists (tools/integration_getters.py is that section in executable form), linked with
the reference's own sources of those classes and with libclsimcuda.so (oracle/Makefile: _ref/libclsim_icetray_mode.so; IceTray
itself is stand-in headers).  The class is driven through a pointer to the REFERENCE's abstract base with REFERENCE objects.

CPU only (Compile() and DescribeTables() are host code); needs /root/reference at build time."""
import numpy as np
import pytest

from clsim_b200 import capi
from oracle import pyoracle
from tests.scenes import add_flasher_generator, make_scene

pytestmark = pytest.mark.skipif(not pyoracle.icetray_mode_available(), reason="oracle/_ref not built (no /root/reference at build time)")


@pytest.mark.parametrize("name,flasher,opts", [
    ("homogeneous", False, {}),
    ("spice_mie", False, {}),
    ("spice_lea", True, {}),                                   # tilt (I3Matrix of corrections), anisotropy (ublas matrices), two generators
    ("spice_mie_tilt", False, {"stop_detected_photons": False, "photon_history_entries": 4}),
    ("spice_lea_notilt", False, {"pancake_factor": 1.0}),
])
def test_reference_objects_through_the_reference_interface_give_the_same_device_tables(name, flasher, opts):
    sc = make_scene(name)
    if flasher:
        sc = add_flasher_generator(sc)
    opt = sc.options(**opts)
    want = capi.describe_tables(sc.medium, sc.geo, sc.generators, sc.bias, opt)        # the Python path of the tests and bench.py
    got = pyoracle.icetray_mode_describe_tables(sc.medium, sc.geo, sc.generators, sc.bias, opt)
    assert set(got) == set(want)
    for key in want:
        assert got[key] == want[key], key
    assert len(got["medium"]["b400"]) == sc.medium.layersNum and got["num_strings"] == len(np.unique(sc.geo.stringIDs))


def test_ring_geometry_and_subdetector_names():
    sc = make_scene("spice_mie", geo_kind="ring")
    got = pyoracle.icetray_mode_describe_tables(sc.medium, sc.geo, sc.generators, sc.bias, sc.options())
    assert got == capi.describe_tables(sc.medium, sc.geo, sc.generators, sc.bias, sc.options())


def test_unknown_description_class_is_refused_with_the_references_exception():
    """A scattering model outside the path (here: a constant) -> I3CLSimStepToPhotonConverter_exception naming the class."""
    sc = make_scene("spice_mie")
    rc, message = pyoracle.icetray_mode_unknown_class_message(sc.medium, sc.generators, sc.bias)
    assert rc == 1 and "scattering angle distribution is of a class the CUDA converter does not know" in message
    assert "I3CLSimRandomValueConstant" in message
