"""The C-ABI library loads without a GPU and exports exactly what include/clsimcuda.h declares.
No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from clsim_b200 import capi
from clsim_b200.description import PHOTON_DTYPE, STEP_DTYPE, ConfigStruct, ConverterOptions
from tests.scenes import make_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "clsimcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clsimcu_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = header_functions()
    assert len(names) >= 20
    lib = capi.lib()
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(capi.SYMBOLS) == names
    nm = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = set(re.findall(r" T (clsimcu_[a-z_0-9]+)", nm))
    assert exported == set(names)


def test_wire_format_sizes():
    assert STEP_DTYPE.itemsize == 48 and PHOTON_DTYPE.itemsize == 80
    assert capi.lib().clsimcu_sizeof_config() == C.sizeof(ConfigStruct)
    from oracle import pyoracle
    assert pyoracle.lib().oracle_sizeof_config() == C.sizeof(ConfigStruct)
    assert STEP_DTYPE.fields["num_photons"][1] == 32 and STEP_DTYPE.fields["source_type"][1] == 44
    assert PHOTON_DTYPE.fields["string_id"][1] == 44 and PHOTON_DTYPE.fields["om_id"][1] == 46
    assert PHOTON_DTYPE.fields["start_x"][1] == 48 and PHOTON_DTYPE.fields["dist_in_abs_lens"][1] == 76


def test_sass_is_sm100_and_has_no_tensor_or_ptx_fallback():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out


def test_config_validation_without_gpu():
    sc = make_scene("spice_mie")
    # both save-all and stop-on-detection: the reference throws in Compile (…OpenCL.cxx:507-508)
    with pytest.raises(capi.ClsimCudaError) as e:
        capi.describe_tables(sc.medium, sc.geo, sc.generators, sc.bias, sc.options(save_all_photons=True))
    assert "saveAllPhotons and stopDetectedPhotons" in str(e.value)
    with pytest.raises(capi.ClsimCudaError) as e:
        capi.describe_tables(sc.medium, sc.geo, [], sc.bias, sc.options())
    assert "WlenGenerators not set" in str(e.value)
    bad = ConverterOptions()
    cfg = ConfigStruct()
    cfg.struct_size = 3
    need = C.c_size_t(0)
    assert capi.lib().clsimcu_describe_tables_from_config(C.byref(cfg), None, 0, C.byref(need)) != 0
    del bad


def test_create_fails_loudly_without_a_device(has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    sc = make_scene("spice_mie", geo_kind="ring")
    with pytest.raises(capi.ClsimCudaError) as e:
        capi.Engine(sc.medium, sc.geo, sc.generators, sc.bias, sc.options(rng_n=64, max_num_workitems=64, kernel_mode=1))
    assert e.value.code == -3
    assert "no CPU fallback" in str(e.value)


def test_null_engine_calls_are_errors():
    lib = capi.lib()
    r = capi.ResultStruct()
    assert lib.clsimcu_get_result(None, C.byref(r)) != 0
    assert b"not initialized" in lib.clsimcu_last_error()
    steps = np.zeros(4, dtype=STEP_DTYPE)
    assert lib.clsimcu_enqueue(None, steps.ctypes.data, 4, 0) != 0
    assert lib.clsimcu_destroy(None) != 0


def test_product_does_not_use_the_oracle():
    """oracle/ is test infrastructure: the product library must neither link it nor mention it, and no
    module of the package may import it."""
    ldd = subprocess.run(["ldd", capi.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "oracle" not in ldd
    needed = subprocess.run(["readelf", "-d", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in needed
    pkg = os.path.join(ROOT, "clsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".cxx", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "clsim_oracle" not in text, f
