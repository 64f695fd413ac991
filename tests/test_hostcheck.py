"""The CUDA sources of everything but the fast kernel, run on the CPU: engine.cu, kernel_reference.cu, mcpe.cu, stepgen.cu and
tabulate.cu compiled for the host under a stand-in CUDA runtime (tests/hostcheck/: one syntax rewrite, the three `<<<>>>`
launches), and the GPU tests of those units -- the same test files the B200 runs, unmodified -- executed against that build in a
subprocess (CLSIMCU_LIB points clsim_b200.capi at it, CLSIM_HOSTCHECK=1 tells tests/conftest.py that the "device" is the CPU
and tests/scenes.py to shrink the bunches and to take the reference-order kernel where a test only needs SOME kernel).

What this shows without a GPU: the reference-order CUDA kernel's SOURCE gives the oracle's hit lists (on the device it differs
from the oracle in the last bits of CUDA's libm; here both run on glibc), the engine's threads / staging / result assembly /
error propagation work, the photon -> MCPE converter, the step generator and the reference-order table maker equal their
oracles, and a step at infinity ends at once.  What it does not show: anything about the fast kernel (not in the build) or
about the device.  This is a test of source text -- the product has no CPU path and `clsim_b200` never loads this library."""
import os
import subprocess
import sys

import pytest

from tests import hostcheck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SELECTION = {
    # file: -k expression (None = the whole file)
    "tests/test_gpu_reference_kernel.py": None,
    "tests/test_gpu_tabulator.py": "spherical_table_equals or cylindrical_table_with or full_azimuth_table",      # the reference-order table maker
    "tests/test_gpu_stepgen.py": "every_step_of_every or argument_errors or reference_shaped",
    "tests/test_gpu_mcpe.py": "not device_rng_draw_assignment",      # (45 s of oracle work in Python; runs on the GPU box)
    "tests/test_gpu_engine.py": "not two_converters_disjoint and not fast_kernel_converter_history",      # (those name the fast kernel)
    "tests/test_zz_gpu_steps_at_infinity.py": "1]",      # the reference-order kernel's parametrisation
    # not a GPU test: the host-compiled kernel against the oracle, byte for byte (both on glibc)
    "tests/hostcheck/check_bit_identity.py": None,
}


def test_host_check_build_rewrites_only_the_launches():
    lib = hostcheck.build()
    assert os.path.isfile(lib)
    with open(os.path.join(hostcheck.BUILD, "rewrite.log")) as f:
        log = f.read().strip().split("\n")
    assert log[-1] == "3 launches rewritten"
    assert all("<<<" in line and "HOSTCHECK_LAUNCH(" in line for line in log[:-1])
    # the product's loader knows nothing of it
    from clsim_b200 import capi
    assert "hostcheck" not in capi.LIB_PATH or os.environ.get("CLSIMCU_LIB")


# the C++ test programs of the drop-in classes (clsim_b200/host/test_*.cxx, built by __graft_entry__.build_host_class), with
# the arguments tests/test_host_class.py gives them on a GPU box
CXX_PROGRAMS = {
    "test_converter_cuda": ["--gpu"],          # I3CLSimStepToPhotonConverterCUDA: 20 bunches, every setter's contract, history, destructor with work queued
    "test_server_inprocess": ["--gpu", "2"],   # two converters behind I3CLSimServerInProcess, three clients, photon conservation
    "test_neighbours_cuda": ["--gpu"],         # I3CLSimPhotonToMCPEConverterCUDA / I3CLSimStepGeneratorCUDA attached to a converter
}


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    """All selections at once, each in its own pytest process (they are independent; the wall time is the slowest one's)."""
    lib = hostcheck.build()
    # (the table is written once, here, by the product library's host-only entry -- not by ten processes at a time)
    from clsim_b200 import capi
    from clsim_b200.sharding import TOTAL_ROWS_RESERVED
    capi.safeprime_multipliers(0, TOTAL_ROWS_RESERVED)
    env = dict(os.environ, CLSIM_HOSTCHECK="1", CLSIMCU_LIB=lib, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""),
               OMP_NUM_THREADS="2",
               # the MWC multiplier table is memoised next to the product library; the host check build lives elsewhere
               CLSIMCU_SAFEPRIMES_CACHE=os.path.join(ROOT, "clsim_b200", "data", "safeprimes_base32.bin"))
    procs = {}
    for path, expr in SELECTION.items():
        cmd = [sys.executable, "-m", "pytest", path, "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"]
        if expr:
            cmd += ["-k", expr]
        procs[path] = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    # the C++ programs link libclsimcuda.so by name (DT_RUNPATH, so LD_LIBRARY_PATH comes first): a directory in which that
    # name is the host check build
    import __graft_entry__ as entry
    entry.build_host_class()
    libdir = tmp_path_factory.mktemp("hostcheck_lib")
    os.symlink(lib, os.path.join(str(libdir), "libclsimcuda.so"))
    cxx_env = dict(env, LD_LIBRARY_PATH=str(libdir) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    for name, args in CXX_PROGRAMS.items():
        binary = os.path.join(ROOT, "clsim_b200", "host", "build", name)
        procs[name] = subprocess.Popen([binary] + args, cwd=ROOT, env=cxx_env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    yield procs
    for p in procs.values():
        if p.poll() is None:
            p.kill()


@pytest.mark.parametrize("path", sorted(SELECTION))
def test_gpu_tests_pass_on_the_host_compiled_sources(runs, path):
    out, _ = runs[path].communicate(timeout=1500)
    tail = out[-3000:]
    assert runs[path].returncode == 0, tail
    assert " passed" in tail and " failed" not in tail, tail


@pytest.mark.parametrize("name", sorted(CXX_PROGRAMS))
def test_cxx_programs_of_the_drop_in_classes_pass_on_the_host_compiled_sources(runs, name):
    out, _ = runs[name].communicate(timeout=1500)
    tail = out[-3000:]
    assert runs[name].returncode == 0, tail
    assert " 0 failed" in tail, tail
