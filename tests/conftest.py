import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Before collection: the skipif markers of the tests that need oracle/_ref (the reference compiled for the host) look
    # for the built libraries when their module is imported.  In a fresh tree they would all be skipped on the first run.
    if os.environ.get("PYTEST_XDIST_WORKER") is None:
        try:
            import __graft_entry__ as entry
            from clsim_b200 import capi
            if not os.path.isfile(capi.LIB_PATH):
                entry.build_product()
            entry.build_oracle()
        except Exception as ex:   # noqa: BLE001 -- the session fixture below reports it properly
            sys.stderr.write("conftest: building the native libraries failed: %s\n" % (ex,))


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Both native libraries must exist; build them if the tree is fresh (nvcc cross-compiles
    without a GPU)."""
    import __graft_entry__ as entry
    from clsim_b200 import capi
    if not os.path.isfile(capi.LIB_PATH):
        entry.build_product()
    entry.build_oracle()
    yield


def _has_gpu():
    # tests/test_hostcheck.py runs GPU tests of the reference-order path in a subprocess against the CUDA sources compiled
    # for the host (tests/hostcheck): there the "device" is the CPU
    if os.environ.get("CLSIM_HOSTCHECK") == "1":
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def has_gpu():
    return _has_gpu()


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here; run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
