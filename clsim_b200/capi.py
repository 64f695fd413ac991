"""ctypes binding of libclsimcuda.so (include/clsimcuda.h).

The library is built in-tree by ``__graft_entry__.build()``.  There is deliberately no
fallback: if the shared library is missing this module raises, and if no CUDA device is
usable ``Engine`` raises with the library's message.
"""
import ctypes as C
import json
import os

import numpy as np

from .description import PHOTON_DTYPE, STEP_DTYPE, ConfigStruct, ResultStruct, build_config

_HERE = os.path.dirname(os.path.abspath(__file__))
# CLSIMCU_LIB selects another build of the same library: kernel variants for A/B runs (tools/build_variants.py), and -- in the
# test suite only -- the CUDA sources compiled for the host (tests/hostcheck, a test of the source text that holds no fast kernel).
# Unset, as in every product use, the one library there is is libclsimcuda.so, and it needs a CUDA device: no CPU fallback.
LIB_PATH = os.environ.get("CLSIMCU_LIB") or os.path.join(_HERE, "libclsimcuda.so")
_lib = None

# every entry point declared in include/clsimcuda.h
SYMBOLS = [
    "clsimcu_create", "clsimcu_destroy", "clsimcu_enqueue", "clsimcu_get_result", "clsimcu_release_result",
    "clsimcu_queue_size", "clsimcu_more_photons_available", "clsimcu_workgroup_size", "clsimcu_max_num_workitems",
    "clsimcu_get_statistics", "clsimcu_upload_resident", "clsimcu_run_resident", "clsimcu_download_resident",
    "clsimcu_download_resident_history", "clsimcu_rng_get", "clsimcu_rng_set", "clsimcu_describe_tables", "clsimcu_describe_tables_from_config",
    "clsimcu_describe_collision_map_from_config",
    "clsimcu_safeprime_multipliers", "clsimcu_seed_rng_states", "clsimcu_download_resident_rng_tags", "clsimcu_last_error", "clsimcu_version",
    "clsimcu_sizeof_config", "clsimcu_device_count",
    "clsimcu_mcpe_create", "clsimcu_mcpe_destroy", "clsimcu_mcpe_convert", "clsimcu_mcpe_rng_get", "clsimcu_attach_mcpe_converter",
    "clsimcu_stepgen_create", "clsimcu_stepgen_destroy", "clsimcu_stepgen_generate", "clsimcu_stepgen_rng_get", "clsimcu_enqueue_sources",
    "clsimcu_tabulator_create", "clsimcu_tabulator_destroy", "clsimcu_tabulator_enqueue", "clsimcu_tabulator_finish", "clsimcu_tabulator_info",
    "clsimcu_tabulator_get_table",
]

STAT_KEYS = ["TotalDeviceTime", "TotalHostTime", "NumKernelCalls", "TotalNumPhotonsGenerated", "TotalNumPhotonsAtDOMs",
             "AverageDeviceTimePerPhoton", "AverageHostTimePerPhoton", "DeviceUtilization"]


class ClsimCudaError(RuntimeError):
    """Mirrors I3CLSimStepToPhotonConverter_exception (public/clsim/I3CLSimStepToPhotonConverter.h:57-65)."""

    def __init__(self, code, message):
        RuntimeError.__init__(self, message)
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("%s is missing: run `python __graft_entry__.py` (build()) first; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name in SYMBOLS:
            getattr(L, name)  # AttributeError if the ABI and the header drift apart
        L.clsimcu_last_error.restype = C.c_char_p
        L.clsimcu_version.restype = C.c_char_p
        L.clsimcu_sizeof_config.restype = C.c_size_t
        L.clsimcu_create.argtypes = [C.POINTER(ConfigStruct), C.POINTER(C.c_void_p)]
        L.clsimcu_destroy.argtypes = [C.c_void_p]
        L.clsimcu_enqueue.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]
        L.clsimcu_get_result.argtypes = [C.c_void_p, C.POINTER(ResultStruct)]
        L.clsimcu_release_result.argtypes = [C.c_void_p, C.POINTER(ResultStruct)]
        L.clsimcu_queue_size.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.clsimcu_more_photons_available.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.clsimcu_workgroup_size.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.clsimcu_max_num_workitems.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.clsimcu_get_statistics.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.clsimcu_upload_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.clsimcu_run_resident.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.clsimcu_download_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_download_resident_rng_tags.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.clsimcu_download_resident_history.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_rng_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.clsimcu_rng_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.clsimcu_describe_tables.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_describe_tables_from_config.argtypes = [C.POINTER(ConfigStruct), C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_describe_collision_map_from_config.argtypes = [C.POINTER(ConfigStruct), C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_safeprime_multipliers.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
        L.clsimcu_seed_rng_states.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        if L.clsimcu_sizeof_config() != C.sizeof(ConfigStruct):
            raise ImportError("clsimcu_config layout mismatch between description.py and libclsimcuda.so")
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise ClsimCudaError(rc, lib().clsimcu_last_error().decode())


def device_count():
    """Usable CUDA devices; raises when there is none (no CPU fallback)."""
    n = C.c_int(0)
    _check(lib().clsimcu_device_count(C.byref(n)))
    return n.value


def safeprime_multipliers(first, n):
    a = np.zeros(n, dtype=np.uint32)
    _check(lib().clsimcu_safeprime_multipliers(int(first), int(n), a.ctypes.data))
    return a


def seed_rng_states(seed, a):
    """Start states of the MWC streams with multipliers `a`, as an engine created with rng_seed = seed draws them (host only)."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    x = np.zeros(len(a), dtype=np.uint64)
    _check(lib().clsimcu_seed_rng_states(int(seed), a.ctypes.data, x.ctypes.data, len(a)))
    return x


def describe_tables(medium, geometry, wlen_generators, wlen_bias, options):
    """Table building only (host code, no GPU needed)."""
    cfg, keep = build_config(medium, geometry, wlen_generators, wlen_bias, options)
    need = C.c_size_t(0)
    _check(lib().clsimcu_describe_tables_from_config(C.byref(cfg), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    _check(lib().clsimcu_describe_tables_from_config(C.byref(cfg), buf, need.value, C.byref(need)))
    del keep
    return json.loads(buf.value.decode())


def describe_collision_map(medium, geometry, wlen_generators, wlen_bias, options, pixel_budget=40000):
    """The fast kernel's xy collision map for a pixel budget (host code, no GPU needed)."""
    cfg, keep = build_config(medium, geometry, wlen_generators, wlen_bias, options)
    need = C.c_size_t(0)
    _check(lib().clsimcu_describe_collision_map_from_config(C.byref(cfg), int(pixel_budget), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    _check(lib().clsimcu_describe_collision_map_from_config(C.byref(cfg), int(pixel_budget), buf, need.value, C.byref(need)))
    del keep
    return json.loads(buf.value.decode())


class Result(object):
    __slots__ = ("identifier", "photons", "history", "num_photons_generated", "num_hits_counted", "mcpes")


class Engine(object):
    """Thin object wrapper over the C ABI; see converter.py for the reference-shaped class."""

    def __init__(self, medium, geometry, wlen_generators, wlen_bias, options):
        self._h = C.c_void_p()
        cfg, self._keep = build_config(medium, geometry, wlen_generators, wlen_bias, options)
        self.history_entries = int(options.photon_history_entries)
        _check(lib().clsimcu_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().clsimcu_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enqueue(self, steps, identifier):
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        _check(lib().clsimcu_enqueue(self._h, steps.ctypes.data, len(steps), int(identifier)))

    def get_result(self):
        r = ResultStruct()
        _check(lib().clsimcu_get_result(self._h, C.byref(r)))
        out = Result()
        out.identifier = int(r.identifier)
        n = int(r.num_photons)
        if n:
            # one copy out of the library's buffer (released below)
            out.photons = np.frombuffer((C.c_char * (n * PHOTON_DTYPE.itemsize)).from_address(r.photons), dtype=PHOTON_DTYPE).copy()
        else:
            out.photons = np.zeros(0, dtype=PHOTON_DTYPE)
        out.history = None
        if self.history_entries > 0:
            if n and r.history:
                cnt = n * self.history_entries * 4
                out.history = np.ctypeslib.as_array(r.history, shape=(cnt,)).reshape(n, self.history_entries, 4).copy()
            else:
                out.history = np.zeros((0, self.history_entries, 4), dtype=np.float32)
        out.num_photons_generated = int(r.num_photons_generated)
        out.num_hits_counted = int(r.num_hits_counted)
        out.mcpes = None
        if getattr(self, "_mcpe", None) is not None:
            from .mcpe import MCPE_DTYPE
            m = int(r.num_mcpes)
            out.mcpes = np.frombuffer(C.string_at(r.mcpes, m * 16), dtype=MCPE_DTYPE).copy() if m else np.zeros(0, dtype=MCPE_DTYPE)
        _check(lib().clsimcu_release_result(self._h, C.byref(r)))
        return out

    def queue_size(self):
        v = C.c_size_t(0)
        _check(lib().clsimcu_queue_size(self._h, C.byref(v)))
        return v.value

    def more_photons_available(self):
        v = C.c_int(0)
        _check(lib().clsimcu_more_photons_available(self._h, C.byref(v)))
        return bool(v.value)

    def workgroup_size(self):
        v = C.c_size_t(0)
        _check(lib().clsimcu_workgroup_size(self._h, C.byref(v)))
        return v.value

    def max_num_workitems(self):
        v = C.c_size_t(0)
        _check(lib().clsimcu_max_num_workitems(self._h, C.byref(v)))
        return v.value

    def statistics(self):
        arr = (C.c_double * 8)()
        _check(lib().clsimcu_get_statistics(self._h, arr))
        return dict(zip(STAT_KEYS, [float(v) for v in arr]))

    # resident path -------------------------------------------------------------------------
    def upload_resident(self, steps):
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        _check(lib().clsimcu_upload_resident(self._h, steps.ctypes.data, len(steps)))

    def run_resident(self, repeat=1):
        ms, gen, hits, seg = C.c_double(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().clsimcu_run_resident(self._h, int(repeat), C.byref(ms), C.byref(gen), C.byref(hits), C.byref(seg)))
        return {"kernel_ms": ms.value, "photons": gen.value, "hits": hits.value, "segments": seg.value}

    def download_resident(self, cap=None):
        n = C.c_size_t(0)
        _check(lib().clsimcu_download_resident(self._h, None, 0, C.byref(n)))
        k = n.value if cap is None else min(n.value, cap)
        out = np.zeros(k, dtype=PHOTON_DTYPE)
        if k:
            _check(lib().clsimcu_download_resident(self._h, out.ctypes.data, k, C.byref(n)))
        return out

    def download_resident_history(self, k, entries):
        """[k][entries][4]: the scatter points of the hits of the last resident run, forward order, unused rows NaN"""
        out = np.zeros((k, entries, 4), dtype=np.float32)
        n = C.c_size_t(0)
        _check(lib().clsimcu_download_resident_history(self._h, out.ctypes.data, k, C.byref(n)))
        return out

    def download_resident_rng_tags(self, k):
        x = np.zeros(2 * k, dtype=np.uint64)
        a = np.zeros(2 * k, dtype=np.uint32)
        _check(lib().clsimcu_download_resident_rng_tags(self._h, x.ctypes.data, a.ctypes.data, k))
        return x.reshape(k, 2), a.reshape(k, 2)  # per record: (x_create, x_propagate), (a_create, a_propagate)

    def rng_get(self, n):
        x = np.zeros(n, dtype=np.uint64)
        a = np.zeros(n, dtype=np.uint32)
        _check(lib().clsimcu_rng_get(self._h, x.ctypes.data, a.ctypes.data, n))
        return x, a

    def rng_set(self, x, a):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        a = np.ascontiguousarray(a, dtype=np.uint32)
        _check(lib().clsimcu_rng_set(self._h, x.ctypes.data, a.ctypes.data, len(x)))

    def tables(self):
        need = C.c_size_t(0)
        _check(lib().clsimcu_describe_tables(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().clsimcu_describe_tables(self._h, buf, need.value, C.byref(need)))
        return json.loads(buf.value.decode())
