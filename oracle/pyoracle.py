"""ctypes wrapper of the CPU oracle (TEST INFRASTRUCTURE -- import only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

from clsim_b200.description import PHOTON_DTYPE, STEP_DTYPE, ConfigStruct, build_config

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libclsim_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.isfile(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "clsim_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.oracle_sizeof_config.restype = C.c_size_t
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.POINTER(ConfigStruct)]
        L.oracle_scene_destroy.argtypes = [C.c_void_p]
        L.oracle_propagate.restype = C.c_uint64
        L.oracle_propagate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_propagate_single_photon.restype = C.c_int
        L.oracle_propagate_single_photon.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p,
                                                     C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.oracle_propagate_single_photon_split.restype = C.c_int
        L.oracle_propagate_single_photon_split.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32,
                                                           C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.oracle_rng_uniform_co.argtypes = [C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p, C.c_size_t]
        L.oracle_safeprimes.restype = C.c_int
        L.oracle_safeprimes.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_rng_seed_states.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_describe_tables.restype = C.c_int
        L.oracle_describe_tables.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.oracle_eval_wlen_function.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_eval_scalar_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_eval_vector_transform.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_sample.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p, C.c_size_t]
        L.oracle_scatter_direction.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def safeprimes(first, n):
    a = np.zeros(n, dtype=np.uint32)
    n2 = np.zeros(n, dtype=np.uint64)
    n1 = np.zeros(n, dtype=np.uint64)
    if lib().oracle_safeprimes(first, n, a.ctypes.data, n2.ctypes.data, n1.ctypes.data) != 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return a, n2, n1


def seed_states(seed, a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    x = np.zeros(len(a), dtype=np.uint64)
    lib().oracle_rng_seed_states(seed, a.ctypes.data, x.ctypes.data, len(a))
    return x


def rng_uniform_co(x, a, n):
    out = np.zeros(n, dtype=np.float32)
    xs = C.c_uint64(int(x))
    lib().oracle_rng_uniform_co(C.byref(xs), int(a), out.ctypes.data, n)
    return out, xs.value


def scatter_direction(in6):
    in6 = np.ascontiguousarray(in6, dtype=np.float32)
    out = np.zeros((len(in6), 3), dtype=np.float32)
    lib().oracle_scatter_direction(in6.ctypes.data, out.ctypes.data, len(in6))
    return out


class Scene(object):
    def __init__(self, medium, geometry, wlen_generators, wlen_bias, options):
        self._cfg, self._keep = build_config(medium, geometry, wlen_generators, wlen_bias, options)
        assert lib().oracle_sizeof_config() == C.sizeof(ConfigStruct)
        self.history_entries = int(options.photon_history_entries)
        self._h = lib().oracle_scene_create(C.byref(self._cfg))
        if not self._h:
            raise RuntimeError(lib().oracle_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_scene_destroy(self._h)
            self._h = None

    def propagate(self, steps, rng_x, rng_a, cap=None, num_threads=1):
        """-> (photons[PHOTON_DTYPE], hits_counted, stats dict, new rng_x, history or None)"""
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        n = len(steps)
        x = np.array(rng_x[:n], dtype=np.uint64, copy=True)
        a = np.ascontiguousarray(rng_a[:n], dtype=np.uint32)
        if cap is None:
            cap = max(1000, 10 * n)
        out = np.zeros(cap, dtype=PHOTON_DTYPE)
        hist = np.zeros((cap, self.history_entries, 4), dtype=np.float32) if self.history_entries else None
        stats = np.zeros(4, dtype=np.uint64)
        cnt = lib().oracle_propagate(self._h, steps.ctypes.data, n, x.ctypes.data, a.ctypes.data, out.ctypes.data, cap,
                                     hist.ctypes.data if hist is not None else None, int(num_threads), stats.ctypes.data)
        k = min(int(cnt), cap)
        st = {"photons": int(stats[0]), "segments": int(stats[1]), "crossings": int(stats[2]), "draws": int(stats[3])}
        return out[:k], int(cnt), st, x, (hist[:k] if hist is not None else None)

    def tabulate(self, axes, steps, rng_x, rng_a, reference, n_group, n_phase, angular_coefficients=None, step_length=1.0, squared=False):
        """Table-maker variant: -> (bins[float64], squared or None, entries, new rng_x).  `axes`: an object with
        .geometry and .axes (kind, power, min, max, n_bins), e.g. clsim_b200.tabulator.SphericalAxes."""
        class TableConfig(C.Structure):
            _fields_ = [("geometry", C.c_int32), ("num_axes", C.c_int32), ("axis_kind", C.c_int32 * 5), ("axis_power", C.c_uint32 * 5),
                        ("axis_bins", C.c_uint32 * 5), ("num_angular_coefficients", C.c_int32), ("axis_min", C.c_double * 5),
                        ("axis_max", C.c_double * 5), ("step_length", C.c_double), ("n_group", C.c_double), ("n_phase", C.c_double),
                        ("angular_coefficients", C.POINTER(C.c_double))]
        tc = TableConfig()
        tc.geometry, tc.num_axes = int(axes.geometry), len(axes.axes)
        for i, ax in enumerate(axes.axes):
            tc.axis_kind[i], tc.axis_power[i], tc.axis_bins[i], tc.axis_min[i], tc.axis_max[i] = ax.kind, ax.power, ax.n_bins, ax.min, ax.max
        tc.step_length, tc.n_group, tc.n_phase = float(step_length), float(n_group), float(n_phase)
        coef = np.ascontiguousarray(angular_coefficients if angular_coefficients is not None else [], dtype=np.float64)
        tc.num_angular_coefficients = len(coef)
        tc.angular_coefficients = coef.ctypes.data_as(C.POINTER(C.c_double))
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        n = len(steps)
        x = np.array(rng_x[:n], dtype=np.uint64, copy=True)
        a = np.ascontiguousarray(rng_a[:n], dtype=np.uint32)
        nb = 1
        for ax in axes.axes:
            nb *= ax.n_bins + 2
        bins = np.zeros(nb, dtype=np.float64)
        sq = np.zeros(nb, dtype=np.float64) if squared else None
        ref = (C.c_double * 7)(*[float(v) for v in reference])
        fn = lib().oracle_tabulate
        fn.restype = C.c_uint64
        fn.argtypes = [C.c_void_p, C.POINTER(TableConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        entries = fn(self._h, C.byref(tc), steps.ctypes.data, n, x.ctypes.data, a.ctypes.data, ref, bins.ctypes.data, sq.ctypes.data if squared else None)
        if entries == 0 and lib().oracle_last_error():
            pass
        return bins, sq, int(entries), x

    def single_photon(self, step, x, a, max_points=0):
        step = np.ascontiguousarray(step, dtype=STEP_DTYPE).reshape(1)
        out = np.zeros(1, dtype=PHOTON_DTYPE)
        traj = np.zeros((max(1, max_points), 8), dtype=np.float32)
        npts = C.c_int(0)
        xs = C.c_uint64(int(x))
        saved = lib().oracle_propagate_single_photon(self._h, step.ctypes.data, C.byref(xs), int(a), out.ctypes.data,
                                                     traj.ctypes.data if max_points else None, max_points, C.byref(npts))
        return bool(saved), out[0], traj[:min(npts.value, max_points)], xs.value, npts.value

    def single_photon_split(self, step, x_create, a_create, x_propagate, a_propagate, max_points=0):
        """max_points > 0: also the trajectory, rows (x, y, z, t, dx, dy, dz, abs_lens_left) at creation and after every segment"""
        step = np.ascontiguousarray(step, dtype=STEP_DTYPE).reshape(1)
        out = np.zeros(1, dtype=PHOTON_DTYPE)
        npts = C.c_int(0)
        traj = np.zeros((max(1, max_points), 8), dtype=np.float32)
        saved = lib().oracle_propagate_single_photon_split(self._h, step.ctypes.data, int(x_create), int(a_create), int(x_propagate), int(a_propagate),
                                                           out.ctypes.data, traj.ctypes.data if max_points else None, max_points, C.byref(npts))
        if max_points:
            return bool(saved), out[0], traj[:min(npts.value, max_points)]
        return bool(saved), out[0]

    def tables(self):
        need = C.c_size_t(0)
        lib().oracle_describe_tables(self._h, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        lib().oracle_describe_tables(self._h, buf, need.value, C.byref(need))
        return json.loads(buf.value.decode())

    def eval_wlen_function(self, which, layers, wlens):
        arr = np.ascontiguousarray(np.stack([np.asarray(layers, dtype=np.float32), np.asarray(wlens, dtype=np.float32)], axis=-1))
        out = np.zeros(len(arr), dtype=np.float32)
        lib().oracle_eval_wlen_function(self._h, which, arr.ctypes.data, out.ctypes.data, len(arr))
        return out

    def eval_scalar_field(self, which, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        out = np.zeros(len(xyz), dtype=np.float32)
        lib().oracle_eval_scalar_field(self._h, which, xyz.ctypes.data, out.ctypes.data, len(xyz))
        return out

    def eval_vector_transform(self, which, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        out = np.zeros_like(xyz)
        lib().oracle_eval_vector_transform(self._h, which, xyz.ctypes.data, out.ctypes.data, len(xyz))
        return out

    def sample(self, which, x, a, n):
        out = np.zeros(n, dtype=np.float32)
        xs = C.c_uint64(int(x))
        lib().oracle_sample(self._h, which, C.byref(xs), int(a), out.ctypes.data, n)
        return out, xs.value


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own propKernel text compiled for the host (oracle/ref_shim/).  Exists only where
# the library was built, i.e. in a container that holds /root/reference (the built .so travels to the GPU box).
# ---------------------------------------------------------------------------------------------------------
_REF_LIB = os.path.join(_HERE, "_ref", "libclsim_ref.so")
_ref_lib = None


def ref_available():
    return os.path.isfile(_REF_LIB)


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        L = C.CDLL(_REF_LIB)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.POINTER(ConfigStruct)]
        L.oracle_scene_destroy.argtypes = [C.c_void_p]
        L.oracle_propagate.restype = C.c_uint64
        L.oracle_propagate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_int, C.c_void_p]
        L.ref_variant_name.restype = C.c_char_p
        L.ref_variant_name.argtypes = [C.c_void_p, C.c_int]
        L.ref_propagate.restype = C.c_uint64
        L.ref_propagate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                    C.c_int, C.c_int]
        _ref_lib = L
    return _ref_lib


class RefScene(object):
    """A scene inside oracle/_ref/libclsim_ref.so: propagate() runs the REFERENCE's kernel text, oracle_propagate()
    the hand-written restatement compiled into the same library (the same code as libclsim_oracle.so)."""

    def __init__(self, medium, geometry, wlen_generators, wlen_bias, options):
        self._cfg, self._keep = build_config(medium, geometry, wlen_generators, wlen_bias, options)
        self.history_entries = int(options.photon_history_entries)
        self.with_flasher = len(wlen_generators) > 1   # the reference passes -DNO_FLASHER for a single generator (…OpenCL.cxx:648-649)
        self._h = ref_lib().oracle_scene_create(C.byref(self._cfg))
        if not self._h:
            raise RuntimeError(ref_lib().oracle_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            ref_lib().oracle_scene_destroy(self._h)
            self._h = None

    def variant(self):
        v = ref_lib().ref_variant_name(self._h, int(self.with_flasher))
        return v.decode() if v else None

    def _run(self, which, steps, rng_x, rng_a, cap, num_threads):
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        n = len(steps)
        x = np.array(rng_x[:n], dtype=np.uint64, copy=True)
        a = np.ascontiguousarray(rng_a[:n], dtype=np.uint32)
        if cap is None:
            cap = max(1000, 10 * n)
        out = np.zeros(cap, dtype=PHOTON_DTYPE)
        hist = np.zeros((cap, self.history_entries, 4), dtype=np.float32) if self.history_entries else None
        hp = hist.ctypes.data if hist is not None else None
        if which == "ref":
            cnt = ref_lib().ref_propagate(self._h, steps.ctypes.data, n, x.ctypes.data, a.ctypes.data, out.ctypes.data, cap, hp,
                                          int(self.with_flasher), int(num_threads))
            if cnt == 2 ** 64 - 1:
                raise RuntimeError(ref_lib().oracle_last_error().decode())
        else:
            stats = np.zeros(4, dtype=np.uint64)
            cnt = ref_lib().oracle_propagate(self._h, steps.ctypes.data, n, x.ctypes.data, a.ctypes.data, out.ctypes.data, cap, hp,
                                             int(num_threads), stats.ctypes.data)
        k = min(int(cnt), cap)
        return out[:k], int(cnt), x, (hist[:k] if hist is not None else None)

    def propagate(self, steps, rng_x, rng_a, cap=None, num_threads=1):
        """The reference's kernel text -> (photons, hits counted, new rng_x, history or None)"""
        return self._run("ref", steps, rng_x, rng_a, cap, num_threads)

    def oracle_propagate(self, steps, rng_x, rng_a, cap=None, num_threads=1):
        return self._run("oracle", steps, rng_x, rng_a, cap, num_threads)


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref/libclsim_ref_geometry.so: the reference's own geometry source generator
# (private/opencl/I3CLSimHelperGenerateGeometrySource.cxx) compiled unmodified (oracle/ref_shim/ref_geometry.cpp).
# ---------------------------------------------------------------------------------------------------------
_REF_GEOMETRY_LIB = os.path.join(_HERE, "_ref", "libclsim_ref_geometry.so")
_ref_geometry_lib = None


def ref_geometry_available():
    return os.path.isfile(_REF_GEOMETRY_LIB)


def ref_geometry_source(geometry):
    """-> (OpenCL source text the reference generates for `geometry`, geoLayerToOMNumIndexPerStringSet buffer,
    stringIndex -> string ID, per string: DOM index -> DOM ID).  Raises RuntimeError with the reference's message."""
    global _ref_geometry_lib
    if _ref_geometry_lib is None:
        L = C.CDLL(_REF_GEOMETRY_LIB)
        L.ref_geometry_generate.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_char_p), C.c_double]
        L.ref_geometry_error.restype = C.c_char_p
        L.ref_geometry_text.restype = C.c_char_p
        for name in ("ref_geometry_layer_to_om", "ref_geometry_string_ids"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.POINTER(C.c_void_p)]
        L.ref_geometry_dom_ids.restype = C.c_size_t
        L.ref_geometry_dom_ids.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _ref_geometry_lib = L
    L = _ref_geometry_lib
    g = geometry
    n = g.size()
    names = (C.c_char_p * max(1, n))(*[s.encode() for s in g.subdetectorNames])
    rc = L.ref_geometry_generate(n, g.stringIDs.ctypes.data, g.domIDs.ctypes.data, g.posX.ctypes.data, g.posY.ctypes.data, g.posZ.ctypes.data,
                                 names, float(g.OMRadius))
    if rc != 0:
        raise RuntimeError(L.ref_geometry_error().decode())
    text = L.ref_geometry_text().decode()
    p = C.c_void_p()
    k = L.ref_geometry_layer_to_om(C.byref(p))
    layer_to_om = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint16)), shape=(k,)).copy() if k else np.zeros(0, np.uint16)
    k = L.ref_geometry_string_ids(C.byref(p))
    string_ids = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(k,)).copy() if k else np.zeros(0, np.int32)
    q = C.c_void_p()
    k = L.ref_geometry_dom_ids(C.byref(p), C.byref(q))
    start = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint32)), shape=(k + 1,)).copy()
    flat = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(int(start[-1]),)).copy() if start[-1] else np.zeros(0, np.uint32)
    dom_ids = [flat[start[i]:start[i + 1]].tolist() for i in range(k)]
    return text, layer_to_om, string_ids.tolist(), dom_ids


def parse_generated_source(text):
    """#define NAME value  and  __constant TYPE name[...] = { ... };  of a generated OpenCL snippet -> (defines, arrays).
    Numbers are read the way an OpenCL compiler reads them: `1.5e+00f` is the float nearest to 1.5, `0xFFFF` is 65535."""
    import re

    def number(tok, as_float):
        tok = tok.strip()
        if tok.lower().startswith("0x"):
            return int(tok, 16)
        if tok.endswith("f"):
            return float(np.float32(float(tok[:-1])))
        if as_float or any(c in tok for c in ".eE"):
            return float(tok)
        return int(tok)

    text = re.sub(r"//[^\n]*", "", text)
    defines = {}
    for m in re.finditer(r"^#define\s+(\w+)\s+(\S+)\s*$", text, re.M):
        try:
            defines[m.group(1)] = number(m.group(2), False)
        except ValueError:
            defines[m.group(1)] = m.group(2)
    arrays = {}
    for m in re.finditer(r"__constant\s+(?:const\s+)?([\w ]+?)\s+(\w+)\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S):
        is_float = m.group(1).strip() in ("float", "double")
        arrays[m.group(2)] = [number(t, is_float) for t in m.group(3).replace("\n", " ").split(",") if t.strip()]
    return defines, arrays


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref/libclsim_ref_stepgen.so: the inline samplers of the reference's step generator
# (private/clsim/I3CLSimLightSourceToStepConverterUtils.h) compiled unmodified (oracle/ref_shim/ref_stepgen_utils.cpp).
# ---------------------------------------------------------------------------------------------------------
_REF_STEPGEN_LIB = os.path.join(_HERE, "_ref", "libclsim_ref_stepgen.so")
_ref_stepgen_lib = None


def ref_stepgen_available():
    return os.path.isfile(_REF_STEPGEN_LIB)


def ref_stepgen_lib():
    global _ref_stepgen_lib
    if _ref_stepgen_lib is None:
        L = C.CDLL(_REF_STEPGEN_LIB)
        for name in ("ref_mwc_co", "ref_mwc_oc"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.POINTER(C.c_uint64), C.c_uint32]
        L.ref_gamma_distributed.restype = C.c_double
        L.ref_gamma_distributed.argtypes = [C.c_double, C.POINTER(C.c_uint64), C.c_uint32]
        L.ref_scatter_direction_by_angle.restype = None
        L.ref_scatter_direction_by_angle.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.c_double]
        L.ref_mwc_init_state.restype = C.c_uint64
        L.ref_mwc_init_state.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_size_t)]
        _ref_stepgen_lib = L
    return _ref_stepgen_lib


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref/libclsim_ref_medium.so: the reference's medium / wavelength SOURCE GENERATORS (function, random-value
# and medium-properties classes plus I3CLSimHelperGenerateMediumPropertiesSource{,_Optimizers}.cxx) compiled
# unmodified (oracle/ref_shim/ref_medium.cpp).  Returns the OpenCL text the reference would generate.
# ---------------------------------------------------------------------------------------------------------
_REF_MEDIUM_LIB = os.path.join(_HERE, "_ref", "libclsim_ref_medium.so")
_ref_medium_lib = None


def ref_medium_available():
    return os.path.isfile(_REF_MEDIUM_LIB)


def ref_medium_lib():
    global _ref_medium_lib
    if _ref_medium_lib is None:
        L = C.CDLL(_REF_MEDIUM_LIB)
        L.ref_medium_last_error.restype = C.c_char_p
        L.ref_medium_source.restype = C.c_int64
        L.ref_medium_source.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
        L.ref_wlen_generator_source.restype = C.c_int64
        L.ref_wlen_generator_source.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t]
        L.ref_wlen_bias_source.restype = C.c_int64
        L.ref_wlen_bias_source.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.ref_medium_host_value.restype = C.c_double
        L.ref_medium_host_value.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double]
        _ref_medium_lib = L
    return _ref_medium_lib


def _text_of(call):
    L = ref_medium_lib()
    n = call(None, 0)
    if n < 0:
        raise RuntimeError(L.ref_medium_last_error().decode())
    buf = C.create_string_buffer(n + 1)
    call(buf, n + 1)
    return buf.value.decode()


class RefGeneratedSource(object):
    """The three generated snippets of the reference's OpenCL program for (medium, wavelength generators, bias):
    .medium, .wlen_generators, .wlen_bias -- text, exactly as the reference's generators write it."""

    def __init__(self, medium, wlen_generators, wlen_bias):
        from clsim_b200.description import ConverterOptions
        self._cfg, self._keep = build_config(medium, None, wlen_generators, wlen_bias, ConverterOptions())
        L = ref_medium_lib()
        m = C.byref(self._cfg.medium)
        self._tilt_z = None if medium.tilt is None else np.ascontiguousarray(medium.tilt["zCoordinates"], dtype=np.float64)
        tz = None if self._tilt_z is None else self._tilt_z.ctypes.data
        self.medium = _text_of(lambda out, cap: L.ref_medium_source(m, tz, out, cap))
        self.wlen_generators = _text_of(lambda out, cap: L.ref_wlen_generator_source(self._cfg.wlen_generators, self._cfg.num_wlen_generators, out, cap))
        self.wlen_bias = _text_of(lambda out, cap: L.ref_wlen_bias_source(C.byref(self._cfg.wlen_bias), out, cap))

    def host_value(self, what, layer=0, a=0.0, b=0.0, c=0.0):
        """GetValue of the reference's host-side class: what = 0 phase index, 1 group index, 2 scattering length,
        3 absorption length (a = wavelength), 4 tilt shift, 5 absorption anisotropy factor (a, b, c = vector), 6 / 7 min / max wavelength."""
        tz = None if self._tilt_z is None else self._tilt_z.ctypes.data
        return ref_medium_lib().ref_medium_host_value(C.byref(self._cfg.medium), tz, what, layer, a, b, c)

    @staticmethod
    def preamble(options):
        L = ref_medium_lib()
        L.ref_preamble_source.restype = C.c_int64
        L.ref_preamble_source.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_char_p, C.c_size_t]
        o = options
        return _text_of(lambda out, cap: L.ref_preamble_source(int(bool(o.stop_detected_photons)), int(bool(o.save_all_photons)),
                                                               float(o.save_all_photons_prescale), int(o.photon_history_entries),
                                                               float(o.fixed_number_of_absorption_lengths), float(o.pancake_factor), out, cap))


_REF_KERNEL_DIR = "/root/reference/resources/kernels"


def ref_program_available():
    """A whole program of the reference can be assembled and compiled where its kernel text lies (this container)."""
    return ref_medium_available() and ref_geometry_available() and os.path.isfile(os.path.join(_REF_KERNEL_DIR, "propagation_kernel.c.cl"))


class RefProgram(object):
    """ONE complete OpenCL program of the reference for (medium, geometry, generators, bias, options), every byte of it
    text the reference ships (resources/kernels/*.cl) or writes with its own generators (libclsim_ref_medium.so,
    libclsim_ref_geometry.so), in the order I3CLSimStepToPhotonConverterOpenCL.cxx:655-667 joins them, compiled for the
    host under oracle/ref_shim/ref_program.cpp.  Nothing in it comes from the oracle's restatements.  Compiled
    libraries are cached under oracle/_ref/programs/ by the hash of the program text and of the shim."""

    def __init__(self, medium, geometry, wlen_generators, wlen_bias, options, save_all_dom_stub=False):
        """save_all_dom_stub: in SAVE_ALL_PHOTONS mode the reference joins no geometry source although saveHit names
        geometryGetDomPosition -- the program does not compile (RuntimeError with the compiler's message).  With the
        stub, a one-line stand-in of that function (a DOM at the origin; the oracle's reading) is put where the geometry
        source would be, so that the rest of the program can still be run in that mode."""
        import hashlib
        import sys
        import tempfile
        sys.path.insert(0, os.path.join(_HERE, "ref_shim"))
        try:
            import translate
        finally:
            sys.path.pop(0)
        self.options = options
        self.generated = RefGeneratedSource(medium, wlen_generators, wlen_bias)

        def kernel(name):
            with open(os.path.join(_REF_KERNEL_DIR, name)) as f:
                return f.read()

        parts = [RefGeneratedSource.preamble(options), kernel("mwcrng_kernel.cl"), self.generated.wlen_generators, self.generated.wlen_bias,
                 self.generated.medium]
        self.layer_to_om = np.zeros(1, np.uint16)
        self.string_ids, self.dom_ids = [], []
        if not options.save_all_photons:
            geo_text, self.layer_to_om, self.string_ids, self.dom_ids = ref_geometry_source(geometry)
            parts.append(geo_text)
        elif save_all_dom_stub:
            parts.append("inline void geometryGetDomPosition(unsigned short stringNum, unsigned short domNum, floating_t *domPosX, "
                         "floating_t *domPosY, floating_t *domPosZ) { *domPosX = ZERO; *domPosY = ZERO; *domPosZ = ZERO; }   // NOT reference text\n")
        # :522-527: header, collision detection (header, body), kernel body
        tail = kernel("propagation_kernel.h.cl")
        if not options.save_all_photons:
            tail += kernel("sparse_collision_kernel.h.cl") + kernel("sparse_collision_kernel.c.cl")
        tail += kernel("propagation_kernel.c.cl")
        parts.append(tail)
        self.text = "".join(p + "\n" for p in parts)
        flags = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w"]
        if len(wlen_generators) <= 1:
            flags.append("-DNO_FLASHER")        # :649
        shim = b""
        for name in ("ref_program.cpp", "opencl_c_shim.inc", "opencl_c_shim_generated.inc", "translate.py"):
            with open(os.path.join(_HERE, "ref_shim", name), "rb") as f:
                shim += f.read()
        key = hashlib.sha256(self.text.encode() + shim + " ".join(flags).encode()).hexdigest()[:24]
        cache = os.path.join(_HERE, "_ref", "programs")
        so = os.path.join(cache, key + ".so")
        if not os.path.isfile(so):
            os.makedirs(cache, exist_ok=True)
            with tempfile.TemporaryDirectory() as tmp:
                with open(os.path.join(tmp, "program.cl.inc"), "w") as f:
                    f.write(translate.translate(self.text))
                out = os.path.join(tmp, "prog.so")
                cc = subprocess.run(["g++"] + flags + ["-I" + tmp, "-I" + os.path.join(_HERE, "ref_shim"), "-shared", "-o", out,
                                     os.path.join(_HERE, "ref_shim", "ref_program.cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                if cc.returncode != 0:
                    raise RuntimeError("the reference's program does not compile:\n" + cc.stdout.decode(errors="replace")[-4000:])
                os.replace(out, so)
        L = C.CDLL(so)
        L.prog_eval_wlen_function.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.prog_eval_scalar_field.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.prog_eval_vector_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.prog_sample.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p, C.c_size_t]
        L.prog_propagate.restype = C.c_uint32
        L.prog_propagate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        self._lib = L

    # ---- the generated functions, one by one (same conventions as Scene.eval_* / Scene.sample)
    def eval_wlen_function(self, which, layers, wlens):
        arr = np.ascontiguousarray(np.stack([np.asarray(layers, dtype=np.float32), np.asarray(wlens, dtype=np.float32)], axis=-1))
        out = np.zeros(len(arr), dtype=np.float32)
        self._lib.prog_eval_wlen_function(which, arr.ctypes.data, out.ctypes.data, len(arr))
        return out

    def eval_scalar_field(self, which, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        out = np.zeros(len(xyz), dtype=np.float32)
        self._lib.prog_eval_scalar_field(which, xyz.ctypes.data, out.ctypes.data, len(xyz))
        return out

    def eval_vector_transform(self, which, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        out = np.zeros_like(xyz)
        self._lib.prog_eval_vector_transform(which, xyz.ctypes.data, out.ctypes.data, len(xyz))
        return out

    def sample(self, which, x, a, n):
        out = np.zeros(n, dtype=np.float32)
        xs = C.c_uint64(int(x))
        self._lib.prog_sample(which, C.byref(xs), int(a), out.ctypes.data, n)
        return out, xs.value

    # ---- the kernel: one launch over the steps in work-item order
    def propagate(self, steps, rng_x, rng_a, cap=None):
        """-> (hit records with string / DOM IDs, hit counter, final RNG states, history or None); the contract of
        Scene.propagate with num_threads = 1."""
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        n = len(steps)
        x = np.array(rng_x[:n], dtype=np.uint64)
        a = np.array(rng_a[:n], dtype=np.uint32)
        if cap is None:
            cap = int(steps["num_photons"].sum()) + 16
        out = np.zeros(cap, dtype=PHOTON_DTYPE)
        H = int(self.options.photon_history_entries)
        hist = np.zeros((cap, H, 4), dtype=np.float32) if H > 0 else None
        count = self._lib.prog_propagate(steps.ctypes.data, n, x.ctypes.data, a.ctypes.data, out.ctypes.data, cap,
                                         self.layer_to_om.ctypes.data, hist.ctypes.data if hist is not None else None)
        got = out[:min(count, cap)]
        if not self.options.save_all_photons:
            # I3CLSimStepToPhotonConverterOpenCL.cxx:1565-1602: indices -> IDs on the host after the launch
            s = got["string_id"].astype(np.uint16).astype(np.int64)
            d = got["om_id"].astype(np.int64)
            sid = np.array([self.string_ids[i] for i in s], dtype=np.int16)
            did = np.array([self.dom_ids[i][j] for i, j in zip(s, d)], dtype=np.uint16)
            got["string_id"], got["om_id"] = sid, did
        return got, int(count), x, (hist[:len(got)] if hist is not None else None)


class RefTableProgram(object):
    """ONE complete TABLE-MAKER program of the reference (private/clsim/tabulator/I3CLSimStepToTableConverter.cxx:178-212):
    preamble with -DTABULATE, mwcrng_kernel.cl, generated wavelength generator / bias / medium / angular acceptance,
    propagation_kernel.h.cl, the binning code Axes::GenerateBinningCode writes (with the coordinate kernel it loads) and
    propagation_kernel.c.cl, compiled for the host under oracle/ref_shim/ref_table_program.cpp.

    At this revision saveHit (compiled, never called with TABULATE) names geometryGetDomPosition, which no part of the
    table-maker's program declares: as in save-all mode the text is ill-formed without a declaration of that function.
    dom_stub=True adds the one-line stand-in (NOT reference text); dom_stub=False shows the compiler's message."""

    def __init__(self, medium, wlen_generators, wlen_bias, axes, angular_coefficients, entries_per_stream=4096, step_length=1.0, dom_stub=True):
        import hashlib
        import sys
        import tempfile
        sys.path.insert(0, os.path.join(_HERE, "ref_shim"))
        try:
            import translate
        finally:
            sys.path.pop(0)
        L = ref_medium_lib()
        self.generated = RefGeneratedSource(medium, wlen_generators, wlen_bias)
        self.axes = axes
        n = len(axes.axes)
        kind = (C.c_int32 * n)(*[ax.kind for ax in axes.axes])
        power = (C.c_uint32 * n)(*[ax.power for ax in axes.axes])
        bins = (C.c_uint32 * n)(*[ax.n_bins for ax in axes.axes])
        lo = (C.c_double * n)(*[ax.min for ax in axes.axes])
        hi = (C.c_double * n)(*[ax.max for ax in axes.axes])
        nb = C.c_uint64(0)
        L.ref_binning_source.restype = C.c_int64
        L.ref_angular_acceptance_source.restype = C.c_int64
        L.ref_table_preamble_source.restype = C.c_int64
        L.ref_table_preamble_source.argtypes = [C.c_int32, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_char_p, C.c_size_t]
        with tempfile.TemporaryDirectory() as build_root:
            # Axes.cxx:36-43 reads $I3_BUILD/clsim/resources/kernels/<name>.c.cl
            os.symlink("/root/reference", os.path.join(build_root, "clsim"))
            old = os.environ.get("I3_BUILD")
            os.environ["I3_BUILD"] = build_root
            try:
                self.binning = _text_of(lambda out, cap: L.ref_binning_source(int(axes.geometry), n, kind, power, bins, lo, hi, C.byref(nb), out, C.c_size_t(cap)))
            finally:
                if old is None:
                    del os.environ["I3_BUILD"]
                else:
                    os.environ["I3_BUILD"] = old
        self.num_bins = int(nb.value)
        idx = (C.c_double * 2)()
        if L.ref_minimum_refractive_index(C.byref(self.generated._cfg.medium), idx) != 0:
            raise RuntimeError(L.ref_medium_last_error().decode())
        self.n_group, self.n_phase = float(idx[0]), float(idx[1])
        coef = np.ascontiguousarray(angular_coefficients, dtype=np.float64)
        self.angular = _text_of(lambda out, cap: L.ref_angular_acceptance_source(coef.ctypes.data_as(C.POINTER(C.c_double)), len(coef), out, C.c_size_t(cap)))
        self.preamble = _text_of(lambda out, cap: L.ref_table_preamble_source(n, int(entries_per_stream), float(step_length), self.n_group, self.n_phase, out, cap))

        def kernel(name):
            with open(os.path.join(_REF_KERNEL_DIR, name)) as f:
                return f.read()

        parts = [self.preamble, kernel("mwcrng_kernel.cl"), self.generated.wlen_generators, self.generated.wlen_bias, self.generated.medium, self.angular]
        if dom_stub:
            parts.append("inline void geometryGetDomPosition(unsigned short stringNum, unsigned short domNum, floating_t *domPosX, "
                         "floating_t *domPosY, floating_t *domPosZ) { *domPosX = ZERO; *domPosY = ZERO; *domPosZ = ZERO; }   // NOT reference text\n")
        parts += [kernel("propagation_kernel.h.cl"), self.binning, kernel("propagation_kernel.c.cl")]
        self.text = "".join(parts)      # (the table-maker hands the parts over as separate strings: no separators)
        flags = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w"]
        shim = b""
        for name in ("ref_table_program.cpp", "opencl_c_shim.inc", "opencl_c_shim_generated.inc", "translate.py"):
            with open(os.path.join(_HERE, "ref_shim", name), "rb") as f:
                shim += f.read()
        key = hashlib.sha256(self.text.encode() + shim + " ".join(flags).encode()).hexdigest()[:24]
        cache = os.path.join(_HERE, "_ref", "programs")
        so = os.path.join(cache, key + ".so")
        if not os.path.isfile(so):
            os.makedirs(cache, exist_ok=True)
            with tempfile.TemporaryDirectory() as tmp:
                with open(os.path.join(tmp, "program.cl.inc"), "w") as f:
                    f.write(translate.translate(self.text))
                out = os.path.join(tmp, "prog.so")
                cc = subprocess.run(["g++"] + flags + ["-I" + tmp, "-I" + os.path.join(_HERE, "ref_shim"), "-shared", "-o", out,
                                     os.path.join(_HERE, "ref_shim", "ref_table_program.cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                if cc.returncode != 0:
                    raise RuntimeError("the reference's program does not compile:\n" + cc.stdout.decode(errors="replace")[-4000:])
                os.replace(out, so)
        self._lib = C.CDLL(so)
        self._lib.prog_tabulate.restype = C.c_uint64
        self._lib.prog_tabulate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]

    @staticmethod
    def reference_particle(reference):
        """I3CLSimReferenceParticle (…StepToTableConverter.cxx:64-91) from (x, y, z, t, dx, dy, dz): twelve floats."""
        x, y, z, t, dx, dy, dz = [float(v) for v in reference]
        perpz = float(np.hypot(dx, dy))
        if perpz > 0.:
            p = np.array([-dx * dz / perpz, -dy * dz / perpz, perpz])
            p = p / np.sqrt((p * p).sum())       # I3Direction normalises
        else:
            p = np.array([1., 0., 0.])
        return np.array([x, y, z, t, dx, dy, dz, 0., p[0], p[1], p[2], 0.], dtype=np.float32)

    def tabulate(self, steps, rng_x, rng_a, reference, squared=False):
        """-> (bins[float64], squared or None, entries, new rng_x, launches): the contract of Scene.tabulate."""
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        n = len(steps)
        x = np.array(rng_x[:n], dtype=np.uint64, copy=True)
        a = np.array(rng_a[:n], dtype=np.uint32, copy=True)
        bins = np.zeros(self.num_bins, dtype=np.float64)
        sq = np.zeros(self.num_bins, dtype=np.float64) if squared else None
        ref = self.reference_particle(reference)
        launches = C.c_uint64(0)
        entries = self._lib.prog_tabulate(steps.ctypes.data, n, x.ctypes.data, a.ctypes.data, ref.ctypes.data, bins.ctypes.data,
                                          sq.ctypes.data if squared else None, C.byref(launches))
        if entries == 0xFFFFFFFFFFFFFFFF:
            raise RuntimeError("a single photon needs more than TABLE_ENTRIES_PER_STREAM entries")
        return bins, sq, int(entries), x, int(launches.value)


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref/libclsim_ref_wire.so: the reference's step / photon records (public/clsim/I3CLSimStep.h, I3CLSimPhoton.h)
# and their serialize() members (private/clsim/I3CLSimStep.cxx, I3CLSimPhoton.cxx) compiled unmodified
# (oracle/ref_shim/ref_wire.cpp; the archives are stand-ins, see oracle/ref_shim/host_wire/icetray/serialization.h).
# ---------------------------------------------------------------------------------------------------------
_REF_WIRE_LIB = os.path.join(_HERE, "_ref", "libclsim_ref_wire.so")
_ref_wire_lib = None


def ref_wire_available():
    return os.path.isfile(_REF_WIRE_LIB)


def ref_wire_lib():
    global _ref_wire_lib
    if _ref_wire_lib is None:
        L = C.CDLL(_REF_WIRE_LIB)
        L.ref_wire_error.restype = C.c_char_p
        for name in ("ref_wire_step_size", "ref_wire_photon_size", "ref_wire_step_version", "ref_wire_photon_version"):
            getattr(L, name).restype = C.c_uint32
        L.ref_wire_make_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_wire_make_photon.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_wire_read_step_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("ref_wire_write_steps", "ref_wire_write_photons"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        for name in ("ref_wire_read_steps", "ref_wire_read_photons"):
            getattr(L, name).restype = C.c_int64
            getattr(L, name).argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        _ref_wire_lib = L
    return _ref_wire_lib


def ref_wire_write(records):
    """Body the reference's I3Vector<...>::serialize(portable_binary_oarchive) writes for a step or photon series."""
    L = ref_wire_lib()
    records = np.ascontiguousarray(records)
    fn = L.ref_wire_write_steps if records.dtype.itemsize == 48 else L.ref_wire_write_photons
    p = C.c_void_p()
    n = fn(records.ctypes.data if len(records) else None, len(records), C.byref(p))
    return C.string_at(p, n)


def ref_wire_read(body, dtype):
    """Records the reference's I3Vector<...>::serialize(portable_binary_iarchive) reads from a body, and the bytes it left
    unread.  RuntimeError with the reference's message when it refuses."""
    L = ref_wire_lib()
    fn = L.ref_wire_read_steps if np.dtype(dtype).itemsize == 48 else L.ref_wire_read_photons
    left = C.c_uint64(0)
    n = fn(bytes(body), len(body), None, 0, C.byref(left))
    if n < 0:
        raise RuntimeError(L.ref_wire_error().decode())
    out = np.zeros(n, dtype=dtype)
    fn(bytes(body), len(body), out.ctypes.data, n, C.byref(left))
    return out, int(left.value)


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref/libclsim_ref_rng.so: init_MWC_RNG of private/opencl/mwcrng_init.h compiled unmodified (oracle/ref_shim/ref_rng_init.cpp)
# ---------------------------------------------------------------------------------------------------------
_REF_RNG_LIB = os.path.join(_HERE, "_ref", "libclsim_ref_rng.so")


def ref_rng_available():
    return os.path.isfile(_REF_RNG_LIB)


def ref_init_mwc_rng(n, safeprimes_file, values):
    """-> (x, a, number of 64-bit values used): the reference's init_MWC_RNG reading `safeprimes_file`, its random service
    handing out the upper, then the lower 32 bits of each of `values`.  RuntimeError when it reports failure."""
    L = C.CDLL(_REF_RNG_LIB)
    L.ref_init_mwc_rng.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    values = np.ascontiguousarray(values, dtype=np.uint64)
    x, a = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32)
    used = C.c_size_t(0)
    rc = L.ref_init_mwc_rng(x.ctypes.data, a.ctypes.data, n, safeprimes_file.encode(), values.ctypes.data, len(values), C.byref(used))
    if rc != 0:
        raise RuntimeError("init_MWC_RNG returned %d" % rc)
    return x, a, int(used.value)


def ref_made_wlen_generator_source(medium, wlen_bias, without_dispersion=False, spectrum=None):
    """Text of the wavelength generator the REFERENCE'S OWN FACTORY makes for (bias, medium): makeCherenkovWavelengthGenerator, or
    makeWavelengthGenerator for a tabulated spectrum = (wavelengths, values) (private/clsim/I3CLSimModuleHelper.cxx:75-300,
    compiled unmodified into libclsim_ref_medium.so)."""
    from clsim_b200.description import ConverterOptions
    from clsim_b200.description import WlenGenerator
    L = ref_medium_lib()
    cfg, keep = build_config(medium, None, [WlenGenerator.constant(4e-7)], wlen_bias, ConverterOptions())
    tz = None if medium.tilt is None else np.ascontiguousarray(medium.tilt["zCoordinates"], dtype=np.float64)
    sx = sy = None
    n = 0
    if spectrum is not None:
        sx, sy = [np.ascontiguousarray(v, dtype=np.float64) for v in spectrum]
        n = len(sx)
    L.ref_made_wlen_generator_source.restype = C.c_int64
    L.ref_made_wlen_generator_source.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t]
    text = _text_of(lambda out, cap: L.ref_made_wlen_generator_source(C.byref(cfg.medium), None if tz is None else tz.ctypes.data, C.byref(cfg.wlen_bias),
                                                                     int(bool(without_dispersion)), None if sx is None else sx.ctypes.data,
                                                                     None if sy is None else sy.ctypes.data, n, out, cap))
    del keep
    return text


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref/libclsim_ref_mcpe.so: the reference's photon -> photo-electron converters
# (private/clsim/dom/I3PhotonToMCPEConverter.cxx) compiled unmodified (oracle/ref_shim/ref_mcpe.cpp)
# ---------------------------------------------------------------------------------------------------------
_REF_MCPE_LIB = os.path.join(_HERE, "_ref", "libclsim_ref_mcpe.so")
_ref_mcpe_lib = None


def ref_mcpe_available():
    return os.path.isfile(_REF_MCPE_LIB)


def _ref_mcpe():
    global _ref_mcpe_lib
    if _ref_mcpe_lib is None:
        L = C.CDLL(_REF_MCPE_LIB)
        L.ref_mcpe_error.restype = C.c_char_p
        L.ref_mcpe_convert_inloop.restype = C.c_int64
        L.ref_mcpe_convert_inloop.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.ref_mcpe_convert_module.restype = C.c_int64
        L.ref_mcpe_convert_module.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double,
                                              C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_double,
                                              C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                              C.POINTER(C.c_uint64)]
        _ref_mcpe_lib = L
    return _ref_mcpe_lib


def _acceptance_args(acc):
    if acc.values is None:
        return None, None, 0, 0.0, 0.0, float(acc.constant)
    v = np.ascontiguousarray(acc.values, dtype=np.float64)
    return v, v.ctypes.data, len(v), float(acc.start_wlen), float(acc.wlen_step), 0.0


def ref_mcpe_convert_inloop(photons, acceptance, angular_coefficients, uniforms):
    """I3CLSimPhotonToMCPEConverterForDOMs::Convert of the reference, photon by photon (photon i is offered uniforms[i]).
    -> (survivor mask, times of the survivors' photo-electrons (0 elsewhere), uniforms drawn).  RuntimeError = log_fatal."""
    L = _ref_mcpe()
    photons = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
    n = len(photons)
    keep_v, vp, vn, x0, dx, const = _acceptance_args(acceptance)
    ang = np.ascontiguousarray(angular_coefficients, dtype=np.float64)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    survive, t = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.float64)
    used = C.c_uint64(0)
    rc = L.ref_mcpe_convert_inloop(photons.ctypes.data, n, vp, vn, x0, dx, const, ang.ctypes.data, len(ang), u.ctypes.data, survive.ctypes.data,
                                   t.ctypes.data, C.byref(used))
    if rc < 0:
        raise RuntimeError(L.ref_mcpe_error().decode())
    return survive.astype(bool), t, int(used.value)


def ref_mcpe_convert_module(photons, dom_positions, acceptance, angular_coefficients, efficiency, uniforms_in_order, oversize=1.0, pancake=1.0,
                            dom_radius=0.1651, default_efficiency=1.0, replace_with_default=False, only_warn=False):
    """The reference's I3PhotonToMCPEConverter MODULE on one frame.  photons: DOM-relative records; dom_positions[n, 3]: where
    each photon's DOM sits; efficiency[n]: relative DOM efficiency from the calibration for each photon's DOM (NaN = no entry).
    uniforms_in_order: handed out as the module asks (DOMs in key order, photons of a DOM in input order, weight 0 skipped).
    -> (string[k], om[k], time[k], input photon index[k]) in the module's output order, and the number of uniforms drawn."""
    L = _ref_mcpe()
    photons = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
    n = len(photons)
    keep_v, vp, vn, x0, dx, const = _acceptance_args(acceptance)
    ang = np.ascontiguousarray(angular_coefficients, dtype=np.float64)
    u = np.ascontiguousarray(uniforms_in_order, dtype=np.float64)
    pos = np.ascontiguousarray(dom_positions, dtype=np.float64)
    eff = np.ascontiguousarray(efficiency, dtype=np.float64)
    s, o, t, idx = np.zeros(n, np.int32), np.zeros(n, np.uint32), np.zeros(n, np.float64), np.zeros(n, np.int64)
    used = C.c_uint64(0)
    k = L.ref_mcpe_convert_module(photons.ctypes.data, n, pos.ctypes.data, vp, vn, x0, dx, const, ang.ctypes.data, len(ang), eff.ctypes.data,
                                  float(default_efficiency), int(bool(replace_with_default)), float(oversize), float(pancake), float(dom_radius),
                                  int(bool(only_warn)), u.ctypes.data, len(u), s.ctypes.data, o.ctypes.data, t.ctypes.data, idx.ctypes.data, n,
                                  C.byref(used))
    if k < 0:
        raise RuntimeError(L.ref_mcpe_error().decode())
    return s[:k], o[:k], t[:k], idx[:k], int(used.value)


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref/libclsim_icetray_mode.so: the PRODUCT's converter class (clsim_b200/host/I3CLSimStepToPhotonConverterCUDA.cxx)
# compiled with -DCLSIM_CUDA_IN_ICETRAY against the reference's own public headers (+ the getters of INTEGRATION.md section 2)
# and linked with the reference's description-class sources (oracle/ref_shim/ref_icetray_mode.cpp)
# ---------------------------------------------------------------------------------------------------------
_ICETRAY_MODE_LIB = os.path.join(_HERE, "_ref", "libclsim_icetray_mode.so")


def icetray_mode_available():
    return os.path.isfile(_ICETRAY_MODE_LIB)


def icetray_mode_describe_tables(medium, geometry, wlen_generators, wlen_bias, options):
    """JSON of the device tables after SetWlenGenerators / SetWlenBias / SetMediumProperties / SetGeometry / Compile() on the
    product's converter class, fed REFERENCE objects through the reference's abstract interface."""
    L = C.CDLL(_ICETRAY_MODE_LIB)
    L.icetray_mode_error.restype = C.c_char_p
    L.icetray_mode_describe_tables.restype = C.c_int64
    L.icetray_mode_describe_tables.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    cfg, keep = build_config(medium, geometry, wlen_generators, wlen_bias, options)
    tz = None if medium.tilt is None else np.ascontiguousarray(medium.tilt["zCoordinates"], dtype=np.float64)
    names = (C.c_char_p * max(1, len(geometry.subdetectorNames)))(*[s.encode() for s in geometry.subdetectorNames])
    out = C.c_char_p()
    n = L.icetray_mode_describe_tables(C.byref(cfg), None if tz is None else tz.ctypes.data, names, C.byref(out))
    if n < 0:
        raise RuntimeError(L.icetray_mode_error().decode())
    text = C.string_at(out, n).decode()
    del keep
    return json.loads(text)


def icetray_mode_unknown_class_message(medium, wlen_generators, wlen_bias):
    """What the class throws (as the reference's I3CLSimStepToPhotonConverter_exception) for a scattering model it does not know."""
    from clsim_b200.description import ConverterOptions
    L = C.CDLL(_ICETRAY_MODE_LIB)
    L.icetray_mode_error.restype = C.c_char_p
    cfg, keep = build_config(medium, None, wlen_generators, wlen_bias, ConverterOptions())
    rc = L.icetray_mode_unknown_class_is_refused(C.byref(cfg))
    del keep
    return rc, L.icetray_mode_error().decode()


def ref_medium_host_values(generated, what, abc, layer=0):
    """GetValue / ApplyTransform of the reference's host-side classes for many arguments at once: `generated` is a
    RefGeneratedSource; what as in RefGeneratedSource.host_value, plus 8 / 9 = pre / post scattering direction transform
    (three values out per triple)."""
    L = ref_medium_lib()
    abc = np.ascontiguousarray(abc, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((len(abc), 3) if what in (8, 9) else len(abc), dtype=np.float64)
    tz = None if generated._tilt_z is None else generated._tilt_z.ctypes.data
    L.ref_medium_host_values.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64]
    if L.ref_medium_host_values(C.byref(generated._cfg.medium), tz, what, layer, abc.ctypes.data, out.ctypes.data, len(abc)) != 0:
        raise RuntimeError(L.ref_medium_last_error().decode())
    return out
