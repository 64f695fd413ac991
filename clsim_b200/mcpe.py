"""Photon -> MCPE conversion on the device: Python mirror of the reference's two converters over the C ABI
(include/clsimcuda.h, "photon -> MCPE on the device"; kernels in csrc/mcpe.cu).

* ``I3CLSimPhotonToMCPEConverterForDOMs`` -- private/clsim/dom/I3PhotonToMCPEConverter.cxx:595-669, the in-loop
  converter I3CLSimClientModule runs on every detected photon (…ClientModule.cxx:424-433).
* ``I3PhotonToMCPEConverter`` -- the per-photon part of the stand-alone module (…cxx:395-523) with the module's
  parameter names (…cxx:40-140).
* ``GetIceCubeDOMAngularSensitivity`` -- python/GetIceCubeDOMAngularSensitivity.py:30-42.

There is no CPU path: the classes raise when the library or a CUDA device is missing.
"""
import ctypes as C
import json
import math
import os

import numpy as np

from . import capi
from .description import PHOTON_DTYPE, WlenBias, WlenBiasStruct

MCPE_DTYPE = np.dtype([("string_id", "<i2"), ("om_id", "<u2"), ("time", "<f4"), ("npe", "<u4"), ("identifier", "<u4")])
assert MCPE_DTYPE.itemsize == 16

FLAVOUR_INLOOP, FLAVOUR_MODULE = 0, 1
DOM_RADIUS = 0.16510

_HERE = os.path.dirname(os.path.abspath(__file__))


class McpeConfigStruct(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("flavour", C.c_int32), ("num_acceptances", C.c_int32),
        ("acceptances", C.POINTER(WlenBiasStruct)),
        ("num_doms", C.c_int32), ("num_angular_coefficients", C.c_int32),
        ("string_id", C.POINTER(C.c_int32)), ("dom_id", C.POINTER(C.c_uint32)),
        ("acceptance_of_dom", C.POINTER(C.c_uint8)), ("efficiency_of_dom", C.POINTER(C.c_double)),
        ("angular_coefficients", C.POINTER(C.c_double)),
        ("dom_dir", C.c_double * 3), ("dom_radius", C.c_double), ("oversize_factor", C.c_double), ("pancake_factor", C.c_double),
        ("only_warn_about_positions", C.c_int32), ("reserved0", C.c_int32),
        ("rng_seed", C.c_uint64), ("rng_first_multiplier", C.c_uint64),
    ]


class I3CLSimFunctionPolynomial(object):
    """public/clsim/function/I3CLSimFunctionPolynomial.h, the range-less constructor."""

    def __init__(self, coefficients):
        self.coefficients = [float(c) for c in coefficients]

    def GetValue(self, x):
        # private/clsim/function/I3CLSimFunctionPolynomial.cxx:86-102
        if not self.coefficients:
            return 0.0
        total, multiplier = self.coefficients[0], 1.0
        for c in self.coefficients[1:]:
            multiplier *= x
            total += c * multiplier
        return total


def _packaged_angsens():
    with open(os.path.join(_HERE, "data", "ice_models.json")) as f:
        return json.load(f)["_angsens_holeice"]


def GetIceCubeDOMAngularSensitivity(holeIce=None):
    """Relative collection efficiency of the DOM as a polynomial in the cosine of the impact angle.  `holeIce`
    names a parameterisation file (first row: peak value, then the coefficients); None = the hole-ice curve the
    reference's tree ships (resources/ice/ppc_aha_0.80/as.holeice)."""
    if holeIce is None:
        return I3CLSimFunctionPolynomial(_packaged_angsens()["coefficients"])
    return I3CLSimFunctionPolynomial(np.loadtxt(holeIce)[1:].tolist())


def GetHoleIcePeak(holeIce=None):
    """Row 0 of the parameterisation file: the value the tray segment folds into the DOM efficiency
    (python/traysegments/common.py:183-184)."""
    if holeIce is None:
        return float(_packaged_angsens()["peak"])
    return float(np.loadtxt(holeIce)[0])


def _lib():
    L = capi.lib()
    if not getattr(L, "_mcpe_bound", False):
        L.clsimcu_mcpe_create.argtypes = [C.POINTER(McpeConfigStruct), C.POINTER(C.c_void_p)]
        L.clsimcu_mcpe_destroy.argtypes = [C.c_void_p]
        L.clsimcu_mcpe_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_mcpe_rng_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_attach_mcpe_converter.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L._mcpe_bound = True
    return L


class _DeviceConverter(object):
    """Owns a clsimcu_mcpe_converter."""

    def __init__(self, flavour, acceptances, dom_keys, acceptance_of_dom, efficiency_of_dom, angular, device=0, dom_dir=(0.0, 0.0, -1.0),
                 dom_radius=DOM_RADIUS, oversize=1.0, pancake=1.0, only_warn=False, rng_seed=0, rng_first_multiplier=0):
        self._h = C.c_void_p()
        cfg = McpeConfigStruct()
        cfg.struct_size = C.sizeof(McpeConfigStruct)
        cfg.device, cfg.flavour = int(device), int(flavour)
        self._keep = []
        acc = (WlenBiasStruct * max(1, len(acceptances)))()
        for i, f in enumerate(acceptances):
            if not isinstance(f, WlenBias):
                raise TypeError("wavelength acceptance must be a FromTable (equal spacing) or Constant function, got %s" % type(f).__name__)
            if f.values is None:
                acc[i].kind, acc[i].value = 0, float(f.constant)
            else:
                v = np.ascontiguousarray(f.values, dtype=np.float64)
                self._keep.append(v)
                acc[i].kind, acc[i].n, acc[i].x0, acc[i].dx = 1, len(v), f.start_wlen, f.wlen_step
                acc[i].v = v.ctypes.data_as(C.POINTER(C.c_double))
        self._keep.append(acc)
        cfg.num_acceptances, cfg.acceptances = len(acceptances), acc
        sid = np.ascontiguousarray([k[0] for k in dom_keys], dtype=np.int32)
        did = np.ascontiguousarray([k[1] for k in dom_keys], dtype=np.uint32)
        aod = np.ascontiguousarray(acceptance_of_dom, dtype=np.uint8)
        coef = np.ascontiguousarray(angular.coefficients, dtype=np.float64)
        self._keep += [sid, did, aod, coef]
        cfg.num_doms, cfg.num_angular_coefficients = len(sid), len(coef)
        cfg.string_id = sid.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.dom_id = did.ctypes.data_as(C.POINTER(C.c_uint32))
        cfg.acceptance_of_dom = aod.ctypes.data_as(C.POINTER(C.c_uint8))
        if efficiency_of_dom is not None:
            eff = np.ascontiguousarray(efficiency_of_dom, dtype=np.float64)
            self._keep.append(eff)
            cfg.efficiency_of_dom = eff.ctypes.data_as(C.POINTER(C.c_double))
        cfg.angular_coefficients = coef.ctypes.data_as(C.POINTER(C.c_double))
        cfg.dom_dir[0], cfg.dom_dir[1], cfg.dom_dir[2] = [float(v) for v in dom_dir]
        cfg.dom_radius, cfg.oversize_factor, cfg.pancake_factor = float(dom_radius), float(oversize), float(pancake)
        cfg.only_warn_about_positions = int(bool(only_warn))
        cfg.rng_seed, cfg.rng_first_multiplier = int(rng_seed), int(rng_first_multiplier)
        capi._check(_lib().clsimcu_mcpe_create(C.byref(cfg), C.byref(self._h)))

    def convert(self, photons, uniforms=None):
        photons = np.ascontiguousarray(photons, dtype=PHOTON_DTYPE)
        n = len(photons)
        out = np.zeros(n, dtype=MCPE_DTYPE)
        u = None
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float32)
            if len(u) != n:
                raise ValueError("one uniform per photon")
        m = C.c_size_t(0)
        capi._check(_lib().clsimcu_mcpe_convert(self._h, photons.ctypes.data if n else None, n, u.ctypes.data if u is not None else None,
                                                out.ctypes.data if n else None, n, C.byref(m)))
        return out[:m.value]

    def rng_state(self):
        k = C.c_size_t(0)
        capi._check(_lib().clsimcu_mcpe_rng_get(self._h, None, None, 0, C.byref(k)))
        x = np.zeros(k.value, dtype=np.uint64)
        a = np.zeros(k.value, dtype=np.uint32)
        capi._check(_lib().clsimcu_mcpe_rng_get(self._h, x.ctypes.data, a.ctypes.data, k.value, C.byref(k)))
        return x, a

    def attach_to(self, engine, keep_photons=False):
        """Run the conversion on `engine`'s stream behind every propagation launch (results carry `.mcpes`)."""
        capi._check(_lib().clsimcu_attach_mcpe_converter(engine._h, self._h, int(bool(keep_photons))))
        engine._mcpe = self  # the converter must outlive the engine

    def close(self):
        if self._h:
            _lib().clsimcu_mcpe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _split_acceptance_map(wavelengthAcceptance):
    """{(string, om): function} -> (list of distinct functions, keys, index per key)."""
    funcs, keys, index = [], [], []
    for key in sorted(wavelengthAcceptance):
        f = wavelengthAcceptance[key]
        for i, g in enumerate(funcs):
            if g is f:
                break
        else:
            funcs.append(f)
            i = len(funcs) - 1
        keys.append((int(key[0]), int(key[1])))
        index.append(i)
    return funcs, keys, index


class I3CLSimPhotonToMCPEConverterForDOMs(_DeviceConverter):
    """I3CLSimPhotonToMCPEConverterForDOMs(randomService, {OMKey: wavelengthAcceptance}, angularAcceptance)
    (public/clsim/dom/I3PhotonToMCPEConverter.h:156-165).  `randomService` is a seed here: the thinning draws
    come from device MWC streams."""

    def __init__(self, randomService, wavelengthAcceptance, angularAcceptance, device=0, rngFirstMultiplierRow=0):
        if not wavelengthAcceptance:
            raise capi.ClsimCudaError(-1, "The \"WavelengthAcceptance\" parameter must not be empty.")
        funcs, keys, index = _split_acceptance_map(wavelengthAcceptance)
        _DeviceConverter.__init__(self, FLAVOUR_INLOOP, funcs, keys, index, None, angularAcceptance, device=device, rng_seed=int(randomService),
                                  rng_first_multiplier=rngFirstMultiplierRow)

    def Convert(self, photons, uniforms=None):
        """Photon series (PHOTON_DTYPE records, positions relative to their DOM) -> surviving MCPEs."""
        return self.convert(photons, uniforms)


class I3PhotonToMCPEConverter(_DeviceConverter):
    """Per-photon part of the I3PhotonToMCPEConverter module; parameter names as the module's (…cxx:40-140)."""

    def __init__(self, RandomService, Geometry, WavelengthAcceptance, AngularAcceptance, DOMOversizeFactor=1.0, DOMPancakeFactor=1.0,
                 DOMRadiusWithoutOversize=DOM_RADIUS, DefaultRelativeDOMEfficiency=1.0, RelativeDOMEfficiencies=None,
                 OnlyWarnAboutInvalidPhotonPositions=False, DOMDirection=(0.0, 0.0, -1.0), device=0, rngFirstMultiplierRow=0):
        if WavelengthAcceptance is None:
            raise capi.ClsimCudaError(-1, "The \"WavelengthAcceptance\" parameter must not be empty.")
        if AngularAcceptance is None:
            raise capi.ClsimCudaError(-1, "The \"AngularAcceptance\" parameter must not be empty.")
        if DefaultRelativeDOMEfficiency < 0.0:
            raise capi.ClsimCudaError(-1, "The \"DefaultRelativeDOMEfficiency\" parameter must not be < 0!")
        keys = [(int(s), int(d)) for s, d in zip(Geometry.stringIDs, Geometry.domIDs)]
        eff = []
        for k in keys:
            v = (RelativeDOMEfficiencies or {}).get(k, float("nan"))
            if math.isnan(v):
                if math.isnan(DefaultRelativeDOMEfficiency):
                    raise capi.ClsimCudaError(-1, "OM (%i/%u) not found in the current calibration map! (Consider setting \"DefaultRelativeDOMEfficiency\" != NaN)" % k)
                v = DefaultRelativeDOMEfficiency
            eff.append(v)
        _DeviceConverter.__init__(self, FLAVOUR_MODULE, [WavelengthAcceptance], keys, [0] * len(keys), eff, AngularAcceptance, device=device,
                                  dom_dir=DOMDirection, dom_radius=DOMRadiusWithoutOversize, oversize=DOMOversizeFactor, pancake=DOMPancakeFactor,
                                  only_warn=OnlyWarnAboutInvalidPhotonPositions, rng_seed=int(RandomService), rng_first_multiplier=rngFirstMultiplierRow)

    def Convert(self, photons, uniforms=None):
        return self.convert(photons, uniforms)


def sort_mcpes(m):
    """Per OM, by time (the module sorts each series, …cxx:526-527)."""
    return m[np.lexsort((m["identifier"], m["time"], m["om_id"], m["string_id"]))]
