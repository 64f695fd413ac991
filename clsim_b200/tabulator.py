"""Table-maker variant: Python mirror of the reference's tabulator classes over the C ABI
(include/clsimcuda.h, "table-maker variant"; host side csrc/tabulate.cu, kernel csrc/kernel_reference.cu).

* ``LinearAxis`` / ``PowerAxis``            private/clsim/tabulator/Axis.cxx
* ``SphericalAxes`` / ``CylindricalAxes``    private/clsim/tabulator/Axes.cxx
* ``I3CLSimStepToTableConverter``            private/clsim/tabulator/I3CLSimStepToTableConverter.cxx
  (EnqueueSteps(steps, reference), Finish, the normalised table and its header values; no FITS writer)

There is no CPU path: the converter raises when the library or a CUDA device is missing.
"""
import ctypes as C
import math

import numpy as np

from . import capi
from .description import KERNEL_FAST, KERNEL_REFERENCE, STEP_DTYPE, ConfigStruct, ConverterOptions, build_config

LINEAR, POWER = 0, 1
SPHERICAL, CYLINDRICAL = 0, 1


class AxisStruct(C.Structure):
    _fields_ = [("kind", C.c_int32), ("power", C.c_uint32), ("min", C.c_double), ("max", C.c_double), ("n_bins", C.c_uint32), ("reserved0", C.c_uint32)]


class TabulatorConfigStruct(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("geometry", C.c_int32), ("num_axes", C.c_int32), ("store_squared_weights", C.c_int32),
                ("axes", AxisStruct * 5), ("step_length", C.c_double), ("reference_area", C.c_double),
                ("num_angular_coefficients", C.c_int32), ("reserved0", C.c_int32), ("angular_coefficients", C.POINTER(C.c_double)),
                ("min_wavelength", C.c_double), ("max_wavelength", C.c_double)]


class ReferenceParticleStruct(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double), ("t", C.c_double),
                ("dir_x", C.c_double), ("dir_y", C.c_double), ("dir_z", C.c_double)]


class Axis(object):
    kind, power = LINEAR, 1

    def __init__(self, min, max, n_bins):
        self.min, self.max, self.n_bins = float(min), float(max), int(n_bins)

    def Transform(self, v):
        return v

    def InverseTransform(self, v):
        return v

    def GetNBins(self):
        return self.n_bins

    def GetMin(self):
        return self.min

    def GetMax(self):
        return self.max

    def GetBinEdge(self, i):
        # Axis.cxx:72-78
        imin, imax = self.InverseTransform(self.min), self.InverseTransform(self.max)
        return self.Transform(imin + i * ((imax - imin) / self.n_bins))

    def GetBinEdges(self):
        return np.array([self.GetBinEdge(i) for i in range(self.n_bins + 1)])


class LinearAxis(Axis):
    pass


class PowerAxis(Axis):
    kind = POWER

    def __init__(self, min, max, n_bins, power=1):
        Axis.__init__(self, min, max, n_bins)
        self.power = int(power)

    def Transform(self, v):
        return math.pow(v, self.power)

    def InverseTransform(self, v):
        return math.pow(v, 1.0 / self.power)


class Axes(object):
    geometry = SPHERICAL

    def __init__(self, axes):
        # Axes.cxx:51-64: over- and underflow bin on every axis, last axis contiguous
        self.axes = list(axes)
        n = len(self.axes)
        self.shape = [a.GetNBins() + 2 for a in self.axes]
        self.strides = [1] * n
        for i in range(n - 2, -1, -1):
            self.strides[i] = self.strides[i + 1] * self.shape[i + 1]
        self.n_bins = self.strides[0] * self.shape[0]

    def GetNDim(self):
        return len(self.axes)

    def GetNBins(self):
        return self.n_bins

    def GetShape(self):
        return list(self.shape)

    def GetStrides(self):
        return list(self.strides)

    def at(self, i):
        return self.axes[i]


class SphericalAxes(Axes):
    geometry = SPHERICAL

    def GetBinVolume(self, idxs):
        # Axes.cxx:125-140
        a = self.axes
        scale = 1 if a[1].GetMax() > 180.0 else 2
        return ((a[0].GetBinEdge(idxs[0] + 1) ** 3 - a[0].GetBinEdge(idxs[0]) ** 3) / 3.0) * scale * (math.pi / 180.0) * \
            (a[1].GetBinEdge(idxs[1] + 1) - a[1].GetBinEdge(idxs[1])) * (a[2].GetBinEdge(idxs[2] + 1) - a[2].GetBinEdge(idxs[2]))


class CylindricalAxes(Axes):
    geometry = CYLINDRICAL

    def GetBinVolume(self, idxs):
        # Axes.cxx:161-172
        a = self.axes
        return ((a[0].GetBinEdge(idxs[0] + 1) ** 2 - a[0].GetBinEdge(idxs[0]) ** 2) / 2.0) * 2 * \
            (a[1].GetBinEdge(idxs[1] + 1) - a[1].GetBinEdge(idxs[1])) * (a[2].GetBinEdge(idxs[2] + 1) - a[2].GetBinEdge(idxs[2]))


def default_axes(infinite_muon=False, impact_angle=False):
    """The layouts python/tablemaker/tabulator.py:621-640 uses."""
    if not infinite_muon:
        dims = [PowerAxis(0, 580, 200, 2), LinearAxis(0, 180, 36), LinearAxis(-1, 1, 100), PowerAxis(0, 7e3, 105, 2)]
        geo = SphericalAxes
    else:
        dims = [PowerAxis(0, 580, 100, 2), LinearAxis(0, math.pi, 36), LinearAxis(-8e2, 8e2, 80), PowerAxis(0, 7e3, 105, 2)]
        geo = CylindricalAxes
    if impact_angle:
        dims.append(LinearAxis(-1, 1, 20))
    return geo(dims)


def fill_config(axes, step_length, reference_area, store_squared, angular, keep):
    cfg = TabulatorConfigStruct()
    cfg.struct_size = C.sizeof(TabulatorConfigStruct)
    cfg.geometry, cfg.num_axes, cfg.store_squared_weights = axes.geometry, axes.GetNDim(), int(bool(store_squared))
    for i, a in enumerate(axes.axes):
        cfg.axes[i].kind, cfg.axes[i].power, cfg.axes[i].min, cfg.axes[i].max, cfg.axes[i].n_bins = a.kind, a.power, a.min, a.max, a.n_bins
    cfg.step_length, cfg.reference_area = float(step_length), float(reference_area)
    if angular is not None:
        coef = np.ascontiguousarray(angular.coefficients, dtype=np.float64)
        keep.append(coef)
        cfg.num_angular_coefficients = len(coef)
        cfg.angular_coefficients = coef.ctypes.data_as(C.POINTER(C.c_double))
    return cfg


def _lib():
    L = capi.lib()
    if not getattr(L, "_tab_bound", False):
        L.clsimcu_tabulator_create.argtypes = [C.POINTER(ConfigStruct), C.POINTER(TabulatorConfigStruct), C.POINTER(C.c_void_p)]
        L.clsimcu_tabulator_destroy.argtypes = [C.c_void_p]
        L.clsimcu_tabulator_enqueue.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(ReferenceParticleStruct)]
        L.clsimcu_tabulator_finish.argtypes = [C.c_void_p]
        L.clsimcu_tabulator_info.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.clsimcu_tabulator_get_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L._tab_bound = True
    return L


class I3CLSimStepToTableConverter(object):
    """I3CLSimStepToTableConverter(device, axes, entriesPerStream, storeSquaredWeights, mediumProperties, spectrumTable,
    referenceArea, wavelengthAcceptance, angularAcceptance, rng) (…StepToTableConverter.cxx:123-130).  `entriesPerStream`
    is accepted and ignored (there are no entry buffers), `rng` is a seed, `spectrumTable` must be None."""

    def __init__(self, device, axes, entriesPerStream, storeSquaredWeights, mediumProperties, spectrumTable, referenceArea,
                 wavelengthAcceptance, angularAcceptance, rng, maxNumWorkitems=1 << 15, rng_a=None, rng_x=None, kernelMode=None):
        from . import ice
        if spectrumTable is not None:
            raise capi.ClsimCudaError(-2, "spectrum tables (flasher spectra) are not supported by the table-maker variant yet")
        self.axes = axes
        self.stepLength, self.domArea = 1.0, float(referenceArea)
        gen = ice.makeCherenkovWavelengthGenerator(wavelengthAcceptance, False, mediumProperties)
        # four-axis tables are made by the persistent kernel; the reference-order twin (one work item per step, the reference's
        # stream <-> step mapping) when explicit streams are given, when asked for, and for the impact-angle axis
        if kernelMode is None:
            kernelMode = KERNEL_FAST if (axes.GetNDim() <= 4 and rng_a is None) else KERNEL_REFERENCE
        self.kernelMode = KERNEL_FAST if (kernelMode == KERNEL_FAST and axes.GetNDim() <= 4) else KERNEL_REFERENCE
        opt = ConverterOptions(device=int(device), stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=1.0,
                               fixed_number_of_absorption_lengths=42.0, kernel_mode=self.kernelMode, max_num_workitems=int(maxNumWorkitems),
                               rng_seed=int(rng), rng_n=(int(maxNumWorkitems) if self.kernelMode == KERNEL_REFERENCE else (0 if rng_a is None else len(rng_a))),
                               rng_a=rng_a, rng_x=rng_x)
        scene, self._keep = build_config(mediumProperties, None, [gen], wavelengthAcceptance, opt)
        keep = []
        cfg = fill_config(axes, self.stepLength, referenceArea, storeSquaredWeights, angularAcceptance if axes.GetNDim() <= 4 else None, keep)
        cfg.min_wavelength, cfg.max_wavelength = mediumProperties.GetMinWavelength(), mediumProperties.GetMaxWavelength()
        self._keep.append(keep)
        self._squared = bool(storeSquaredWeights)
        self._h = C.c_void_p()
        capi._check(_lib().clsimcu_tabulator_create(C.byref(scene), C.byref(cfg), C.byref(self._h)))

    def EnqueueSteps(self, steps, reference):
        """reference: (x, y, z, t, dir_x, dir_y, dir_z) of the reference particle."""
        if steps is None:
            return
        steps = np.ascontiguousarray(steps, dtype=STEP_DTYPE)
        ref = ReferenceParticleStruct(*[float(v) for v in reference])
        capi._check(_lib().clsimcu_tabulator_enqueue(self._h, steps.ctypes.data if len(steps) else None, len(steps), C.byref(ref)))

    def Finish(self):
        capi._check(_lib().clsimcu_tabulator_finish(self._h))

    def info(self):
        out = (C.c_double * 16)()
        capi._check(_lib().clsimcu_tabulator_info(self._h, out))
        nd = self.axes.GetNDim()
        return {"n_bins": int(out[0]), "shape": [int(v) for v in out[1:1 + nd]], "strides": [int(v) for v in out[6:6 + nd]],
                "n_photons": out[11], "n_group": out[12], "n_phase": out[13], "spectral_bias_factor": out[14], "photons": int(out[15])}

    def GetTable(self, normalize=False):
        """(bin content, squared weights or None), flat float32 arrays of GetNBins() entries."""
        n = self.axes.GetNBins()
        bins = np.zeros(n, dtype=np.float32)
        sq = np.zeros(n, dtype=np.float32) if self._squared else None
        capi._check(_lib().clsimcu_tabulator_get_table(self._h, int(bool(normalize)), bins.ctypes.data, sq.ctypes.data if sq is not None else None, n))
        return bins, sq

    def close(self):
        if self._h:
            _lib().clsimcu_tabulator_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
