"""Python twin of the reference's converter class, over the C ABI.

``I3CLSimStepToPhotonConverterCUDA`` has the method names, argument meaning and error
behaviour of ``I3CLSimStepToPhotonConverterOpenCL`` as seen through the reference's Python
bindings (private/pybindings/I3CLSimStepToPhotonConverter.cxx:97-247; C++ declarations
public/clsim/I3CLSimStepToPhotonConverterOpenCL.h:78-258).  The compiled-language twin that a
reference maintainer would link is clsim_b200/host/I3CLSimStepToPhotonConverterCUDA.{h,cxx};
both sit on the same C ABI.

``initializeCUDA`` mirrors the factory ``I3CLSimModuleHelper::initializeOpenCL``
(private/clsim/I3CLSimModuleHelper.cxx:303-372) and ``configureCUDADevices`` the option
handling of ``configureOpenCLDevices`` (python/traysegments/common.py:10-77).
"""
import math

import numpy as np

from . import capi
from .description import KERNEL_FAST, KERNEL_REFERENCE, ConverterOptions


class I3CLSimStepToPhotonConverter_exception(RuntimeError):
    pass


class ConversionResult_t(object):
    """identifier, photons (never None), photonHistories (None unless requested)
    (public/clsim/I3CLSimStepToPhotonConverter.h:71-86)."""

    def __init__(self, identifier=0, photons=None, photonHistories=None):
        self.identifier = identifier
        self.photons = photons
        self.photonHistories = photonHistories


def _already():
    return I3CLSimStepToPhotonConverter_exception("I3CLSimStepToPhotonConverterCUDA already initialized!")


class I3CLSimStepToPhotonConverterCUDA(object):
    default_useNativeMath = True

    def __init__(self, randomSeed=0, useNativeMath=True):
        # class defaults of the reference constructor (…ConverterOpenCL.cxx:68-95)
        self._opt = ConverterOptions()
        self._opt.rng_seed = int(randomSeed)
        self._opt.kernel_mode = KERNEL_FAST if useNativeMath else KERNEL_REFERENCE
        self._wlenGenerators = None
        self._wlenBias = None
        self._medium = None
        self._geometry = None
        self._deviceSelected = False
        self._engine = None
        self._compiled = False

    # ---- setters: every one throws after Initialize (…OpenCL.cxx:1324-1523) -----------------
    def _guard(self):
        if self._engine is not None:
            raise _already()
        self._compiled = False

    def SetDevice(self, device):
        self._guard()
        self._opt.device = int(device)
        self._deviceSelected = True

    def SetWlenGenerators(self, wlenGenerators):
        self._guard()
        self._wlenGenerators = list(wlenGenerators)

    def SetWlenBias(self, wlenBias):
        self._guard()
        self._wlenBias = wlenBias

    def SetMediumProperties(self, mediumProperties):
        self._guard()
        self._medium = mediumProperties

    def SetGeometry(self, geometry):
        self._guard()
        self._geometry = geometry

    def SetEnableDoubleBuffering(self, value):
        self._guard()
        self._opt.enable_double_buffering = bool(value)

    def GetEnableDoubleBuffering(self):
        return self._opt.enable_double_buffering

    def SetDoublePrecision(self, value):
        self._guard()
        if value:
            raise I3CLSimStepToPhotonConverter_exception(
                "DoublePrecision is not available in the CUDA converter (fp64 throughput makes it pointless on B200)")

    def GetDoublePrecision(self):
        return False

    def SetStopDetectedPhotons(self, value):
        self._guard()
        self._opt.stop_detected_photons = bool(value)

    def GetStopDetectedPhotons(self):
        return self._opt.stop_detected_photons

    def SetSaveAllPhotons(self, value):
        self._guard()
        self._opt.save_all_photons = bool(value)

    def GetSaveAllPhotons(self):
        return self._opt.save_all_photons

    def SetSaveAllPhotonsPrescale(self, value):
        self._guard()
        self._opt.save_all_photons_prescale = float(value)

    def GetSaveAllPhotonsPrescale(self):
        return self._opt.save_all_photons_prescale

    def SetFixedNumberOfAbsorptionLengths(self, value):
        self._guard()
        self._opt.fixed_number_of_absorption_lengths = float(value)

    def GetFixedNumberOfAbsorptionLengths(self):
        return self._opt.fixed_number_of_absorption_lengths

    def SetDOMPancakeFactor(self, value):
        self._guard()
        self._opt.pancake_factor = float(value)

    def GetDOMPancakeFactor(self):
        return self._opt.pancake_factor

    def SetPhotonHistoryEntries(self, value):
        self._guard()
        self._opt.photon_history_entries = int(value)

    def GetPhotonHistoryEntries(self):
        return self._opt.photon_history_entries

    def SetWorkgroupSize(self, val):
        self._guard()
        self._opt.workgroup_size = int(val)

    def SetMaxNumWorkitems(self, val):
        self._guard()
        if val <= 0:
            raise I3CLSimStepToPhotonConverter_exception("Invalid maximum number of work items!")
        self._opt.max_num_workitems = int(val)

    def SetKernelMode(self, mode):
        """CUDA-only knob: KERNEL_FAST (default) or KERNEL_REFERENCE (exact twin, all options)."""
        self._guard()
        self._opt.kernel_mode = int(mode)

    def SetRNGStreams(self, a, x):
        """Test hook: explicit MWC multipliers/seeds instead of seed-derived ones."""
        self._guard()
        self._opt.rng_a = np.ascontiguousarray(a, dtype=np.uint32)
        self._opt.rng_x = np.ascontiguousarray(x, dtype=np.uint64)
        self._opt.rng_n = len(self._opt.rng_a)

    def SetFirstRNGMultiplierRow(self, row):
        """Multi-GPU: each device takes its own slice of the safe-prime table."""
        self._guard()
        self._opt.rng_first_multiplier = int(row)

    def GetMaxWorkgroupSize(self):
        return 1024

    # ---- life cycle ---------------------------------------------------------------------------
    def Compile(self):
        """Validation only: nothing is JIT-compiled (kernels are built ahead of time for sm_100a)."""
        if self._engine is not None:
            raise _already()
        if self._compiled:
            return
        # same checks, same order as …OpenCL.cxx:492-508
        if not self._wlenGenerators:
            raise I3CLSimStepToPhotonConverter_exception("WlenGenerators not set!")
        if self._wlenBias is None:
            raise I3CLSimStepToPhotonConverter_exception("WlenBias not set!")
        if self._medium is None:
            raise I3CLSimStepToPhotonConverter_exception("MediumProperties not set!")
        if self._geometry is None:
            raise I3CLSimStepToPhotonConverter_exception("Geometry not set!")
        if not self._deviceSelected:
            raise I3CLSimStepToPhotonConverter_exception("Device not selected!")
        if self._opt.save_all_photons and self._opt.stop_detected_photons:
            raise I3CLSimStepToPhotonConverter_exception(
                "Internal error: both the saveAllPhotons and stopDetectedPhotons options are set at the same time.")
        self._compiled = True

    def Initialize(self):
        if self._engine is not None:
            raise _already()
        self.Compile()
        try:
            self._engine = capi.Engine(self._medium, None if self._opt.save_all_photons else self._geometry,
                                       self._wlenGenerators, self._wlenBias, self._opt)
        except capi.ClsimCudaError as e:
            raise I3CLSimStepToPhotonConverter_exception(str(e))

    def IsInitialized(self):
        return self._engine is not None

    def _need(self):
        if self._engine is None:
            raise I3CLSimStepToPhotonConverter_exception("I3CLSimStepToPhotonConverterCUDA is not initialized!")
        return self._engine

    def EnqueueSteps(self, steps, identifier):
        eng = self._need()
        if steps is None:
            raise I3CLSimStepToPhotonConverter_exception("Steps pointer is (null)!")
        try:
            eng.enqueue(steps, identifier)
        except capi.ClsimCudaError as e:
            raise I3CLSimStepToPhotonConverter_exception(str(e))

    def GetWorkgroupSize(self):
        if self._engine is not None:
            return self._engine.workgroup_size()
        return self._opt.workgroup_size or 1

    def GetMaxNumWorkitems(self):
        if self._engine is not None:
            return self._engine.max_num_workitems()
        return self._opt.max_num_workitems

    def QueueSize(self):
        return self._need().queue_size()

    def MorePhotonsAvailable(self):
        return self._need().more_photons_available()

    def GetConversionResult(self):
        try:
            r = self._need().get_result()
        except capi.ClsimCudaError as e:
            raise I3CLSimStepToPhotonConverter_exception(str(e))
        return ConversionResult_t(r.identifier, r.photons, r.history)

    def GetStatistics(self):
        return self._need().statistics()

    def Close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass


def configureCUDADevices(UseGPUs=True, UseCPUs=False, UseOnlyDeviceNumber=None, DoNotParallelize=False,
                         OverrideApproximateNumberOfWorkItems=None, numDevices=None):
    """Device list for the UseGPUs-style tray-segment options (python/traysegments/common.py:10-77).
    Returns a list of dicts(ordinal, approximateNumberOfWorkItems).  UseCPUs is an error: the CUDA
    converter has no CPU path by mandate."""
    if UseCPUs:
        raise RuntimeError("UseCPUs=True is not supported by the CUDA converter (no CPU fallback)")
    if not UseGPUs:
        raise RuntimeError("No matching OpenCL devices. Nothing to do.")
    if numDevices is None:
        try:
            numDevices = capi.device_count()
        except capi.ClsimCudaError:
            numDevices = 0
    if numDevices == 0:
        raise RuntimeError("No matching CUDA devices. Nothing to do.")
    ordinals = list(range(numDevices))
    if UseOnlyDeviceNumber is not None:
        ordinals = [ordinals[UseOnlyDeviceNumber]]
    approx = OverrideApproximateNumberOfWorkItems if OverrideApproximateNumberOfWorkItems is not None else (1 << 20)
    return [{"ordinal": o, "approximateNumberOfWorkItems": int(approx)} for o in ordinals]


def initializeCUDA(device, randomSeed, geometry, medium, wavelengthGenerationBias, wavelengthGenerators,
                   enableDoubleBuffering=False, doublePrecision=False, stopDetectedPhotons=True, saveAllPhotons=False,
                   saveAllPhotonsPrescale=0.01, fixedNumberOfAbsorptionLengths=math.nan, pancakeFactor=1.0,
                   photonHistoryEntries=0, limitWorkgroupSize=0, kernelMode=KERNEL_FAST, rngFirstMultiplierRow=0):
    """Same call sequence as I3CLSimModuleHelper::initializeOpenCL (I3CLSimModuleHelper.cxx:319-369)."""
    conv = I3CLSimStepToPhotonConverterCUDA(randomSeed, useNativeMath=(kernelMode == KERNEL_FAST))
    conv.SetDevice(device["ordinal"])
    conv.SetWlenGenerators(wavelengthGenerators)
    conv.SetWlenBias(wavelengthGenerationBias)
    conv.SetMediumProperties(medium)
    conv.SetGeometry(geometry)
    conv.SetEnableDoubleBuffering(enableDoubleBuffering)
    conv.SetDoublePrecision(doublePrecision)
    conv.SetStopDetectedPhotons(stopDetectedPhotons)
    conv.SetSaveAllPhotons(saveAllPhotons)
    conv.SetSaveAllPhotonsPrescale(saveAllPhotonsPrescale)
    conv.SetFixedNumberOfAbsorptionLengths(fixedNumberOfAbsorptionLengths)
    conv.SetDOMPancakeFactor(pancakeFactor)
    conv.SetPhotonHistoryEntries(photonHistoryEntries)
    conv.SetKernelMode(kernelMode)
    conv.SetFirstRNGMultiplierRow(rngFirstMultiplierRow)
    conv.Compile()
    maxWorkgroupSize = conv.GetMaxWorkgroupSize()
    if limitWorkgroupSize != 0:
        maxWorkgroupSize = min(int(limitWorkgroupSize), maxWorkgroupSize)
    # The CUDA kernels do not tie bunch size to a thread block: granularity 1 is advertised
    # unless the caller limits it (SURVEY.md 8b "sizing handshake").
    conv.SetWorkgroupSize(1 if limitWorkgroupSize == 0 else maxWorkgroupSize)
    workgroupSize = conv.GetWorkgroupSize()
    maxNumWorkitems = (int(device["approximateNumberOfWorkItems"]) // workgroupSize) * workgroupSize
    if maxNumWorkitems == 0:
        maxNumWorkitems = workgroupSize
    conv.SetMaxNumWorkitems(maxNumWorkitems)
    conv.Initialize()
    return conv
