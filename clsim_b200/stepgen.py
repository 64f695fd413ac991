"""Light source -> steps: Python mirror of the reference's ppc-style step converter with the per-step work on the
device (include/clsimcuda.h, "light source -> steps on the device"; kernel in csrc/stepgen.cu).

``I3CLSimLightSourceToStepConverterPPC`` keeps the reference's interface (public/clsim/I3CLSimLightSourceToStepConverterPPC.h;
private/clsim/I3CLSimLightSourceToStepConverterPPC.cxx): setters, Initialize, EnqueueLightSource, EnqueueBarrier,
MoreStepsAvailable, GetConversionResult.  As in the reference, EnqueueLightSource turns a particle into entries of
the step generation queue on the host (light yield, Poisson draw); unlike the reference, the entries are turned
into steps by a CUDA kernel -- either returned to the host (GetConversionResult, reference behaviour) or consumed in
place by the propagation engine (``EnqueueInto``), in which case the steps never exist on the host.

Un-vendored inputs, restated and labelled as such: ``ShowerParameters`` (sim-services' I3SimConstants, used at
…PPC.cxx:288) and the particle type lists.  They shape the workload only; the C ABI takes the resulting numbers.
There is no CPU path for the per-step work.
"""
import ctypes as C
import math

import numpy as np

from . import capi
from .description import STEP_DTYPE

CASCADE, TRACK_CASCADE_LIKE, TRACK_MUON_LIKE = 0, 1, 2
SOURCE_DTYPE = np.dtype([
    ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("t", "<f8"), ("dir_x", "<f8"), ("dir_y", "<f8"), ("dir_z", "<f8"), ("length", "<f8"),
    ("pa", "<f8"), ("pb", "<f8"), ("num_steps", "<u8"), ("photons_per_step", "<u4"), ("photons_in_last_step", "<u4"),
    ("identifier", "<u4"), ("kind", "<i4")])
assert SOURCE_DTYPE.itemsize == 104


class StepGeneratorConfigStruct(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("angular_a", C.c_double), ("angular_b", C.c_double),
                ("rng_seed", C.c_uint64), ("rng_first_multiplier", C.c_uint64)]


def _lib():
    L = capi.lib()
    if not getattr(L, "_stepgen_bound", False):
        L.clsimcu_stepgen_create.argtypes = [C.POINTER(StepGeneratorConfigStruct), C.POINTER(C.c_void_p)]
        L.clsimcu_stepgen_destroy.argtypes = [C.c_void_p]
        L.clsimcu_stepgen_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_stepgen_rng_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clsimcu_enqueue_sources.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]
        L._stepgen_bound = True
    return L


class StepGenerator(object):
    """Owns a clsimcu_step_generator."""

    def __init__(self, device=0, angular_a=0.39, angular_b=2.61, rng_seed=0, rng_first_multiplier=0):
        self._h = C.c_void_p()
        cfg = StepGeneratorConfigStruct()
        cfg.struct_size = C.sizeof(StepGeneratorConfigStruct)
        cfg.device, cfg.angular_a, cfg.angular_b = int(device), float(angular_a), float(angular_b)
        cfg.rng_seed, cfg.rng_first_multiplier = int(rng_seed), int(rng_first_multiplier)
        capi._check(_lib().clsimcu_stepgen_create(C.byref(cfg), C.byref(self._h)))

    def generate(self, sources):
        sources = np.ascontiguousarray(sources, dtype=SOURCE_DTYPE)
        n = C.c_size_t(0)
        total = int((sources["num_steps"].astype(np.int64) + (sources["photons_in_last_step"] > 0)).sum())
        out = np.zeros(total, dtype=STEP_DTYPE)
        capi._check(_lib().clsimcu_stepgen_generate(self._h, sources.ctypes.data if len(sources) else None, len(sources),
                                                    out.ctypes.data if total else None, total, C.byref(n)))
        assert n.value == total
        return out

    def rng_state(self):
        k = C.c_size_t(0)
        capi._check(_lib().clsimcu_stepgen_rng_get(self._h, None, None, 0, C.byref(k)))
        x = np.zeros(k.value, dtype=np.uint64)
        a = np.zeros(k.value, dtype=np.uint32)
        capi._check(_lib().clsimcu_stepgen_rng_get(self._h, x.ctypes.data, a.ctypes.data, k.value, C.byref(k)))
        return x, a

    def enqueue_into(self, engine, sources, identifier):
        """EnqueueSteps on `engine` for a bunch that is made on the device from `sources`."""
        sources = np.ascontiguousarray(sources, dtype=SOURCE_DTYPE)
        capi._check(_lib().clsimcu_enqueue_sources(engine._h, self._h, sources.ctypes.data if len(sources) else None, len(sources), int(identifier)))

    def close(self):
        if self._h:
            _lib().clsimcu_stepgen_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------- host side
class I3CLSimLightSourceToStepConverter_exception(RuntimeError):
    pass


class Particle(object):
    """The fields of I3Particle the converter reads."""
    EM_TYPES = ("EMinus", "EPlus", "Brems", "DeltaE", "PairProd", "Gamma", "Pi0")
    HADRON_TYPES = ("Hadrons", "Neutron", "PiPlus", "PiMinus", "K0_Long", "KPlus", "KMinus", "PPlus", "PMinus", "K0_Short", "NuclInt")
    MUON_TYPES = ("MuMinus", "MuPlus")
    TAU_TYPES = ("TauMinus", "TauPlus")

    def __init__(self, type, energy, pos, direction, time=0.0, length=float("nan"), shape="Null"):
        self.type, self.energy, self.pos, self.time, self.length, self.shape = type, float(energy), tuple(float(v) for v in pos), float(time), float(length), shape
        d = np.asarray(direction, dtype=float)
        self.dir = tuple(d / np.linalg.norm(d))


def ShowerParameters(particle_type, E, density=0.9216):
    """Longitudinal profile and EM-equivalent scale of a cascade: RESTATED FROM MEMORY of sim-services'
    I3SimConstants::ShowerParameters (un-vendored; the reference only calls it, …PPC.cxx:288) -- the standard
    ppc parameterisation.  Unpinned; workload shape only.  E in GeV, density in g/cm^3.
    Returns (a, b [m], emScale, emScaleSigma)."""
    l_rad = 0.358 / density       # radiation length [m]
    logE = max(0.0, math.log(E))
    if particle_type in Particle.EM_TYPES:
        return 2.01849 + 0.63176 * logE, l_rad / 0.63207, 1.0, 0.0
    E0, m, f0, rms0, gamma = 0.18791678, 0.16267529, 0.30974123, 0.95899551, 1.35589541
    e = max(2.71828183, E)
    em_scale = 1.0 - math.pow(e / E0, -m) * (1.0 - f0)
    em_sigma = em_scale * rms0 * math.pow(math.log(e), -gamma)
    return 1.58357292 + 0.41886807 * logE, l_rad / 0.33833116, em_scale, em_sigma


def NumberOfPhotonsPerMeter(medium, wlen_bias, from_wlen, to_wlen):
    """I3CLSimLightSourceToStepConverterUtils::NumberOfPhotonsPerMeter (…Utils.cxx:44-110): Frank-Tamm yield at
    beta = 1 times the generation bias, integrated over photon energy (the reference: GSL QAG to 1e-5)."""
    from scipy import integrate
    h_times_c = 1.0  # only the product energy * wavelength matters: integrate in 1/wavelength

    def f(inv_wlen):
        wlen = h_times_c / inv_wlen
        n = medium.GetPhaseRefractiveIndex(wlen)
        return wlen_bias.GetValue(wlen) * (2.0 * math.pi / 137.0) * (1.0 - 1.0 / (n * n))

    # a tabulated bias is piecewise linear in wavelength: integrate node to node
    edges = [1.0 / to_wlen, 1.0 / from_wlen]
    if getattr(wlen_bias, "values", None) is not None:
        nodes = wlen_bias.start_wlen + wlen_bias.wlen_step * np.arange(len(wlen_bias.values))
        edges += [1.0 / w for w in nodes if from_wlen < w < to_wlen]
    edges = sorted(edges)
    return sum(integrate.quad(f, lo, hi, epsrel=1e-5)[0] for lo, hi in zip(edges[:-1], edges[1:]))


class I3CLSimLightSourceToStepConverterPPC(object):
    def __init__(self, photonsPerStep=200, highPhotonsPerStep=0, useHighPhotonsPerStepStartingFromNumPhotons=1.0e9, device=0):
        # …PPC.cxx:52-73
        if photonsPerStep <= 0:
            raise I3CLSimLightSourceToStepConverter_exception("photonsPerStep may not be <= 0!")
        self.photonsPerStep = int(photonsPerStep)
        self.highPhotonsPerStep = int(highPhotonsPerStep) if highPhotonsPerStep else self.photonsPerStep
        self.useHighFrom = float(useHighPhotonsPerStepStartingFromNumPhotons)
        self.useCascadeExtension = True
        self.maxBunchSize = 512000
        self.bunchSizeGranularity = 1
        self.medium = self.wlenBias = None
        self.seed = None
        self.initialized = False
        self.barrier = False
        self.queue = []
        self.device = device
        self.generator = None

    def _guard(self):
        if self.initialized:
            raise I3CLSimLightSourceToStepConverter_exception("I3CLSimLightSourceToStepConverterPPC already initialized!")

    def SetUseCascadeExtension(self, flag):
        self.useCascadeExtension = bool(flag)

    def SetBunchSizeGranularity(self, num):
        self._guard()
        if num <= 0:
            raise I3CLSimLightSourceToStepConverter_exception("BunchSizeGranularity of 0 is invalid!")
        if num != 1:
            raise I3CLSimLightSourceToStepConverter_exception("A BunchSizeGranularity != 1 is currently not supported!")
        self.bunchSizeGranularity = num

    def SetMaxBunchSize(self, num):
        self._guard()
        if num <= 0:
            raise I3CLSimLightSourceToStepConverter_exception("MaxBunchSize of 0 is invalid!")
        self.maxBunchSize = int(num)

    def SetWlenBias(self, bias):
        self._guard()
        self.wlenBias = bias

    def SetMediumProperties(self, medium):
        self._guard()
        self.medium = medium

    def SetRandomService(self, seed):
        self._guard()
        self.seed = int(seed)

    def Initialize(self, rngFirstMultiplierRow=0):
        # …PPC.cxx:80-135
        self._guard()
        if self.seed is None:
            raise I3CLSimLightSourceToStepConverter_exception("RandomService not set!")
        if self.wlenBias is None:
            raise I3CLSimLightSourceToStepConverter_exception("WlenBias not set!")
        if self.medium is None:
            raise I3CLSimLightSourceToStepConverter_exception("MediumProperties not set!")
        if self.bunchSizeGranularity > self.maxBunchSize:
            raise I3CLSimLightSourceToStepConverter_exception("BunchSizeGranularity must not be greater than MaxBunchSize!")
        if self.maxBunchSize % self.bunchSizeGranularity != 0:
            raise I3CLSimLightSourceToStepConverter_exception("MaxBunchSize is not a multiple of BunchSizeGranularity!")
        self.host_rng = np.random.default_rng(self.seed)
        self.generator = StepGenerator(device=self.device, rng_seed=self.seed, rng_first_multiplier=rngFirstMultiplierRow)
        # the refractive index is layer independent in the IceCube models: one yield for all layers
        self.meanPhotonsPerMeter = NumberOfPhotonsPerMeter(self.medium, self.wlenBias, self.medium.GetMinWavelength(), self.medium.GetMaxWavelength())
        self.initialized = True

    def IsInitialized(self):
        return self.initialized

    def _draw_count(self, mean):
        # Poisson, Gaussian above 1e7 (…PPC.cxx:301-321)
        if mean > 1e7:
            while True:
                v = self.host_rng.normal(mean, math.sqrt(mean))
                if v >= 0.0:
                    return int(v)
        return int(self.host_rng.poisson(mean))

    def _entry(self, p, identifier, kind, num_photons, length=0.0, pa=0.0, pb=0.0):
        per_step = self.highPhotonsPerStep if float(num_photons) > self.useHighFrom else self.photonsPerStep
        e = np.zeros((), dtype=SOURCE_DTYPE)
        e["x"], e["y"], e["z"], e["t"] = p.pos[0], p.pos[1], p.pos[2], p.time
        e["dir_x"], e["dir_y"], e["dir_z"] = p.dir
        e["length"], e["pa"], e["pb"] = length, pa, pb
        e["num_steps"], e["photons_per_step"] = num_photons // per_step, per_step
        e["photons_in_last_step"] = num_photons % per_step
        e["identifier"], e["kind"] = identifier, kind
        return e

    def EnqueueLightSource(self, particle, identifier):
        # …PPC.cxx:190-481
        if not self.initialized:
            raise I3CLSimLightSourceToStepConverter_exception("I3CLSimLightSourceToStepConverterPPC is not initialized!")
        if self.barrier:
            raise I3CLSimLightSourceToStepConverter_exception("A barrier is enqueued! You must receive all steps before enqueuing a new particle.")
        p = particle
        density = 0.9216
        E = p.energy
        logE = max(0.0, math.log(E))
        if p.type in Particle.EM_TYPES or p.type in Particle.HADRON_TYPES:
            nph = 5.21 * 0.924 / density
            a, b, em_scale, em_sigma = ShowerParameters(p.type, E, density)
            f = 1.0
            if em_sigma != 0.0:
                while True:
                    f = em_scale + em_sigma * self.host_rng.normal(0.0, 1.0)
                    if 0.0 <= f <= 1.0:
                        break
            n = self._draw_count(f * self.meanPhotonsPerMeter * nph * E)
            if p.shape == "CascadeSegment":
                if not (p.length > 0):
                    raise I3CLSimLightSourceToStepConverter_exception("Found a cascade segment with length %g. This should not be." % p.length)
                self.queue.append(self._entry(p, identifier, TRACK_CASCADE_LIKE, n, length=p.length))
            else:
                self.queue.append(self._entry(p, identifier, CASCADE, n, pa=a, pb=(b if self.useCascadeExtension else 0.0)))
        elif p.type in Particle.MUON_TYPES or p.type in Particle.TAU_TYPES:
            length = 2000.0 if math.isnan(p.length) else p.length
            extr = 1.0 + max(0.0, 0.1880 + 0.0206 * logE)
            total = self.meanPhotonsPerMeter * length * extr
            n_mu = self._draw_count(total / extr)
            n_ca = self._draw_count(total * (1.0 - 1.0 / extr))
            self.queue.append(self._entry(p, identifier, TRACK_MUON_LIKE, n_mu, length=length))
            e = self._entry(p, identifier, TRACK_CASCADE_LIKE, n_ca, length=length)
            # the reference takes the remainder of the STEP count here, not of the photon count (…PPC.cxx:457):
            # reproduced
            e["photons_in_last_step"] = int(e["num_steps"]) % int(e["photons_per_step"])
            self.queue.append(e)
        else:
            raise I3CLSimLightSourceToStepConverter_exception("I3CLSimLightSourceToStepConverterPPC cannot handle a %s." % p.type)

    def EnqueueBarrier(self):
        if not self.initialized:
            raise I3CLSimLightSourceToStepConverter_exception("I3CLSimLightSourceToStepConverterPPC is not initialized!")
        if self.barrier:
            raise I3CLSimLightSourceToStepConverter_exception("A barrier is already enqueued!")
        self.queue.append(None)
        self.barrier = True

    def BarrierActive(self):
        return self.barrier

    def MoreStepsAvailable(self):
        return len(self.queue) > 0

    def _next_sources(self):
        """One call of MakeSteps (…PPC.cxx:566-640): the front entry, at most maxBunchSize steps of it."""
        front = self.queue[0]
        if front is None:
            self.queue.pop(0)
            return None
        take = min(int(front["num_steps"]), self.maxBunchSize)
        part = front.copy()
        part["num_steps"] = take
        front["num_steps"] = int(front["num_steps"]) - take
        if int(front["num_steps"]) == 0 and take < self.maxBunchSize:
            self.queue.pop(0)   # the partial last step goes with this part
        else:
            part["photons_in_last_step"] = 0
        return np.array([part], dtype=SOURCE_DTYPE)

    def GetConversionResultWithBarrierInfo(self):
        """(steps, barrierWasReset), steps made on the device and copied back (reference behaviour)."""
        if not self.initialized:
            raise I3CLSimLightSourceToStepConverter_exception("I3CLSimLightSourceToStepConverterPPC is not initialized!")
        if not self.queue:
            raise I3CLSimLightSourceToStepConverter_exception("I3CLSimLightSourceToStepConverterPPC: no particle is enqueued!")
        src = self._next_sources()
        if src is None:
            self.barrier = False
            return np.zeros(0, dtype=STEP_DTYPE), True
        return self.generator.generate(src), False

    def GetConversionResult(self):
        return self.GetConversionResultWithBarrierInfo()[0]

    def EnqueueInto(self, engine, identifier, max_steps=None):
        """Device-resident variant: packs queue entries up to `max_steps` steps (default: the engine's maximum bunch)
        into one bunch that the engine generates and propagates on the device.  Returns the number of steps, 0 when a
        barrier was reached or the queue is empty."""
        cap = int(max_steps or engine.max_num_workitems())
        picked, total = [], 0
        while self.queue and self.queue[0] is not None and total < cap and len(picked) < 65536:
            front = self.queue[0]
            whole = int(front["num_steps"]) + (1 if int(front["photons_in_last_step"]) > 0 else 0)
            if total + whole <= cap:
                picked.append(front.copy())
                total += whole
                self.queue.pop(0)
            else:
                take = min(int(front["num_steps"]), cap - total)
                if take == 0:
                    break
                part = front.copy()
                part["num_steps"], part["photons_in_last_step"] = take, 0
                front["num_steps"] = int(front["num_steps"]) - take
                picked.append(part)
                total += take
        if not picked:
            if self.queue and self.queue[0] is None:
                self.queue.pop(0)
                self.barrier = False
            return 0
        self.generator.enqueue_into(engine, np.array(picked, dtype=SOURCE_DTYPE), identifier)
        return total
