"""CPU restatement of the reference's step generation (MakeSteps of the ppc-style parameterisation), plain Python
doubles, scalar loops -- for small cases.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under clsim_b200/).  Follows:

* I3CLSimLightSourceToStepConverterPPC::MakeSteps_visitor  private/clsim/I3CLSimLightSourceToStepConverterPPC.cxx:523-607
* GenerateStep / GenerateStepForMuon                       …PPC.cxx:785-842
* GenerateStepPreCalculator::FeederThread                  …PPC.cxx:740-760 (angular smearing a=0.39, b=2.61, :105)
* gammaDistributedNumber, scatterDirectionByAngle, mwcRngRandomNumber_co/oc
                                                           private/clsim/I3CLSimLightSourceToStepConverterUtils.h:63-198
* I3CLSimStep::SetDir(x,y,z)                               public/clsim/I3CLSimStep.h:128-133 (through I3Direction,
  dataclasses, un-vendored: theta = pi - zenith, phi = pi + azimuth of the direction the particle comes from)

The reference has no known-answer test for its step generator and draws from racing feeder threads, so the
exact step sequence is not defined by it: parity unpinned for whole sequences.  What is pinned: the samplers
(MWC draws, gammaDistributedNumber, scatterDirectionByAngle) BIT FOR BIT against the reference's own inline
functions, compiled unmodified into oracle/_ref/libclsim_ref_stepgen.so (oracle/ref_shim/ref_stepgen_utils.cpp;
tests/test_stepgen_oracle.py), the gamma and angular samplers against their analytic distributions, and the record
layout.
"""
import ctypes as C
import math

import numpy as np

C_LIGHT = 0.299792458  # I3Constants::c [m/ns]
CASCADE, TRACK_CASCADE_LIKE, TRACK_MUON_LIKE = 0, 1, 2

SOURCE_DTYPE = np.dtype([
    ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("t", "<f8"), ("dir_x", "<f8"), ("dir_y", "<f8"), ("dir_z", "<f8"), ("length", "<f8"),
    ("pa", "<f8"), ("pb", "<f8"), ("num_steps", "<u8"), ("photons_per_step", "<u4"), ("photons_in_last_step", "<u4"),
    ("identifier", "<u4"), ("kind", "<i4")])
assert SOURCE_DTYPE.itemsize == 104


# glibc's single-precision exp and log, the functions the reference's float temporaries select (same libm as the reference
# library this oracle is held against; numpy's float32 exp/log are its own SIMD kernels and differ in the last bit)
_libm = C.CDLL("libm.so.6")
_libm.expf.restype = _libm.logf.restype = C.c_float
_libm.expf.argtypes = _libm.logf.argtypes = [C.c_float]


def _expf(v):
    return float(_libm.expf(v))


def _logf(v):
    return float(_libm.logf(v))


class Mwc(object):
    def __init__(self, x, a):
        self.x, self.a = int(x), int(a)

    def co(self):
        self.x = (self.x & 0xFFFFFFFF) * self.a + (self.x >> 32)
        return float(self.x & 0xFFFFFFFF) / 4294967296.0

    def oc(self):
        return 1.0 - self.co()


def gamma_distributed(shape, rng):
    f32 = np.float32
    if shape < 1.0:
        c = 1.0 / shape
        d = (1.0 - shape) * math.pow(shape, shape / (1.0 - shape))
        while True:
            z = -math.log(rng.oc())
            e = -math.log(rng.oc())
            x = math.pow(z, c)
            if not (z + e < d + x):
                return x
    b = shape - math.log(4.0)
    l = math.sqrt(2.0 * shape - 1.0)
    cheng = 1.0 + math.log(4.5)
    while True:
        rx = rng.oc()
        ry = rng.oc()
        # (ry == 1, one draw in 2^32: the reference divides by zero and goes on with y = +inf, x = +inf and a NaN in the
        # rejection test, which accepts -- a step at infinity; Python would raise instead, so it is spelled out)
        odds = ry / (1.0 - ry) if ry < 1.0 else math.inf
        y = float(f32(math.log(odds) / l))                       # the reference keeps y, z, r in float ...
        x = shape * _expf(y)                                     # ... so its std::exp(y) is the FLOAT overload (expf),
        z = float(f32(rx * ry * ry))
        r = float(f32(b + (shape + l) * y - x))
        if not (r < 4.5 * z - cheng and r < _logf(z)):           # and its std::log(z) is logf
            return x


def scatter_direction(cosa, sina, x, y, z, random_value):
    b = 2.0 * math.pi * random_value
    cosb, sinb = math.cos(b), math.sin(b)
    sinth = math.sqrt(max(0.0, 1.0 - z * z))
    if sinth > 0.0:
        ox, oy, oz = x, y, z
        x = ox * cosa - (oy * cosb + oz * ox * sinb) * sina / sinth
        y = oy * cosa + (ox * cosb - oz * oy * sinb) * sina / sinth
        z = oz * cosa + sina * sinb * sinth
    else:
        x, y = sina * cosb, sina * sinb
        z = cosa if z >= 0.0 else -cosa
    recip = 1.0 / math.sqrt(x * x + y * y + z * z)
    return x * recip, y * recip, z * recip


def theta_phi(x, y, z):
    r = math.sqrt(x * x + y * y + z * z)
    zenith = math.acos(max(-1.0, min(1.0, -z / r)))
    azimuth = math.atan2(-y / r, -x / r)
    if azimuth < 0.0:
        azimuth += 2.0 * math.pi
    phi = math.pi + azimuth
    if phi >= 2.0 * math.pi:
        phi -= 2.0 * math.pi
    return math.pi - zenith, phi


def first_steps(sources):
    counts = sources["num_steps"].astype(np.int64) + (sources["photons_in_last_step"] > 0)
    return np.concatenate([[0], np.cumsum(counts)])


def make_steps(sources, x, a, step_dtype, angular_a=0.39, angular_b=2.61):
    """Steps of the queue entries `sources` (SOURCE_DTYPE) with T = len(x) MWC streams: step j is made from stream
    j % T, the steps of one stream in ascending order.  Returns (steps, advanced states)."""
    first = first_steps(sources)
    total = int(first[-1])
    T = len(x)
    rngs = [Mwc(x[i], a[i]) for i in range(min(T, total))]
    out = np.zeros(total, dtype=step_dtype)
    one_over_a = 1.0 / angular_a
    big_i = 1.0 - math.exp(-angular_b * math.pow(2.0, angular_a))
    entry = np.searchsorted(first, np.arange(total), side="right") - 1
    f32 = np.float32
    for j in range(total):
        src = sources[entry[j]]
        local = j - int(first[entry[j]])
        rng = rngs[j % T]
        s = out[j]
        s["num_photons"] = src["photons_per_step"] if local < int(src["num_steps"]) else src["photons_in_last_step"]
        s["weight"], s["beta"], s["identifier"], s["source_type"] = 1.0, 1.0, src["identifier"], 0
        dx, dy, dz = float(src["dir_x"]), float(src["dir_y"]), float(src["dir_z"])
        if src["kind"] == TRACK_MUON_LIKE:
            s["x"], s["y"], s["z"], s["t"] = f32(src["x"]), f32(src["y"]), f32(src["z"]), f32(src["t"])
            s["length"] = f32(src["length"])
        else:
            along = float(src["pb"]) * gamma_distributed(float(src["pa"]), rng) if src["kind"] == CASCADE else rng.co() * float(src["length"])
            cos_val = max(1.0 - math.pow(-math.log(1.0 - rng.co() * big_i) / angular_b, one_over_a), -1.0)
            sin_val = math.sqrt(1.0 - cos_val * cos_val)
            random_value = rng.co()
            s["x"] = f32(float(src["x"]) + along * dx)
            s["y"] = f32(float(src["y"]) + along * dy)
            s["z"] = f32(float(src["z"]) + along * dz)
            s["t"] = f32(float(src["t"]) + along / C_LIGHT)
            s["length"] = f32(0.001)
            dx, dy, dz = scatter_direction(cos_val, sin_val, dx, dy, dz, random_value)
        th, ph = theta_phi(dx, dy, dz)
        s["theta"], s["phi"] = f32(th), f32(ph)
    x_after = np.array(x, dtype=np.uint64)
    for i, r in enumerate(rngs):
        x_after[i] = r.x
    return out, x_after
