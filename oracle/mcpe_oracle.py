"""CPU restatement of the reference's photon -> MCPE conversion (numpy, double precision).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under clsim_b200/).  Follows, line by line:

* I3CLSimPhotonToMCPEConverterForDOMs::Convert   private/clsim/dom/I3PhotonToMCPEConverter.cxx:602-669
* I3PhotonToMCPEConverter::Convert, per photon    private/clsim/dom/I3PhotonToMCPEConverter.cxx:395-523
* I3CLSimFunctionFromTable::GetValue              private/clsim/function/I3CLSimFunctionFromTable.cxx:106-124
* I3CLSimFunctionPolynomial::GetValue             private/clsim/function/I3CLSimFunctionPolynomial.cxx:86-102
* I3Direction::SetThetaPhi / GetX,Y,Z (dataclasses, un-vendored): direction of TRAVEL from (theta, phi) is
  (sin t cos p, sin t sin p, cos t); clsim fills it with the photon's travel direction
  (private/clsim/I3CLSimClientModule.cxx:341-346).

Pinned by reference data: the DOM acceptance table and the hole-ice polynomial are golden fixtures generated
from the reference's own Python (tests/golden/make_golden.py).  Pinned to the reference's own code: the source above is
compiled unmodified into oracle/_ref/libclsim_ref_mcpe.so (oracle/ref_shim/ref_mcpe.cpp; IceTray's module protocol and
data classes are stand-ins, its random service hands out the uniforms the test supplies), and tests/test_mcpe_oracle.py
holds both functions below against it: the same survivors and times from the same uniforms, the same fatal conditions,
the module's per-DOM time ordering.  Not pinned: hit merging (MCHitMerging, sim-services, un-vendored; off by default).
"""
import numpy as np


def from_table(values, x0, dx, wlen):
    """I3CLSimFunctionFromTable::GetValue, equal spacing mode; vectorised over wlen (float64)."""
    values = np.asarray(values, dtype=np.float64)
    q = (np.asarray(wlen, dtype=np.float64) - x0) / dx
    fbin = np.trunc(q)
    frac = q - fbin                       # modf: fraction carries the sign of q
    ibin = fbin.astype(np.int64)
    low = (ibin < 0) | ((ibin == 0) & (frac < 0))
    high = ~low & (ibin >= len(values) - 1)
    ibin = np.where(low, 0, np.where(high, len(values) - 2, ibin))
    frac = np.where(low, 0.0, np.where(high, 1.0, frac))
    return values[ibin] + (values[ibin + 1] - values[ibin]) * frac


def acceptance_value(f, wlen):
    """f: clsim_b200.description.WlenBias-like (values / start_wlen / wlen_step, or constant)."""
    if f.values is None:
        return np.full(np.shape(wlen), float(f.constant))
    return from_table(f.values, f.start_wlen, f.wlen_step, wlen)


def polynomial(coefficients, x):
    x = np.asarray(x, dtype=np.float64)
    if len(coefficients) == 0:
        return np.zeros_like(x)
    total = np.full(x.shape, float(coefficients[0]))
    mult = np.ones_like(x)
    for c in coefficients[1:]:
        mult = mult * x
        total = total + float(c) * mult
    return total


class Fatal(RuntimeError):
    """A condition the reference answers with log_fatal."""


def convert_inloop(photons, acceptance_of, angular_coefficients, uniforms):
    """I3CLSimPhotonToMCPEConverterForDOMs::Convert over a photon series.

    photons: structured array (clsim_b200.description.PHOTON_DTYPE); acceptance_of: {(string, om): WlenBias};
    uniforms: one draw per photon.  Returns (survivor mask, probability, time)."""
    n = len(photons)
    w = photons["weight"].astype(np.float64)
    if np.any(w < 0):
        raise Fatal("Photon with negative weight found.")
    live = w != 0.0
    r = np.sqrt(photons["x"].astype(np.float64) ** 2 + photons["y"].astype(np.float64) ** 2 + photons["z"].astype(np.float64) ** 2)
    if np.any(live & (np.abs(r - 0.1651) > 0.03)):
        raise Fatal("distance not 165.1mm")
    cos_angle = np.clip(-np.cos(photons["theta"].astype(np.float64)), -1.0, 1.0)
    acc = np.zeros(n)
    keys = photons["string_id"].astype(np.int64) * 65536 + photons["om_id"].astype(np.int64)
    for k in np.unique(keys[live]):
        key = (int(k // 65536), int(k % 65536))
        if key not in acceptance_of:
            raise Fatal("No wavelength acceptance configured for OMKey%s" % (key,))
        sel = live & (keys == k)
        acc[sel] = acceptance_value(acceptance_of[key], photons["wavelength"][sel].astype(np.float64))
    p = w * acc
    p = p * polynomial(angular_coefficients, cos_angle)
    if np.any(live & (p > 1.0)):
        raise Fatal("hitProbability > 1: your hit weights are too high.")
    survive = live & ~(p <= np.asarray(uniforms, dtype=np.float64))
    return survive, p, photons["t"].astype(np.float64)


def convert_module(photons, acceptance, angular_coefficients, efficiency_of, uniforms, oversize=1.0, pancake=1.0, dom_radius=0.1651,
                   dom_dir=(0.0, 0.0, -1.0), only_warn=False):
    """Per-photon part of I3PhotonToMCPEConverter::Convert; photon positions are relative to their DOM, so
    `om.position - photon.GetPos()` is -r."""
    w = photons["weight"].astype(np.float64)
    if np.any(w < 0):
        raise Fatal("Photon with negative weight found.")
    live = w != 0.0
    th, ph = photons["theta"].astype(np.float64), photons["phi"].astype(np.float64)
    dx, dy, dz = np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)
    px, py, pz = -photons["x"].astype(np.float64), -photons["y"].astype(np.float64), -photons["z"].astype(np.float64)
    dist = np.sqrt(px * px + py * py + pz * pz)
    cos_angle = np.clip(-(dx * dom_dir[0] + dy * dom_dir[1] + dz * dom_dir[2]), -1.0, 1.0)
    if pancake == 1.0 and not only_warn and np.any(live & (np.abs(dist - oversize * dom_radius) > 0.03)):
        raise Fatal("distance not %f*%f" % (oversize, dom_radius))
    p = w * acceptance_value(acceptance, photons["wavelength"].astype(np.float64))
    p = p * polynomial(angular_coefficients, cos_angle)
    eff = np.array([efficiency_of[(int(s), int(o))] for s, o in zip(photons["string_id"], photons["om_id"])], dtype=np.float64) if len(photons) else np.zeros(0)
    p = p * eff
    if np.any(live & (p > 1.0)):
        raise Fatal("hitProbability > 1: your hit weights are too high.")
    survive = live & ~(p <= np.asarray(uniforms, dtype=np.float64))
    dot = px * dx + py * dy + pz * dz
    time = photons["t"].astype(np.float64) + dot * (1.0 - pancake / oversize) / photons["group_velocity"].astype(np.float64)
    return survive, p, time


def mwc_uniforms(x, a, n):
    """The converter's draw assignment: with T = len(x) streams, photon j takes draw number j // T of stream
    j % T (mwcrng_kernel.cl:12-28, conversion rounds toward zero).  Returns (uniforms[n], advanced states)."""
    x = np.array(x, dtype=np.uint64)
    a64 = np.asarray(a, dtype=np.uint64)
    T = len(x)
    u = np.zeros(n, dtype=np.float32)
    for first in range(0, n, T):
        m = min(T, n - first)
        lo = x[:m] & np.uint64(0xFFFFFFFF)
        x[:m] = lo * a64[:m] + (x[:m] >> np.uint64(32))
        w = (x[:m] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        # __uint2float_rz: truncate to 24 significant bits
        f = w.astype(np.float64)
        e = np.floor(np.log2(np.maximum(f, 1.0)))
        q = np.exp2(np.maximum(e - 23.0, 0.0))
        f = np.floor(f / q) * q
        u[first:first + m] = (f.astype(np.float32) * np.float32(2.3283064365386963e-10))
    return u, x
