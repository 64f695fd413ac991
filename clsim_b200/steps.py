"""Synthetic step bundles (48-byte I3CLSimStep records) for the BASELINE configurations.

The real step generator is upstream of the drop-in boundary and out of scope (SURVEY.md 2,
row 20); these functions only imitate the statistics of its output so that the path can be
exercised and timed:

* muon-like steps are what ``GenerateStepForMuon`` emits (I3CLSimLightSourceToStepConverterPPC.cxx:821-842):
  position = track start, length = whole track, so photons are spread uniformly along it;
* cascade-like steps are what ``GenerateStep`` emits (:785-819): 1 mm long, placed along the
  axis, direction smeared with ``cos = max(1-(-ln(1-U*I)/b)^(1/a), -1)``, a=0.39, b=2.61,
  I = 1-exp(-b*2^a) (:680-760);
* the muon : cascade photon ratio follows ``1 : max(0, 0.1880+0.0206 ln E)`` (:371-378).

Everything is seeded through ``numpy.random.default_rng``.
"""
import math

import numpy as np

from .description import STEP_DTYPE

C_LIGHT = 0.299792458  # m/ns


def _dir_to_theta_phi(d):
    """I3Direction::CalcTheta/CalcPhi of the direction of travel (I3CLSimStep.h:123-133)."""
    d = np.asarray(d, dtype=float)
    n = np.sqrt((d ** 2).sum(-1))
    theta = np.arccos(np.clip(d[..., 2] / n, -1.0, 1.0))
    phi = np.arctan2(d[..., 1], d[..., 0])
    phi = np.where(phi < 0.0, phi + 2.0 * math.pi, phi)
    return theta, phi


def _rotate_by_angle(axis, cosa, rnd):
    """Host double twin of scatterDirectionByAngle (I3CLSimLightSourceToStepConverterUtils.h:160-198)."""
    axis = np.asarray(axis, dtype=float)
    sina = np.sqrt(np.maximum(0.0, 1.0 - cosa * cosa))
    b = 2.0 * math.pi * rnd
    cosb, sinb = np.cos(b), np.sin(b)
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    sinth = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    safe = np.where(sinth > 0.0, sinth, 1.0)
    nx = np.where(sinth > 0.0, x * cosa - (y * cosb + z * x * sinb) * sina / safe, sina * cosb)
    ny = np.where(sinth > 0.0, y * cosa + (x * cosb - z * y * sinb) * sina / safe, sina * sinb)
    nz = np.where(sinth > 0.0, z * cosa + sina * sinb * sinth, cosa * np.sign(z))
    out = np.stack([nx, ny, nz], axis=-1)
    return out / np.sqrt((out ** 2).sum(-1))[..., None]


def _fill(n, pos, t, direction, length, num_photons, identifier, source_type=0, weight=1.0, beta=1.0):
    steps = np.zeros(n, dtype=STEP_DTYPE)
    pos = np.broadcast_to(np.asarray(pos, dtype=float), (n, 3))
    theta, phi = _dir_to_theta_phi(np.broadcast_to(np.asarray(direction, dtype=float), (n, 3)))
    steps["x"], steps["y"], steps["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
    steps["t"] = t
    steps["theta"], steps["phi"] = theta, phi
    steps["length"] = length
    steps["beta"] = beta
    steps["num_photons"] = num_photons
    steps["weight"] = weight
    steps["identifier"] = identifier
    steps["source_type"] = source_type
    return steps


def point_source_steps(num_steps=5000, photons_per_step=200, pos=(0.0, 0.0, 0.0), seed=1):
    """BASELINE config 1: isotropic 1 mm steps at one point."""
    rng = np.random.default_rng(seed)
    cz = rng.uniform(-1.0, 1.0, num_steps)
    ph = rng.uniform(0.0, 2.0 * math.pi, num_steps)
    s = np.sqrt(1.0 - cz * cz)
    d = np.stack([s * np.cos(ph), s * np.sin(ph), cz], axis=-1)
    return _fill(num_steps, pos, 0.0, d, 1e-3, photons_per_step, np.arange(num_steps) % 64)


def cascade_smeared_directions(axis, n, rng, a=0.39, b=2.61):
    big_i = 1.0 - math.exp(-b * 2.0 ** a)
    u = rng.uniform(0.0, 1.0, n)
    cs = np.maximum(1.0 - np.power(-np.log(1.0 - u * big_i) / b, 1.0 / a), -1.0)
    return _rotate_by_angle(np.broadcast_to(np.asarray(axis, dtype=float), (n, 3)), cs, rng.uniform(0.0, 1.0, n))


def muon_track_steps(num_steps, photons_per_step=200, energy_gev=1e4, track_length=1000.0, zenith_deg=45.0, azimuth_deg=30.0,
                     center=(0.0, 0.0, 0.0), lateral_offset=(0.0, 0.0), t0=0.0, seed=2, identifier=0):
    """BASELINE config 2: one muon through the detector centre, ppc parameterisation shape."""
    rng = np.random.default_rng(seed)
    # I3Particle zenith/azimuth give where the particle comes FROM; travel direction is the opposite
    zen, azi = math.radians(zenith_deg), math.radians(azimuth_deg)
    travel = -np.array([math.sin(zen) * math.cos(azi), math.sin(zen) * math.sin(azi), math.cos(zen)])
    perp1 = np.cross(travel, [0.0, 0.0, 1.0])
    perp1 /= np.linalg.norm(perp1)
    perp2 = np.cross(travel, perp1)
    start = np.asarray(center, dtype=float) - 0.5 * track_length * travel + lateral_offset[0] * perp1 + lateral_offset[1] * perp2
    extr = 1.0 + max(0.0, 0.1880 + 0.0206 * math.log(energy_gev))
    n_muon = int(round(num_steps / extr))
    n_casc = num_steps - n_muon
    mu = _fill(n_muon, start, t0, travel, track_length, photons_per_step, identifier)
    along = rng.uniform(0.0, track_length, n_casc)
    cpos = start[None, :] + along[:, None] * travel[None, :]
    cdir = cascade_smeared_directions(travel, n_casc, rng)
    ca = _fill(n_casc, cpos, t0 + along / C_LIGHT, cdir, 1e-3, photons_per_step, identifier)
    steps = np.concatenate([mu, ca])
    rng.shuffle(steps)
    return steps


def muon_track_sources(num_steps, photons_per_step=200, energy_gev=1e4, track_length=1000.0, zenith_deg=45.0, azimuth_deg=30.0,
                       center=(0.0, 0.0, 0.0), t0=0.0, identifier=0):
    """The same workload as ``muon_track_steps`` as two entries of the step generation queue (muon-like and
    cascade-like steps of one track), for bunches that are made on the device (stepgen.py)."""
    from .stepgen import SOURCE_DTYPE, TRACK_CASCADE_LIKE, TRACK_MUON_LIKE
    zen, azi = math.radians(zenith_deg), math.radians(azimuth_deg)
    travel = -np.array([math.sin(zen) * math.cos(azi), math.sin(zen) * math.sin(azi), math.cos(zen)])
    start = np.asarray(center, dtype=float) - 0.5 * track_length * travel
    extr = 1.0 + max(0.0, 0.1880 + 0.0206 * math.log(energy_gev))
    n_muon = int(round(num_steps / extr))
    src = np.zeros(2, dtype=SOURCE_DTYPE)
    src["x"], src["y"], src["z"], src["t"] = start[0], start[1], start[2], t0
    src["dir_x"], src["dir_y"], src["dir_z"] = travel
    src["length"] = track_length
    src["kind"] = [TRACK_MUON_LIKE, TRACK_CASCADE_LIKE]
    src["num_steps"] = [n_muon, num_steps - n_muon]
    src["photons_per_step"] = photons_per_step
    src["identifier"] = identifier
    return src


def muon_bundle_steps(num_steps, num_muons=100, spread=20.0, seed=3, **kw):
    """BASELINE config 3: parallel muons with a lateral spread."""
    rng = np.random.default_rng(seed)
    per = max(1, num_steps // num_muons)
    parts = []
    for k in range(num_muons):
        off = rng.normal(0.0, spread, 2)
        n = per if k < num_muons - 1 else num_steps - per * (num_muons - 1)
        if n <= 0:
            continue
        parts.append(muon_track_steps(n, lateral_offset=off, seed=seed * 1000 + k, identifier=k, **kw))
    return np.concatenate(parts)[:num_steps]


def cascade_steps(num_steps, photons_per_step=200, energy_gev=1e6, pos=(0.0, 0.0, 0.0), zenith_deg=60.0, azimuth_deg=120.0,
                  seed=4, identifier=0):
    """BASELINE config 4: EM cascade; longitudinal profile b*Gamma(a) along the axis with the
    standard ppc constants (a = 2.03+0.604 ln E, b = 0.633/L_rad, L_rad = 0.358/0.9216 m);
    workload shape only -- the constants live in un-vendored sim-services."""
    rng = np.random.default_rng(seed)
    zen, azi = math.radians(zenith_deg), math.radians(azimuth_deg)
    travel = -np.array([math.sin(zen) * math.cos(azi), math.sin(zen) * math.sin(azi), math.cos(zen)])
    a = 2.03 + 0.604 * math.log(energy_gev)
    l_rad = 0.358 / 0.9216
    along = rng.gamma(a, l_rad / 0.633, num_steps)
    cpos = np.asarray(pos, dtype=float)[None, :] + along[:, None] * travel[None, :]
    cdir = cascade_smeared_directions(travel, num_steps, rng)
    return _fill(num_steps, cpos, along / C_LIGHT, cdir, 1e-3, photons_per_step, identifier)


def flasher_steps(num_steps, dom_pos, photons_per_step=200, led_azimuth_deg=0.0, tilted=False, seed=5, identifier=0):
    """BASELINE config 5: LED light starting inside a DOM (sourceType 1), Gaussian angular
    smearing sigma = (9.2 deg polar, 10.1 deg azimuthal)
    (python/FlasherInfoVectToFlasherPulseSeriesConverter.py:86-92), zero-length steps."""
    rng = np.random.default_rng(seed)
    elev = math.radians(48.0 if tilted else 0.0)
    azi = math.radians(led_azimuth_deg) + rng.normal(0.0, math.radians(10.1), num_steps)
    pol = elev + rng.normal(0.0, math.radians(9.2), num_steps)
    d = np.stack([np.cos(pol) * np.cos(azi), np.cos(pol) * np.sin(azi), np.sin(pol)], axis=-1)
    t = rng.exponential(5.0, num_steps)
    return _fill(num_steps, dom_pos, t, d, 0.0, photons_per_step, identifier, source_type=1)


def pad_to_granularity(steps, granularity):
    """Dummy-step padding of the last bunch (I3CLSimLightSourceToStepConverterAsync.cxx:210-263):
    numPhotons=0, weight=0."""
    rem = len(steps) % granularity
    if rem == 0:
        return steps
    pad = np.zeros(granularity - rem, dtype=STEP_DTYPE)
    pad["beta"] = 1.0
    return np.concatenate([steps, pad])
