"""Shared scene builders for the tests (the five BASELINE configurations, small)."""
import math
import os

import numpy as np

from clsim_b200 import geometry, ice, steps
from clsim_b200.description import KERNEL_FAST, KERNEL_REFERENCE, ConverterOptions  # noqa: F401


# tests/test_hostcheck.py re-runs GPU tests of everything but the fast kernel in a subprocess against the CUDA sources compiled
# for the host (tests/hostcheck): there the "device" is one CPU thread and holds no fast kernel, so the tests that only need SOME
# kernel to make photons take the reference-order one, on smaller bunches.  On a GPU (the variable unset) nothing changes.
HOSTCHECK = os.environ.get("CLSIM_HOSTCHECK") == "1"
DEVICE_KERNEL = KERNEL_REFERENCE if HOSTCHECK else KERNEL_FAST


def sized(on_gpu, on_host):
    return on_host if HOSTCHECK else on_gpu


class Scene(object):
    def __init__(self, medium, geo, generators, bias, pancake, oversize):
        self.medium, self.geo, self.generators, self.bias = medium, geo, generators, bias
        self.pancake, self.oversize = pancake, oversize

    def options(self, **kw):
        base = dict(stop_detected_photons=True, pancake_factor=self.pancake)
        base.update(kw)
        return ConverterOptions(**base)


def make_scene(name, oversize=5.0, geo_kind="ic86"):
    if name == "homogeneous":       # config 1
        medium = ice.MakeHomogeneousIceMediumProperties("spice_mie")
    elif name == "spice_mie":       # config 2
        medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_mie", useTiltIfAvailable=False)
    elif name == "spice_mie_tilt":
        medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_mie", useTiltIfAvailable=True)
    elif name == "spice_lea":       # configs 3-5
        medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=True)
    elif name == "spice_lea_notilt":
        medium = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=False)
    else:
        raise KeyError(name)
    geo = geometry.make_ic86_like_geometry(oversize) if geo_kind == "ic86" else geometry.make_ring_geometry(oversize)
    bias = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * oversize)
    gens = [ice.makeCherenkovWavelengthGenerator(bias, False, medium)]
    return Scene(medium, geo, gens, bias, oversize, oversize)


def add_flasher_generator(scene):
    wl, val = ice.GetFlasherLED405Spectrum()
    scene.generators = scene.generators + [ice.makeWavelengthGenerator(wl, val, scene.bias, scene.medium)]
    return scene


def rng_streams(n, seed=1234):
    from oracle import pyoracle
    from clsim_b200 import capi
    a = capi.safeprime_multipliers(0, n)
    x = pyoracle.seed_states(seed, a)
    return a, x


def sort_photons(p):
    """Canonical order for comparing hit lists whose emission order is unspecified."""
    key = np.lexsort((p["start_phi"], p["start_theta"], p["wavelength"], p["om_id"], p["string_id"], p["identifier"]))
    return p[key]


def dom_near(geo, point):
    d2 = (geo.posX - point[0]) ** 2 + (geo.posY - point[1]) ** 2 + (geo.posZ - point[2]) ** 2
    i = int(np.argmin(d2))
    return np.array([geo.posX[i], geo.posY[i], geo.posZ[i]])


def match_photons(got, want, wl_digits=4):
    """Pair records of two hit lists that describe the same photon: same bunch identifier,
    DOM, scatter count and (to ~1e-4 relative) wavelength and start direction.
    Returns (index pairs, fraction of `want` matched)."""
    def key(p):
        return (int(p["identifier"]), int(p["string_id"]), int(p["om_id"]), int(p["num_scatters"]),
                round(float(p["wavelength"]) * 1e9, wl_digits - 2), round(float(p["start_theta"]), 3), round(float(p["start_phi"]), 3))
    table = {}
    for i, p in enumerate(want):
        table.setdefault(key(p), []).append(i)
    pairs = []
    for j, p in enumerate(got):
        lst = table.get(key(p))
        if lst:
            pairs.append((j, lst.pop()))
    return np.array(pairs, dtype=np.int64).reshape(-1, 2), (len(pairs) / float(max(1, len(want))))
