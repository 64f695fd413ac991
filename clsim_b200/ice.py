"""Builders for the inputs of the step->photon path: ice models, DOM acceptance, wavelength
generators.  Host-side (Python where the reference is Python), same names, arguments and
error behaviour as the reference:

* ``MakeIceCubeMediumProperties``  <- python/MakeIceCubeMediumProperties.py:49-245
* ``GetIceTiltZShift``             <- python/util/GetIceTiltZShift.py:40-61
* ``GetSpiceLeaAnisotropyTransforms`` <- python/util/GetSpiceLeaAnisotropyTransforms.py:39-101
* ``GetIceCubeDOMAcceptance``      <- python/GetIceCubeDOMAcceptance.py:36-114
* ``makeCherenkovWavelengthGenerator`` / ``makeWavelengthGenerator``
                                   <- private/clsim/I3CLSimModuleHelper.cxx:52-300 (C++ in the reference)

Ice tables are read either from a ppc-style directory (icemodel.dat, icemodel.par, cfg.txt,
tilt.par, tilt.dat -- same files the reference reads) or, by model name, from the numeric
tables packaged in ``clsim_b200/data/ice_models.json``.
"""
import json
import math
import os

import numpy as np

from .description import MediumProperties, WlenBias, WlenGenerator

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ice_models.json")
_cache = {}

DEG = math.pi / 180.0
NANOMETER = 1e-9


def _packaged():
    if "d" not in _cache:
        with open(_DATA) as f:
            _cache["d"] = json.load(f)
    return _cache["d"]


def _load_tables(iceDataDirectory):
    """-> dict(icemodel_dat[rows], icemodel_par[rows], cfg[], tilt_par|None, tilt_dat|None)."""
    if os.path.isdir(iceDataDirectory):
        d = iceDataDirectory
        has_par = os.path.isfile(d + "/tilt.par")
        has_dat = os.path.isfile(d + "/tilt.dat")
        if has_par and not has_dat:
            raise RuntimeError("ice model directory has tilt.par but tilt.dat is missing!")
        if has_dat and not has_par:
            raise RuntimeError("ice model directory has tilt.dat but tilt.par is missing!")
        out = {
            "icemodel_dat": np.loadtxt(d + "/icemodel.dat"),
            "icemodel_par": np.loadtxt(d + "/icemodel.par"),
            "cfg": np.loadtxt(d + "/cfg.txt"),
            "tilt_par": np.loadtxt(d + "/tilt.par") if has_par else None,
            "tilt_dat": np.loadtxt(d + "/tilt.dat") if has_dat else None,
        }
        return out
    name = os.path.basename(os.path.normpath(iceDataDirectory))
    models = _packaged()
    if name not in models or name.startswith("_"):
        raise RuntimeError("ice model %r is neither a directory nor one of the packaged models %s"
                           % (iceDataDirectory, sorted(k for k in models if not k.startswith("_"))))
    e = models[name]
    return {
        "icemodel_dat": np.array(e["icemodel_dat"], dtype=float),
        "icemodel_par": np.array(e["icemodel_par"], dtype=float),
        "cfg": np.array(e["cfg"], dtype=float),
        "tilt_par": np.array(e["tilt_par"], dtype=float) if "tilt_par" in e else None,
        "tilt_dat": np.array(e["tilt_dat"], dtype=float) if "tilt_dat" in e else None,
    }


def GetSpiceLeaAnisotropyTransforms(anisotropyDirAzimuth=216.0 * DEG, magnitudeAlongDir=0.04, magnitudePerpToDir=-0.08):
    """Returns (absLenScaling parameters, Cpre, Cpost): C = T^T diag(k1,k2,1/(k1 k2))^(+-1) T."""
    k1 = np.exp(magnitudeAlongDir)
    k2 = np.exp(magnitudePerpToDir)
    kz = 1.0 / (k1 * k2)
    stretch = np.diag([k1, k2, kz])
    s, c = np.sin(anisotropyDirAzimuth), np.cos(anisotropyDirAzimuth)
    rot = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])
    c_pre = np.dot(np.dot(rot.T, stretch), rot)
    c_post = np.dot(np.dot(rot.T, np.linalg.inv(stretch)), rot)
    scaling = {
        "anisotropyDirAzimuth": float(anisotropyDirAzimuth),
        "magnitudeAlongDir": float(magnitudeAlongDir),
        "magnitudePerpToDir": float(magnitudePerpToDir),
    }
    return scaling, c_pre, c_post


def GetIceTiltZShift(tiltDirAzimuth=225.0 * DEG, tiltDirectory="spice_lea", detectorCenterDepth=1948.07, _tables=None):
    t = _tables if _tables is not None else _load_tables(tiltDirectory)
    if t["tilt_par"] is None:
        raise RuntimeError("no tilt tables in %r" % (tiltDirectory,))
    distances = t["tilt_par"][:, 1]
    dat = t["tilt_dat"]
    zcoords = (detectorCenterDepth - dat[:, 0])[::-1]
    zshift = np.array([dat[:, i + 1][::-1] for i in range(len(distances))])
    return {
        "distancesFromOriginAlongTilt": np.array(distances, dtype=float),
        "zCoordinates": np.array(zcoords, dtype=float),
        "zCorrections": zshift,
        "directionOfTiltAzimuth": float(tiltDirAzimuth),
    }


def MakeIceCubeMediumProperties(detectorCenterDepth=1948.07, iceDataDirectory="spice_mie", useTiltIfAvailable=True,
                                returnParameters=False):
    t = _load_tables(iceDataDirectory)
    use_tilt = bool(useTiltIfAvailable) and t["tilt_par"] is not None

    par = np.atleast_2d(t["icemodel_par"])
    if len(par) == 6:
        alpha, kappa, A, B, D, E = (par[i][0] for i in range(6))
    elif len(par) == 4:
        alpha, kappa, A, B = (par[i][0] for i in range(4))
        D = 400.0 ** kappa  # what ppc does for the 4-parameter files
        E = 0.0
    else:
        raise RuntimeError("%s/icemodel.par is not a valid Dima-icemodel file. (needs either 4 or 6 entries, this one has %u entries)"
                           % (iceDataDirectory, len(par)))
    cfg = t["cfg"]
    if len(cfg) < 4:
        raise RuntimeError(iceDataDirectory + "/cfg.txt does not have enough configuration lines. It needs at least 4.")
    efficiency = cfg[1]
    f_liu = cfg[2]
    mean_cos = cfg[3]
    has_aniso = False
    if 4 < len(cfg) < 7:
        raise RuntimeError(iceDataDirectory + "/cfg.txt has more than 4 lines (this means you probably get ice anisotropy), but it needs at least 7 lines in this case.")
    elif len(cfg) > 4:
        has_aniso = True
        aniso_az, aniso_along, aniso_perp = cfg[4] * DEG, cfg[5], cfg[6]
    if f_liu < 0.0 or f_liu > 1.0:
        raise RuntimeError("Invalid Liu(SAM) scattering fraction configured in cfg.txt: value=%g" % f_liu)
    if mean_cos < -1.0 or mean_cos > 1.0:
        raise RuntimeError("Invalid <cos(theta)> configured in cfg.txt: value=%g" % mean_cos)

    dat = t["icemodel_dat"]
    depth, b_e400, a_dust400, delta_tau = dat[:, 0], dat[:, 1], dat[:, 2], dat[:, 3]
    if len(depth) < 2:
        raise RuntimeError("There is only a single layer in your layer definition file")
    height = depth[1] - depth[0]
    if height <= 0.0:
        raise RuntimeError("ice layer depths are not in increasing order")
    if np.any(np.abs(np.diff(depth) - height) > 1e-5):
        raise RuntimeError("ice layers are not spaced evenly")

    # file order is top-to-bottom; the medium wants ascending z
    depth = depth[::-1]
    b_400 = b_e400[::-1] / (1.0 - mean_cos)
    depth_top = depth - height / 2.0  # ppc depths are layer centres
    z_start = detectorCenterDepth - (depth_top + height)

    m = MediumProperties()
    m.layersNum = len(z_start)
    m.layersZStart = float(z_start[0])
    m.layersHeight = float(height)
    m.ForcedMinWlen = 265.0 * NANOMETER
    m.ForcedMaxWlen = 675.0 * NANOMETER
    m.efficiency = float(efficiency)
    m.kappa, m.A, m.B, m.D, m.E, m.alpha = float(kappa), float(A), float(B), float(D), float(E), float(alpha)
    m.aDust400 = np.array(a_dust400[::-1], dtype=float)
    m.deltaTau = np.array(delta_tau[::-1], dtype=float)
    m.b400 = np.array(b_400, dtype=float)
    m.scat_kind = 0  # Mixed(SimplifiedLiu, HenyeyGreenstein)
    m.fractionOfFirstDistribution = float(f_liu)
    m.meanCosine = float(mean_cos)
    if has_aniso:
        m.anisotropy, m.preMatrix, m.postMatrix = GetSpiceLeaAnisotropyTransforms(aniso_az, aniso_along, aniso_perp)
        m.preRenormalize = m.postRenormalize = True
    if use_tilt:
        m.tilt = GetIceTiltZShift(tiltDirectory=iceDataDirectory, detectorCenterDepth=detectorCenterDepth, _tables=t)
    if not returnParameters:
        return m
    nan = float("NaN")
    return m, {
        "anisotropyDirAzimuth": aniso_az if has_aniso else nan,
        "anisotropyMagnitudeAlongDir": aniso_along if has_aniso else nan,
        "anisotropyMagnitudePerpToDir": aniso_perp if has_aniso else nan,
    }


def MakeHomogeneousIceMediumProperties(iceDataDirectory="spice_mie", atZ=0.0, detectorCenterDepth=1948.07):
    """BASELINE config 1: one 10 km thick layer (the class defaults layersZStart=-5000 m,
    layersHeight=10000 m of I3CLSimMediumProperties.cxx:43-45) carrying the optical properties
    of the layered model's layer at z = atZ; no tilt, no anisotropy."""
    full = MakeIceCubeMediumProperties(detectorCenterDepth, iceDataDirectory, useTiltIfAvailable=False)
    layer = int((atZ - full.layersZStart) / full.layersHeight)
    layer = min(max(layer, 0), full.layersNum - 1)
    m = MediumProperties()
    m.layersNum, m.layersZStart, m.layersHeight = 1, -5000.0, 10000.0
    m.ForcedMinWlen, m.ForcedMaxWlen = full.ForcedMinWlen, full.ForcedMaxWlen
    m.kappa, m.A, m.B, m.D, m.E, m.alpha = full.kappa, full.A, full.B, full.D, full.E, full.alpha
    m.aDust400 = full.aDust400[layer:layer + 1].copy()
    m.deltaTau = full.deltaTau[layer:layer + 1].copy()
    m.b400 = full.b400[layer:layer + 1].copy()
    m.scat_kind = 0
    m.fractionOfFirstDistribution, m.meanCosine = full.fractionOfFirstDistribution, full.meanCosine
    return m


def GetIceCubeDOMAcceptance(domRadius=0.16510, efficiency=1.0, highQE=False, highQERatio=None):
    """Wavelength acceptance of the IceCube DOM as a 43-entry table starting at 260 nm in
    10 nm steps, normalised to the DOM cross-section (python/GetIceCubeDOMAcceptance.py:36-135).

    highQE multiplies by the wavelength-dependent relative efficiency of the DeepCore PMTs.  The reference
    reads it from ice-models' wv.rde (python/GetIceCubeDOMAcceptance.py:128-130), which is not part of its
    tree: pass the two columns as highQERatio=(wavelengths [nm], ratio)."""
    eff_area = np.array(_packaged()["_dom2007a_eff_area"], dtype=float)
    dom_area = math.pi * domRadius ** 2.0
    values = efficiency * (eff_area / dom_area)
    if highQE:
        if highQERatio is None:
            raise RuntimeError("highQE needs the wv.rde table of ice-models (not vendored by the reference): pass highQERatio=(wv_nm, rde)")
        wv, rde = highQERatio
        values = values * np.interp(260 + 10 * np.arange(len(values)), np.asarray(wv, dtype=float), np.asarray(rde, dtype=float))
    return WlenBias(values=values, start_wlen=260.0 * NANOMETER, wlen_step=10.0 * NANOMETER)


def envelope(functions):
    """Point-wise maximum of acceptance tables on one grid: the generation bias when DOM types differ
    (python/traysegments/common.py:191)."""
    first = functions[0]
    return WlenBias(values=np.max([f.values for f in functions], axis=0), start_wlen=first.start_wlen, wlen_step=first.wlen_step)


def GetFlasherLED405Spectrum():
    """(wavelengths [m], values) of the 405 nm LED data-sheet spectrum (unequal spacing),
    the table GetIceCubeFlasherSpectrumData reads (python/GetIceCubeFlasherSpectrum.py:37-65)."""
    e = _packaged()["_flasher_led_405nm"]
    return np.array(e["wlen_nm"], dtype=float) * NANOMETER, np.array(e["value"], dtype=float)


def _cherenkov_yield(wlen, medium, beta=1.0):
    n_phase = medium.GetPhaseRefractiveIndex(wlen)
    return (2.0 * math.pi / (137.0 * (wlen * wlen))) * (1.0 - 1.0 / (math.pow(beta * n_phase, 2.0)))


def makeCherenkovWavelengthGenerator(wavelengthGenerationBias, generateCherenkovPhotonsWithoutDispersion, mediumProperties):
    """I3CLSimModuleHelper::makeCherenkovWavelengthGenerator (I3CLSimModuleHelper.cxx:176-300)."""
    min_wlen = mediumProperties.GetMinWavelength()
    max_wlen = mediumProperties.GetMaxWavelength()
    wlen_range = max_wlen - min_wlen
    if wlen_range <= 0.0:
        raise RuntimeError("Internal error, wavelength range <= 0!")
    if wavelengthGenerationBias.GetMinWlen() > min_wlen or wavelengthGenerationBias.GetMaxWlen() < max_wlen:
        raise RuntimeError("wavelength generation bias has to have a wavelength range larger or equal to the medium property range!")
    bias = wavelengthGenerationBias
    no_bias = bias.values is None and abs(float(bias.constant) - 1.0) < 1e-10

    def spectrum_at(wlen, b):
        if generateCherenkovPhotonsWithoutDispersion:
            return b * (1.0 / (wlen * wlen))
        return b * _cherenkov_yield(wlen, mediumProperties)

    if bias.values is not None:
        # tabulated bias: tabulate the spectrum on the bias table's own grid
        n = len(bias.values)
        spectrum = np.empty(n)
        for i in range(n):
            wlen = bias.start_wlen + float(i) * bias.wlen_step
            spectrum[i] = spectrum_at(wlen, bias.values[i])
        return WlenGenerator.interpolated(bias.start_wlen, bias.wlen_step, spectrum)
    if no_bias and generateCherenkovPhotonsWithoutDispersion:
        return WlenGenerator.cherenkov_no_dispersion(min_wlen, max_wlen)
    n = int(wlen_range / (10.0 * NANOMETER)) + 2
    step = wlen_range / float(n - 1)
    spectrum = np.empty(n)
    for i in range(n):
        wlen = min_wlen + float(i) * step
        spectrum[i] = spectrum_at(wlen, bias.GetValue(wlen))
    return WlenGenerator.interpolated(min_wlen, step, spectrum)


def makeWavelengthGenerator(spectrumWlens, spectrumValues, wavelengthGenerationBias, mediumProperties):
    """I3CLSimModuleHelper::makeWavelengthGenerator for a tabulated (unequally spaced) spectrum
    (I3CLSimModuleHelper.cxx:75-174): the table binning is re-used and never clipped."""
    wl = np.asarray(spectrumWlens, dtype=float)
    vals = np.asarray(spectrumValues, dtype=float)
    if wavelengthGenerationBias.GetMinWlen() > wl[0] or wavelengthGenerationBias.GetMaxWlen() < wl[-1]:
        raise RuntimeError("wavelength generation bias has to have a wavelength range larger or equal to the spectrum wavelength range!")
    spectrum = np.array([wavelengthGenerationBias.GetValue(w) * v for w, v in zip(wl, vals)])
    return WlenGenerator.interpolated_unequal(wl, spectrum)
