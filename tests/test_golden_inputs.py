"""Host-side input builders against golden vectors produced by the reference's own Python
(tests/golden/make_golden.py): ice tables, tilt, anisotropy transforms, DOM acceptance,
safe primes.  CPU only."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

from clsim_b200 import capi, ice

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.mark.parametrize("model,tilt,fixture", [
    ("spice_mie", True, "medium_spice_mie.json"),
    ("spice_mie", False, "medium_spice_mie_notilt.json"),
    ("spice_lea", True, "medium_spice_lea.json"),
    ("spice_lea", False, "medium_spice_lea_notilt.json"),
    ("spice_1", True, "medium_spice_1.json"),
    ("ppc_aha_0.80", True, "medium_ppc_aha_0.80.json"),
])
def test_medium_matches_reference_loader(model, tilt, fixture):
    g = gold(fixture)
    m = ice.MakeIceCubeMediumProperties(iceDataDirectory=model, useTiltIfAvailable=tilt)
    # bit-exact: same arithmetic on the same table values
    assert m.layersNum == g["layersNum"]
    assert m.layersZStart == g["layersZStart"]
    assert m.layersHeight == g["layersHeight"]
    assert m.ForcedMinWlen == g["ForcedMinWlen"] and m.ForcedMaxWlen == g["ForcedMaxWlen"]
    for k in ("kappa", "A", "B", "D", "E", "alpha"):
        assert getattr(m, k) == g[k], k
    assert np.array_equal(m.aDust400, g["aDust400"])
    assert np.array_equal(m.deltaTau, g["deltaTau"])
    assert np.array_equal(m.b400, g["b400"])
    assert m.fractionOfFirstDistribution == g["fractionOfFirstDistribution"]
    assert m.meanCosine == g["meanCosine"]
    if g["anisotropy"] is None:
        assert m.anisotropy is None
    else:
        assert m.anisotropy == g["anisotropy"]
        assert np.array_equal(np.asarray(m.preMatrix).ravel(), g["preMatrix"])
        assert np.array_equal(np.asarray(m.postMatrix).ravel(), g["postMatrix"])
        assert m.preRenormalize == g["preRenormalize"] and m.postRenormalize == g["postRenormalize"]
    if g["tilt"] is None:
        assert m.tilt is None
    else:
        assert np.array_equal(m.tilt["distancesFromOriginAlongTilt"], g["tilt"]["distancesFromOriginAlongTilt"])
        assert np.array_equal(m.tilt["zCoordinates"], g["tilt"]["zCoordinates"])
        assert np.array_equal(m.tilt["zCorrections"], np.array(g["tilt"]["zCorrections"]))
        assert m.tilt["directionOfTiltAzimuth"] == g["tilt"]["directionOfTiltAzimuth"]


def test_dom_acceptance_matches_reference():
    g = gold("dom_acceptance.json")
    b = ice.GetIceCubeDOMAcceptance()
    assert b.start_wlen == g["startWlen"] and b.wlen_step == g["wlenStep"]
    np.testing.assert_allclose(b.values, g["values"], rtol=1e-15)
    assert len(b.values) == 43
    # oversized radius only rescales the table
    b5 = ice.GetIceCubeDOMAcceptance(domRadius=0.16510 * 5.0)
    np.testing.assert_allclose(b5.values * 25.0, b.values, rtol=1e-14)


def test_anisotropy_transforms_match_ppc_formulas():
    """resources/tests/testSpiceLeaTransforms.py: |clsim - ppc| <= 1e-14 for the pre/post matrices."""
    g = gold("ppc_formulas.json")
    p = g["params"]
    _, cpre, cpost = ice.GetSpiceLeaAnisotropyTransforms(p["thx_deg"] * math.pi / 180.0, p["logk1"], p["logk2"])
    np.testing.assert_array_equal(cpre.ravel(), g["Cpre"])
    np.testing.assert_array_equal(cpost.ravel(), g["Cpost"])
    v = np.array(g["unit_vectors"])
    for mat, want in ((cpre, g["PPCPre"]), (cpost, g["PPCPost"])):
        out = v @ mat.T
        out /= np.sqrt((out ** 2).sum(1))[:, None]
        assert np.abs(out - np.array(want)).max() <= 1e-14


def test_safeprimes_match_rnd_txt():
    """The first rows of the reference's CUDAMCML-compatible table (rnd.txt, 16 028 rows)."""
    g = gold("safeprimes.json")
    a = capi.safeprime_multipliers(0, g["rows"])
    assert list(a[:32]) == g["first_32"]
    assert list(a[-8:]) == g["last_8"]
    assert int(a[1000]) == g["row_1000"] and int(a[10000]) == g["row_10000"]
    assert hashlib.sha256(a.astype("<u4").tobytes()).hexdigest() == g["sha256_a_le_u32"]
    # n2 = a*2^32-1 and n1 = (n2-1)/2 as in the table's other two columns
    for ai, n2, n1 in zip(a[:4], g["n2_first"], g["n1_first"]):
        assert (int(ai) << 32) - 1 == n2 and (n2 - 1) // 2 == n1
    assert np.all(np.diff(a.astype(np.int64)) < 0)
    # a slice further down is consistent with the full sequence
    assert np.array_equal(capi.safeprime_multipliers(5000, 100), a[5000:5100])


def test_oracle_safeprimes_match_rnd_txt():
    from oracle import pyoracle
    g = gold("safeprimes.json")
    a, n2, n1 = pyoracle.safeprimes(0, 1200)
    assert list(a[:32]) == g["first_32"]
    assert int(a[1000]) == g["row_1000"]
    assert [int(v) for v in n2[:4]] == g["n2_first"] and [int(v) for v in n1[:4]] == g["n1_first"]


def test_flasher_spectrum_table():
    g = gold("flasher_405nm.json")
    wl, val = ice.GetFlasherLED405Spectrum()
    np.testing.assert_allclose(wl, np.array(g["wlen_nm"]) * 1e-9, rtol=1e-15)
    np.testing.assert_array_equal(val, g["value"])
    assert np.all(np.diff(wl) > 0)


def test_cherenkov_generator_tabulated_on_bias_grid():
    """makeCherenkovWavelengthGenerator with a tabulated bias re-uses the bias binning
    (I3CLSimModuleHelper.cxx:224-256)."""
    m = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_mie", useTiltIfAvailable=False)
    b = ice.GetIceCubeDOMAcceptance()
    gen = ice.makeCherenkovWavelengthGenerator(b, False, m)
    assert gen.kind == gen.INTERP_EQUAL and len(gen.y) == 43
    assert gen.x0 == 260e-9 and gen.dx == 10e-9
    wl = 260e-9 + 7 * 10e-9
    n = m.GetPhaseRefractiveIndex(wl)
    assert gen.y[7] == pytest.approx(b.values[7] * (2.0 * math.pi / (137.0 * (wl * wl))) * (1.0 - 1.0 / (n * n)), rel=1e-14)
    # no bias + no dispersion -> the analytic generator
    flat = ice.WlenBias(constant=1.0)
    g2 = ice.makeCherenkovWavelengthGenerator(flat, True, m)
    assert g2.kind == g2.NO_DISPERSION and g2.from_wlen == 265e-9 and g2.to_wlen == 675e-9
    # no bias, dispersion -> 10 nm grid with int(range/10nm)+2 points
    g3 = ice.makeCherenkovWavelengthGenerator(flat, False, m)
    assert len(g3.y) == int((675e-9 - 265e-9) / 10e-9) + 2
    with pytest.raises(RuntimeError):
        ice.makeCherenkovWavelengthGenerator(ice.WlenBias(values=[1.0, 1.0], start_wlen=300e-9, wlen_step=10e-9), False, m)


def test_ice_loader_errors():
    with pytest.raises(RuntimeError):
        ice.MakeIceCubeMediumProperties(iceDataDirectory="no_such_model")
