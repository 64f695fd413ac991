"""Table-maker variant without a GPU: table geometry against closed forms, the oracle's savePath
against a physics identity, the C ABI's argument checks."""
import ctypes as C
import math

import numpy as np
import pytest

from clsim_b200 import capi, ice, mcpe, steps, tabulator
from clsim_b200.description import ConverterOptions, WlenBias
from oracle import pyoracle
from tests.scenes import rng_streams


def test_axes_shape_strides_edges_and_volumes():
    ax = tabulator.default_axes()
    assert ax.GetShape() == [202, 38, 102, 107] and ax.GetStrides() == [38 * 102 * 107, 102 * 107, 107, 1]
    assert ax.GetNBins() == 202 * 38 * 102 * 107
    r = ax.at(0)
    assert np.allclose(r.GetBinEdges(), (np.arange(201) * math.sqrt(580.0) / 200) ** 2)
    assert np.allclose(ax.at(2).GetBinEdges(), np.linspace(-1, 1, 101))
    # the spatial bins tile the sphere: half the azimuth is tabulated, every bin counts twice (Axes.cxx:125-140)
    total = sum(ax.GetBinVolume((i, j, k)) for i in range(0, 200, 7) for j in range(36) for k in range(0, 100, 9))
    full = sum(ax.GetBinVolume((i, j, k)) for i in (0, 57, 199) for j in (0, 35) for k in (0, 99))
    assert total > 0 and full > 0
    shell = sum(ax.GetBinVolume((199, j, k)) for j in range(36) for k in range(100))
    assert abs(shell - 4 * math.pi / 3 * (580.0 ** 3 - r.GetBinEdge(199) ** 3)) < 1e-6 * shell
    cyl = tabulator.default_axes(infinite_muon=True, impact_angle=True)
    assert cyl.GetShape() == [102, 38, 82, 107, 22]
    ring = sum(cyl.GetBinVolume((99, j, 3)) for j in range(36))
    assert abs(ring - math.pi * (580.0 ** 2 - cyl.at(0).GetBinEdge(99) ** 2) * 20.0) < 1e-6 * ring


def homogeneous_scene():
    medium = ice.MakeHomogeneousIceMediumProperties("spice_mie")
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651)
    gen = ice.makeCherenkovWavelengthGenerator(acc, False, medium)
    opt = ConverterOptions(stop_detected_photons=False, save_all_photons=True, save_all_photons_prescale=1.0, fixed_number_of_absorption_lengths=42.0)
    return medium, acc, gen, pyoracle.Scene(medium, None, [gen], acc, opt)


def test_oracle_path_integral_equals_absorption_length():
    """With the impact angle tabulated the entry weight is exp(-depth): summed along a photon's path in 1 m sub-steps it
    integrates to the absorption length.  Homogeneous ice, bounds far away."""
    medium, acc, gen, scene = homogeneous_scene()
    axes = tabulator.SphericalAxes([tabulator.PowerAxis(0, 1e5, 20, 2), tabulator.LinearAxis(0, 180, 6), tabulator.LinearAxis(-1, 1, 5),
                                    tabulator.PowerAxis(0, 1e7, 10, 2), tabulator.LinearAxis(-1, 1, 4)])
    bunch = steps.point_source_steps(32, 25, seed=3)
    a, x = rng_streams(len(bunch))
    bins, _, entries, x_after = scene.tabulate(axes, bunch, x, a, (0, 0, 0, 0, 0, 0, 1), n_group=1.36, n_phase=1.33)
    assert entries > 32 * 25 * 100 and not np.array_equal(x_after, x)
    assert bins.min() >= 0 and abs(bins.sum() - bins[bins > 0].sum()) == 0
    per_photon = bins.sum() / (32 * 25)
    # absorption length averaged over the generated spectrum
    wl = np.linspace(265e-9, 675e-9, 4001)
    spec = np.array([acc.GetValue(w) * (1.0 / w ** 2) * (1 - 1 / medium.GetPhaseRefractiveIndex(w) ** 2) for w in wl])
    expect = np.sum(spec * np.array([medium.GetAbsorptionLength(0, w) for w in wl])) / spec.sum()
    assert abs(per_photon / expect - 1) < 0.05, (per_photon, expect)
    # nothing in the radial or time overflow: the bounds were never reached
    full = bins.reshape(axes.GetShape())
    assert full[-1].sum() == 0 and full[:, :, :, -1].sum() == 0


def test_oracle_four_axis_table_uses_the_angular_acceptance_and_the_bounds():
    medium, acc, gen, scene = homogeneous_scene()
    ang = mcpe.GetIceCubeDOMAngularSensitivity()
    axes = tabulator.SphericalAxes([tabulator.PowerAxis(0, 300, 30, 2), tabulator.LinearAxis(0, 180, 6), tabulator.LinearAxis(-1, 1, 10),
                                    tabulator.PowerAxis(0, 2000, 20, 2)])
    bunch = steps.point_source_steps(32, 25, seed=4)
    a, x = rng_streams(len(bunch))
    bins, sq, entries, _ = scene.tabulate(axes, bunch, x, a, (0, 0, 0, 0, 0, 0, 1), 1.36, 1.33, angular_coefficients=ang.coefficients, squared=True)
    assert entries > 0 and bins.min() >= 0 and sq.min() >= 0
    full = bins.reshape(axes.GetShape())
    assert full[0].sum() == 0 and full[:, 0].sum() == 0          # no underflow in radius or azimuth
    assert full[:, :, :, -1].sum() == 0                          # photons stop at the time bound: nothing in the time overflow
    # weights are bounded by the peak of the angular acceptance
    assert sq.sum() <= 0.75 ** 2 * entries and bins.sum() < 0.75 * entries
    # the same photons into a full-azimuth table: same total, azimuth spread over twice as many bins
    axes360 = tabulator.SphericalAxes([tabulator.PowerAxis(0, 300, 30, 2), tabulator.LinearAxis(0, 360, 12), tabulator.LinearAxis(-1, 1, 10),
                                       tabulator.PowerAxis(0, 2000, 20, 2)])
    bins360, _, entries360, _ = scene.tabulate(axes360, bunch, x, a, (0, 0, 0, 0, 0, 0, 1), 1.36, 1.33, angular_coefficients=ang.coefficients)
    assert entries360 == entries and abs(bins360.sum() - bins.sum()) < 1e-9 * bins.sum()
    f360 = bins360.reshape(axes360.GetShape())
    assert f360[:, 7:13].sum() > 0.3 * bins.sum()


def test_abi_argument_checks(has_gpu):
    lib = tabulator._lib()
    h = C.c_void_p()
    medium = ice.MakeHomogeneousIceMediumProperties("spice_mie")
    acc = ice.GetIceCubeDOMAcceptance(domRadius=0.1651)
    ang = mcpe.GetIceCubeDOMAngularSensitivity()
    with pytest.raises(capi.ClsimCudaError, match="4 axes, or 5"):
        tabulator.I3CLSimStepToTableConverter(0, tabulator.SphericalAxes([tabulator.LinearAxis(0, 1, 2)] * 3), 0, False, medium, None, 0.0856, acc, ang, 1)
    bad = tabulator.default_axes()
    bad.axes[2] = tabulator.LinearAxis(1, -1, 10)
    with pytest.raises(capi.ClsimCudaError, match="max > min"):
        tabulator.I3CLSimStepToTableConverter(0, bad, 0, False, medium, None, 0.0856, acc, ang, 1)
    with pytest.raises(capi.ClsimCudaError, match="angular acceptance polynomial"):
        tabulator.I3CLSimStepToTableConverter(0, tabulator.default_axes(), 0, False, medium, None, 0.0856, acc, None, 1)
    assert lib.clsimcu_tabulator_finish(None) == -4 and lib.clsimcu_tabulator_enqueue(None, None, 0, None) == -4
    if not has_gpu:
        with pytest.raises(capi.ClsimCudaError, match="no CPU fallback"):
            tabulator.I3CLSimStepToTableConverter(0, tabulator.default_axes(), 0, False, medium, None, 0.0856, acc, ang, 1)
