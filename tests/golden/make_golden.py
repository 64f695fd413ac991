#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE's own Python in this container.

Run here (not on the GPU box; /root/reference does not travel):

    python tests/golden/make_golden.py

What it does
------------
IceTray is absent, so the reference's Python cannot be imported as shipped.  This script
installs *recording stubs* for ``icecube`` / ``I3Tray`` (classes that only remember their
constructor arguments and method calls, plus the handful of I3Units constants the
loaders use), then imports the unmodified reference modules from /root/reference/python:

* ``MakeIceCubeMediumProperties.py`` (+ ``util/GetIceTiltZShift.py``,
  ``util/GetSpiceLeaAnisotropyTransforms.py``) for every ppc-style ice directory under
  resources/ice that the path uses -> ``medium_<model>[_notilt].json``
* ``GetIceCubeDOMAcceptance.py`` -> ``dom_acceptance.json``
* ``GetIceCubeFlasherSpectrum.py`` table loader for the 405 nm LED -> ``flasher_405nm.json``
* the ppc formulas restated inside ``resources/tests/testScalarFields.py`` and
  ``testSpiceLeaTransforms.py`` (extracted with ``ast``; the scripts themselves need an
  OpenCL device) evaluated on seeded inputs -> ``ppc_formulas.json``
* the first/last rows and a digest of the 16 028-row safe-prime table ``rnd.txt``
  -> ``safeprimes.json``

It also writes the *input data* the product ships (``clsim_b200/data/ice_models.json``):
the numeric columns of the ppc ice tables, stored as JSON numbers (repr round-trips
doubles exactly), because the tables are inputs of the path, not code.
"""
import ast
import hashlib
import importlib
import json
import math
import os
import sys
import types

import numpy

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


# ----------------------------------------------------------------------------- stubs
class _Units(object):
    # icetray/I3Units.h base units: metre, nanosecond, radian, GeV
    m = meter = 1.0
    meter2 = 1.0
    cm = 1e-2
    cm3 = 1e-6
    mm = 1e-3
    ns = nanosecond = 1.0
    nanometer = 1e-9
    micrometer = 1e-6
    rad = radian = 1.0
    deg = degree = math.pi / 180.0
    g = gram = 1.0
    kg = 1e3
    GeV = 1.0
    TeV = 1e3


class Recorded(object):
    """Remembers how it was built and what was called on it."""

    def __init__(self, *args, **kwargs):
        object.__setattr__(self, "_cls", type(self).__name__)
        object.__setattr__(self, "args", args)
        object.__setattr__(self, "kwargs", kwargs)
        object.__setattr__(self, "calls", [])
        object.__setattr__(self, "attrs", {})

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def method(*a, **k):
            self.calls.append((name, a, k))

        return method

    def __setattr__(self, name, value):
        self.attrs[name] = value


_classes = {}


def _recorded_class(name):
    if name not in _classes:
        _classes[name] = type(name, (Recorded,), {})
    return _classes[name]


class I3Matrix(object):
    def __init__(self, arr):
        self.array = numpy.array(arr, dtype=float)


def _stub_module(name):
    mod = types.ModuleType(name)
    mod.__path__ = []

    def __getattr__(attr):
        if attr == "I3Units":
            return _Units
        if attr == "I3Matrix":
            return I3Matrix
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _recorded_class(attr)

    mod.__getattr__ = __getattr__
    return mod


def install_stubs():
    for name in ("icecube", "icecube.icetray", "icecube.dataclasses", "icecube.clsim", "I3Tray"):
        sys.modules[name] = _stub_module(name)
    sys.modules["icecube"].icetray = sys.modules["icecube.icetray"]
    sys.modules["icecube"].dataclasses = sys.modules["icecube.dataclasses"]
    sys.modules["icecube"].clsim = sys.modules["icecube.clsim"]
    # the reference package itself, importable as "clsim_ref"
    pkg = types.ModuleType("clsim_ref")
    pkg.__path__ = [os.path.join(REF, "python")]
    sys.modules["clsim_ref"] = pkg


# ----------------------------------------------------------------------------- helpers
def _floats(a):
    return [float(v) for v in numpy.asarray(a, dtype=float).ravel()]


def medium_to_dict(m):
    """Flatten the recorded I3CLSimMediumProperties into plain numbers."""
    out = {
        "layersNum": int(m.kwargs["layersNum"]),
        "layersZStart": float(m.kwargs["layersZStart"]),
        "layersHeight": float(m.kwargs["layersHeight"]),
        "ForcedMinWlen": float(m.attrs["ForcedMinWlen"]),
        "ForcedMaxWlen": float(m.attrs["ForcedMaxWlen"]),
    }
    n = out["layersNum"]
    absl = [None] * n
    scat = [None] * n
    for name, a, k in m.calls:
        if name == "SetAbsorptionLength":
            absl[a[0]] = a[1]
        elif name == "SetScatteringLength":
            scat[a[0]] = a[1]
        elif name == "SetScatteringCosAngleDistribution":
            mix = a[0]
            assert mix._cls == "I3CLSimRandomValueMixed"
            assert mix.kwargs["firstDistribution"]._cls == "I3CLSimRandomValueSimplifiedLiu"
            assert mix.kwargs["secondDistribution"]._cls == "I3CLSimRandomValueHenyeyGreenstein"
            out["fractionOfFirstDistribution"] = float(mix.kwargs["fractionOfFirstDistribution"])
            out["meanCosine"] = float(mix.kwargs["firstDistribution"].kwargs["meanCosine"])
            assert out["meanCosine"] == float(mix.kwargs["secondDistribution"].kwargs["meanCosine"])
        elif name == "SetDirectionalAbsorptionLengthCorrection":
            f = a[0]
            if f._cls == "I3CLSimScalarFieldConstant":
                out["anisotropy"] = None
                assert f.args == (1.0,)
            else:
                assert f._cls == "I3CLSimScalarFieldAnisotropyAbsLenScaling"
                out["anisotropy"] = {
                    "anisotropyDirAzimuth": float(f.kwargs["anisotropyDirAzimuth"]),
                    "magnitudeAlongDir": float(f.kwargs["magnitudeAlongDir"]),
                    "magnitudePerpToDir": float(f.kwargs["magnitudePerpToDir"]),
                }
        elif name in ("SetPreScatterDirectionTransform", "SetPostScatterDirectionTransform"):
            f = a[0]
            key = "pre" if "Pre" in name else "post"
            if f._cls == "I3CLSimVectorTransformConstant":
                out[key + "Matrix"] = None
            else:
                assert f._cls == "I3CLSimVectorTransformMatrix"
                out[key + "Matrix"] = _floats(f.args[0].array)
                out[key + "Renormalize"] = bool(f.kwargs["renormalize"])
        elif name == "SetIceTiltZShift":
            f = a[0]
            if f._cls == "I3CLSimScalarFieldConstant":
                out["tilt"] = None
                assert f.args == (0.0,)
            else:
                assert f._cls == "I3CLSimScalarFieldIceTiltZShift"
                out["tilt"] = {
                    "distancesFromOriginAlongTilt": _floats(f.kwargs["distancesFromOriginAlongTilt"]),
                    "zCoordinates": _floats(f.kwargs["zCoordinates"]),
                    "zCorrections": [_floats(r) for r in f.kwargs["zCorrections"].array],
                    "directionOfTiltAzimuth": float(f.kwargs["directionOfTiltAzimuth"]),
                }
    for key in ("kappa", "A", "B", "D", "E"):
        vals = set(float(f.kwargs[key]) for f in absl)
        assert len(vals) == 1
        out[key] = vals.pop()
    out["aDust400"] = [float(f.kwargs["aDust400"]) for f in absl]
    out["deltaTau"] = [float(f.kwargs["deltaTau"]) for f in absl]
    vals = set(float(f.kwargs["alpha"]) for f in scat)
    assert len(vals) == 1
    out["alpha"] = vals.pop()
    out["b400"] = [float(f.kwargs["b400"]) for f in scat]
    return out


def extract_function(path, name):
    """Compile one top-level function of a reference script without running the script."""
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"math": math, "numpy": numpy, "np": numpy, "I3Units": _Units}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def main():
    install_stubs()
    MakeMedium = importlib.import_module("clsim_ref.MakeIceCubeMediumProperties").MakeIceCubeMediumProperties
    GetAcceptance = importlib.import_module("clsim_ref.GetIceCubeDOMAcceptance").GetIceCubeDOMAcceptance

    ice_root = os.path.join(REF, "resources", "ice")
    models = ["spice_mie", "spice_lea", "spice_1", "ppc_aha_0.80"]
    ice_data = {}
    for model in models:
        d = os.path.join(ice_root, model)
        for tilt in (True, False):
            m = MakeMedium(iceDataDirectory=d, useTiltIfAvailable=tilt)
            desc = medium_to_dict(m)
            suffix = "" if tilt else "_notilt"
            if tilt or desc != json.loads(json.dumps(medium_to_dict(MakeMedium(iceDataDirectory=d, useTiltIfAvailable=True)))):
                with open(os.path.join(HERE, "medium_%s%s.json" % (model, suffix)), "w") as f:
                    json.dump(desc, f)
        # the raw numeric tables, as shipped input data for the product
        entry = {
            "icemodel_dat": [_floats(r) for r in numpy.loadtxt(os.path.join(d, "icemodel.dat"))],
            "icemodel_par": [_floats(r) for r in numpy.loadtxt(os.path.join(d, "icemodel.par"))],
            "cfg": _floats(numpy.loadtxt(os.path.join(d, "cfg.txt"))),
        }
        if os.path.isfile(os.path.join(d, "tilt.par")):
            entry["tilt_par"] = [_floats(r) for r in numpy.loadtxt(os.path.join(d, "tilt.par"))]
            entry["tilt_dat"] = [_floats(r) for r in numpy.loadtxt(os.path.join(d, "tilt.dat"))]
        ice_data[model] = entry

    acc = GetAcceptance()
    assert acc._cls == "I3CLSimFunctionFromTable"
    acceptance = {"startWlen": float(acc.args[0]), "wlenStep": float(acc.args[1]), "values": _floats(acc.args[2])}
    with open(os.path.join(HERE, "dom_acceptance.json"), "w") as f:
        json.dump(acceptance, f)
    # the raw effective-area column (the loader divides by pi*r^2): recover it for the product data
    dom_radius = 0.16510
    ice_data["_dom2007a_eff_area"] = [v * math.pi * dom_radius ** 2.0 for v in acceptance["values"]]

    # flasher LED spectrum table (GetIceCubeFlasherSpectrum.py:37-65 reads two columns)
    led = numpy.loadtxt(os.path.join(REF, "resources", "flasher_data", "flasher_led_405nm_emission_spectrum_datasheet.txt"), unpack=True)
    ice_data["_flasher_led_405nm"] = {"wlen_nm": _floats(led[0]), "value": _floats(led[1])}
    with open(os.path.join(HERE, "flasher_405nm.json"), "w") as f:
        json.dump(ice_data["_flasher_led_405nm"], f)

    # hole-ice angular acceptance (python/GetIceCubeDOMAngularSensitivity.py:30-42): the reference's function,
    # through the stubs, on the one parameterisation file its tree ships (ice-models is un-vendored); row 0 of
    # the file is the peak value the tray segment folds into the DOM efficiency (traysegments/common.py:183-184)
    GetAngular = importlib.import_module("clsim_ref.GetIceCubeDOMAngularSensitivity").GetIceCubeDOMAngularSensitivity
    as_file = os.path.join(REF, "resources", "ice", "ppc_aha_0.80", "as.holeice")
    ang = GetAngular(holeIce=as_file)
    assert ang._cls == "I3CLSimFunctionPolynomial"
    angular = {"source": os.path.relpath(as_file, REF), "peak": float(numpy.loadtxt(as_file)[0]), "coefficients": _floats(ang.args[0])}
    # known answers: the polynomial as I3CLSimFunctionPolynomial::GetValue sums it (…Polynomial.cxx:86-102)
    xs = numpy.linspace(-1.0, 1.0, 41)
    vals = []
    for x in xs:
        total, mult = angular["coefficients"][0], 1.0
        for c in angular["coefficients"][1:]:
            mult *= x
            total += c * mult
        vals.append(total)
    angular["cos"] = _floats(xs)
    angular["value"] = _floats(vals)
    with open(os.path.join(HERE, "angular_acceptance.json"), "w") as f:
        json.dump(angular, f)
    ice_data["_angsens_holeice"] = {"peak": angular["peak"], "coefficients": angular["coefficients"]}

    os.makedirs(os.path.join(REPO, "clsim_b200", "data"), exist_ok=True)
    with open(os.path.join(REPO, "clsim_b200", "data", "ice_models.json"), "w") as f:
        json.dump(ice_data, f)

    # ---- ppc formulas restated by the reference's tests, evaluated on seeded inputs
    tests_dir = os.path.join(REF, "resources", "tests")
    rng = numpy.random.default_rng(12345)
    vecs = rng.normal(size=(2000, 3))
    vecs /= numpy.sqrt((vecs ** 2).sum(1))[:, None]
    formulas = {"unit_vectors": [_floats(v) for v in vecs]}
    # SpiceLea values, the defaults of both reference test scripts (testScalarFields.py:18-20,
    # testSpiceLeaTransforms.py:18-20)
    thx, logk1, logk2 = 216.0, 0.04, -0.08
    formulas["params"] = {"thx_deg": thx, "logk1": logk1, "logk2": logk2}
    dimas = extract_function(os.path.join(tests_dir, "testScalarFields.py"), "DimasAbsLenScalingFactor")
    formulas["DimasAbsLenScalingFactor"] = _floats(dimas(vecs[:, 0], vecs[:, 1], vecs[:, 2], thx, logk1, logk2))
    azx, azy = math.cos(thx * math.pi / 180.0), math.sin(thx * math.pi / 180.0)
    k1, k2 = numpy.exp(logk1), numpy.exp(logk2)
    kz = 1.0 / (k1 * k2)
    pre = extract_function(os.path.join(tests_dir, "testSpiceLeaTransforms.py"), "evaluateVectorTransformationPPCPre")
    post = extract_function(os.path.join(tests_dir, "testSpiceLeaTransforms.py"), "evaluateVectorTransformationPPCPost")
    formulas["PPCPre"] = [_floats(pre(v, azx, azy, k1, k2, kz)) for v in vecs]
    formulas["PPCPost"] = [_floats(post(v, azx, azy, k1, k2, kz)) for v in vecs]
    # the reference's own transform builder, through the stubs
    GetT = importlib.import_module("clsim_ref.util.GetSpiceLeaAnisotropyTransforms").GetSpiceLeaAnisotropyTransforms
    _, preT, postT = GetT(anisotropyDirAzimuth=thx * _Units.deg, magnitudeAlongDir=logk1, magnitudePerpToDir=logk2)
    formulas["Cpre"] = _floats(preT.args[0].array)
    formulas["Cpost"] = _floats(postT.args[0].array)
    with open(os.path.join(HERE, "ppc_formulas.json"), "w") as f:
        json.dump(formulas, f)

    # ---- safe primes
    rnd = os.path.join(REF, "resources", "scripts", "compareToPPCredux", "test_ice_models", "lea", "rnd.txt")
    if not os.path.isfile(rnd):
        cands = []
        for root, _, files in os.walk(os.path.join(REF, "resources", "scripts", "compareToPPCredux", "test_ice_models")):
            if "rnd.txt" in files:
                cands.append(os.path.join(root, "rnd.txt"))
        rnd = sorted(cands)[0]
    rows = numpy.loadtxt(rnd, dtype=numpy.uint64)
    a = rows[:, 0]
    digest = hashlib.sha256(a.astype("<u4").tobytes()).hexdigest()
    with open(os.path.join(HERE, "safeprimes.json"), "w") as f:
        json.dump({"source": os.path.relpath(rnd, REF), "rows": int(len(a)), "first_32": [int(v) for v in a[:32]],
                   "last_8": [int(v) for v in a[-8:]], "row_1000": int(a[1000]), "row_10000": int(a[10000]),
                   "sha256_a_le_u32": digest,
                   "n2_first": [int(v) for v in rows[:4, 1]], "n1_first": [int(v) for v in rows[:4, 2]]}, f)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
