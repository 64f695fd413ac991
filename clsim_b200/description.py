"""Model-level description objects and their flattening to the C ABI structs.

These are the Python twins of the reference's description objects that reach the
step->photon converter: ``I3CLSimMediumProperties`` (public/clsim/I3CLSimMediumProperties.h),
``I3CLSimSimpleGeometry`` (public/clsim/I3CLSimSimpleGeometry.h), the wavelength generators
(``I3CLSimRandomValue*``) and the wavelength bias (``I3CLSimFunctionFromTable`` /
``I3CLSimFunctionConstant``).  They hold doubles exactly like the reference objects; all
float rounding and table building happens inside ``libclsimcuda`` (the C++ side, as in the
reference's generators).

``ctypes`` mirrors of ``include/clsimcuda.h`` live here as well.
"""
import ctypes as C
import math

import numpy as np

# ------------------------------------------------------------------------------ wire formats
STEP_DTYPE = np.dtype(
    [
        ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("t", "<f4"),
        ("theta", "<f4"), ("phi", "<f4"), ("length", "<f4"), ("beta", "<f4"),
        ("num_photons", "<u4"), ("weight", "<f4"), ("identifier", "<u4"),
        ("source_type", "u1"), ("dummy1", "u1"), ("dummy2", "<u2"),
    ]
)
PHOTON_DTYPE = np.dtype(
    [
        ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("t", "<f4"),
        ("theta", "<f4"), ("phi", "<f4"), ("wavelength", "<f4"), ("cherenkov_dist", "<f4"),
        ("num_scatters", "<u4"), ("weight", "<f4"), ("identifier", "<u4"),
        ("string_id", "<i2"), ("om_id", "<u2"),
        ("start_x", "<f4"), ("start_y", "<f4"), ("start_z", "<f4"), ("start_t", "<f4"),
        ("start_theta", "<f4"), ("start_phi", "<f4"), ("group_velocity", "<f4"), ("dist_in_abs_lens", "<f4"),
    ]
)
assert STEP_DTYPE.itemsize == 48    # private/clsim/I3CLSimStep.cxx:35-37
assert PHOTON_DTYPE.itemsize == 80  # private/clsim/I3CLSimPhoton.cxx:36


# ------------------------------------------------------------------------------ ctypes structs
class WlenGeneratorStruct(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n", C.c_int32), ("x0", C.c_double), ("dx", C.c_double),
        ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)),
        ("from_wlen", C.c_double), ("to_wlen", C.c_double), ("value", C.c_double),
    ]


class WlenBiasStruct(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n", C.c_int32), ("x0", C.c_double), ("dx", C.c_double),
        ("v", C.POINTER(C.c_double)), ("value", C.c_double),
    ]


class MediumStruct(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("scat_kind", C.c_int32),
        ("layers_zstart", C.c_double), ("layers_height", C.c_double),
        ("kappa", C.c_double), ("A", C.c_double), ("B", C.c_double), ("D", C.c_double), ("E", C.c_double),
        ("a_dust400", C.POINTER(C.c_double)), ("delta_tau", C.POINTER(C.c_double)),
        ("alpha", C.c_double), ("b400", C.POINTER(C.c_double)),
        ("n_phase", C.c_double * 5), ("n_group", C.c_double * 5),
        ("f_sl", C.c_double), ("mean_cos", C.c_double),
        ("tilt_num_dist", C.c_int32), ("tilt_num_z", C.c_int32),
        ("tilt_dist", C.POINTER(C.c_double)), ("tilt_corr", C.POINTER(C.c_double)),
        ("tilt_z0", C.c_double), ("tilt_dz", C.c_double), ("tilt_azimuth", C.c_double),
        ("has_anisotropy", C.c_int32), ("pre_renormalize", C.c_int32),
        ("post_renormalize", C.c_int32), ("reserved0", C.c_int32),
        ("aniso_azimuth", C.c_double), ("aniso_along", C.c_double), ("aniso_perp", C.c_double),
        ("pre_matrix", C.c_double * 9), ("post_matrix", C.c_double * 9),
    ]


class GeometryStruct(C.Structure):
    _fields_ = [
        ("num_doms", C.c_int32), ("reserved0", C.c_int32),
        ("string_id", C.POINTER(C.c_int32)), ("dom_id", C.POINTER(C.c_uint32)),
        ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)), ("z", C.POINTER(C.c_double)),
        ("subdetector", C.POINTER(C.c_int32)), ("om_radius", C.c_double),
    ]


class ConfigStruct(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("kernel_mode", C.c_int32),
        ("enable_double_buffering", C.c_int32), ("stop_detected_photons", C.c_int32),
        ("save_all_photons", C.c_int32), ("photon_history_entries", C.c_int32),
        ("num_wlen_generators", C.c_int32),
        ("save_all_photons_prescale", C.c_double), ("fixed_number_of_absorption_lengths", C.c_double),
        ("pancake_factor", C.c_double),
        ("max_num_workitems", C.c_uint64), ("workgroup_size", C.c_uint32),
        ("output_photons_per_workitem", C.c_uint32),
        ("wlen_generators", C.POINTER(WlenGeneratorStruct)),
        ("wlen_bias", WlenBiasStruct), ("medium", MediumStruct), ("geometry", GeometryStruct),
        ("rng_n", C.c_uint64), ("rng_a", C.POINTER(C.c_uint32)), ("rng_x", C.POINTER(C.c_uint64)),
        ("rng_seed", C.c_uint64), ("rng_first_multiplier", C.c_uint64),
    ]


class ResultStruct(C.Structure):
    _fields_ = [
        ("identifier", C.c_uint32), ("reserved0", C.c_uint32), ("num_photons", C.c_size_t),
        ("photons", C.c_void_p), ("history", C.POINTER(C.c_float)),
        ("num_photons_generated", C.c_uint64), ("num_hits_counted", C.c_uint64), ("opaque", C.c_void_p),
        ("mcpes", C.c_void_p), ("num_mcpes", C.c_size_t),
    ]


KERNEL_FAST = 0
KERNEL_REFERENCE = 1


# ------------------------------------------------------------------------------ description objects
class WlenGenerator(object):
    """One entry of SetWlenGenerators (I3CLSimStepToPhotonConverter.h:90-100)."""

    INTERP_EQUAL, INTERP_UNEQUAL, NO_DISPERSION, CONSTANT = 0, 1, 2, 3

    def __init__(self, kind, x0=0.0, dx=0.0, x=None, y=None, from_wlen=0.0, to_wlen=0.0, value=0.0):
        self.kind = kind
        self.x0, self.dx = float(x0), float(dx)
        self.x = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        self.y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        self.from_wlen, self.to_wlen, self.value = float(from_wlen), float(to_wlen), float(value)

    @classmethod
    def interpolated(cls, x_first, x_spacing, y):
        """I3CLSimRandomValueInterpolatedDistribution(xFirst, xSpacing, y)."""
        return cls(cls.INTERP_EQUAL, x0=x_first, dx=x_spacing, y=y)

    @classmethod
    def interpolated_unequal(cls, x, y):
        """I3CLSimRandomValueInterpolatedDistribution(x, y)."""
        return cls(cls.INTERP_UNEQUAL, x=x, y=y)

    @classmethod
    def cherenkov_no_dispersion(cls, from_wlen, to_wlen):
        """I3CLSimRandomValueWlenCherenkovNoDispersion(fromWlen, toWlen)."""
        return cls(cls.NO_DISPERSION, from_wlen=from_wlen, to_wlen=to_wlen)

    @classmethod
    def constant(cls, value):
        """I3CLSimRandomValueConstant(value)."""
        return cls(cls.CONSTANT, value=value)


class WlenBias(object):
    """SetWlenBias argument: FromTable (equal spacing) or Constant."""

    def __init__(self, values=None, start_wlen=0.0, wlen_step=0.0, constant=None):
        self.values = None if values is None else np.ascontiguousarray(values, dtype=np.float64)
        self.start_wlen, self.wlen_step = float(start_wlen), float(wlen_step)
        self.constant = constant

    def GetMinWlen(self):
        return -math.inf if self.values is None else self.start_wlen

    def GetMaxWlen(self):
        return math.inf if self.values is None else self.start_wlen + self.wlen_step * (len(self.values) - 1)

    def GetValue(self, wlen):
        """Host double twin (I3CLSimFunctionFromTable.cxx:107-147)."""
        if self.values is None:
            return float(self.constant)
        frac, fbin = math.modf((wlen - self.start_wlen) / self.wlen_step)
        ibin = int(fbin)
        if ibin < 0 or (ibin == 0 and frac < 0):
            ibin, frac = 0, 0.0
        elif ibin >= len(self.values) - 1:
            ibin, frac = len(self.values) - 2, 1.0
        return self.values[ibin] + (self.values[ibin + 1] - self.values[ibin]) * frac


class MediumProperties(object):
    """The subset of I3CLSimMediumProperties the IceCube ice models populate
    (python/MakeIceCubeMediumProperties.py:166-230)."""

    # I3CLSimFunctionRefIndexIceCube.cxx:38-47
    DEFAULT_N_PHASE = (1.55749, -1.57988, 3.99993, -4.68271, 2.09354)
    DEFAULT_N_GROUP = (1.227106, -0.954648, 1.42568, -0.711832, 0.00000)

    def __init__(self):
        self.layersNum = 0
        self.layersZStart = 0.0
        self.layersHeight = 0.0
        self.ForcedMinWlen = 265e-9
        self.ForcedMaxWlen = 675e-9
        self.kappa = self.A = self.B = self.D = self.E = 0.0
        self.alpha = 0.0
        self.aDust400 = np.zeros(0)
        self.deltaTau = np.zeros(0)
        self.b400 = np.zeros(0)
        self.n_phase = self.DEFAULT_N_PHASE
        self.n_group = self.DEFAULT_N_GROUP
        self.scat_kind = 0
        self.fractionOfFirstDistribution = 0.0
        self.meanCosine = 0.0
        self.tilt = None        # dict(distancesFromOriginAlongTilt, zCoordinates, zCorrections, directionOfTiltAzimuth)
        self.anisotropy = None  # dict(anisotropyDirAzimuth, magnitudeAlongDir, magnitudePerpToDir)
        self.preMatrix = None
        self.postMatrix = None
        self.preRenormalize = True
        self.postRenormalize = True
        self.efficiency = 1.0

    def GetMinWavelength(self):
        return self.ForcedMinWlen

    def GetMaxWavelength(self):
        return self.ForcedMaxWlen

    # host double twins of the device functions ---------------------------------------------
    def GetPhaseRefractiveIndex(self, wlen):
        """I3CLSimFunctionRefIndexIceCube::GetValue, mode "phase" (…RefIndexIceCube.cxx:84-102)."""
        x = wlen / 1e-6
        n = self.n_phase
        return n[0] + x * (n[1] + x * (n[2] + x * (n[3] + x * n[4])))

    def GetGroupRefractiveIndex(self, wlen):
        x = wlen / 1e-6
        g = self.n_group
        return self.GetPhaseRefractiveIndex(wlen) * (g[0] + x * (g[1] + x * (g[2] + x * (g[3] + x * g[4]))))

    def GetAbsorptionLength(self, layer, wlen):
        """I3CLSimFunctionAbsLenIceCube::GetValue (…AbsLenIceCube.cxx:63-67)."""
        x = wlen / 1e-9
        # (`1.f + 0.01f*deltaTau_` in the reference: the literal is a float, 0.00999999977648...)
        return 1.0 / ((self.D * self.aDust400[layer] + self.E) * x ** (-self.kappa)
                      + self.A * math.exp(-self.B / x) * (1.0 + 0.009999999776482582 * self.deltaTau[layer]))

    def GetScatteringLength(self, layer, wlen):
        """I3CLSimFunctionScatLenIceCube::GetValue (…ScatLenIceCube.cxx:53-57)."""
        x = wlen / 1e-9
        return 1.0 / (self.b400[layer] * (x / 400.0) ** (-self.alpha))


class SimpleGeometry(object):
    """I3CLSimSimpleGeometry: flat DOM list + OM radius (oversize included)."""

    def __init__(self, string_ids, dom_ids, x, y, z, om_radius, subdetectors=None):
        self.stringIDs = np.ascontiguousarray(string_ids, dtype=np.int32)
        self.domIDs = np.ascontiguousarray(dom_ids, dtype=np.uint32)
        self.posX = np.ascontiguousarray(x, dtype=np.float64)
        self.posY = np.ascontiguousarray(y, dtype=np.float64)
        self.posZ = np.ascontiguousarray(z, dtype=np.float64)
        n = len(self.stringIDs)
        if subdetectors is None:
            subdetectors = ["Unknown"] * n  # I3CLSimSimpleGeometryFromI3Geometry.cxx:109
        names = sorted(set(subdetectors))   # std::set<std::string> order
        self.subdetectorNames = list(subdetectors)
        self.subdetectorIndex = np.ascontiguousarray([names.index(s) for s in subdetectors], dtype=np.int32)
        self.OMRadius = float(om_radius)
        assert all(len(a) == n for a in (self.domIDs, self.posX, self.posY, self.posZ, self.subdetectorIndex))

    def size(self):
        return len(self.stringIDs)


class ConverterOptions(object):
    """The setters of I3CLSimStepToPhotonConverterOpenCL the factory calls
    (private/clsim/I3CLSimModuleHelper.cxx:319-369), with the class defaults
    (…ConverterOpenCL.cxx:68-95)."""

    def __init__(self, **kw):
        self.device = 0
        self.kernel_mode = KERNEL_FAST
        self.enable_double_buffering = False
        self.stop_detected_photons = False
        self.save_all_photons = False
        self.save_all_photons_prescale = 0.001
        self.fixed_number_of_absorption_lengths = math.nan
        self.pancake_factor = 1.0
        self.photon_history_entries = 0
        self.max_num_workitems = 10240
        self.workgroup_size = 0
        self.output_photons_per_workitem = 0
        self.rng_n = 0
        self.rng_a = None
        self.rng_x = None
        self.rng_seed = 0
        self.rng_first_multiplier = 0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown converter option %r" % k)
            setattr(self, k, v)


def _dptr(arr):
    return arr.ctypes.data_as(C.POINTER(C.c_double))


def build_config(medium, geometry, wlen_generators, wlen_bias, options):
    """Flatten the description into a ``ConfigStruct``.  Returns (struct, keepalive)."""
    keep = []
    cfg = ConfigStruct()
    cfg.struct_size = C.sizeof(ConfigStruct)
    cfg.device = int(options.device)
    cfg.kernel_mode = int(options.kernel_mode)
    cfg.enable_double_buffering = int(bool(options.enable_double_buffering))
    cfg.stop_detected_photons = int(bool(options.stop_detected_photons))
    cfg.save_all_photons = int(bool(options.save_all_photons))
    cfg.photon_history_entries = int(options.photon_history_entries)
    cfg.save_all_photons_prescale = float(options.save_all_photons_prescale)
    cfg.fixed_number_of_absorption_lengths = float(options.fixed_number_of_absorption_lengths)
    cfg.pancake_factor = float(options.pancake_factor)
    cfg.max_num_workitems = int(options.max_num_workitems)
    cfg.workgroup_size = int(options.workgroup_size)
    cfg.output_photons_per_workitem = int(options.output_photons_per_workitem)

    gens = (WlenGeneratorStruct * max(1, len(wlen_generators)))()
    for i, g in enumerate(wlen_generators):
        gens[i].kind = g.kind
        gens[i].x0, gens[i].dx = g.x0, g.dx
        gens[i].from_wlen, gens[i].to_wlen, gens[i].value = g.from_wlen, g.to_wlen, g.value
        if g.y is not None:
            gens[i].n = len(g.y)
            gens[i].y = _dptr(g.y)
            keep.append(g.y)
        if g.x is not None:
            gens[i].x = _dptr(g.x)
            keep.append(g.x)
    keep.append(gens)
    cfg.wlen_generators = C.cast(gens, C.POINTER(WlenGeneratorStruct))
    cfg.num_wlen_generators = len(wlen_generators)

    if wlen_bias.values is None:
        cfg.wlen_bias.kind = 0
        cfg.wlen_bias.value = float(wlen_bias.constant)
    else:
        cfg.wlen_bias.kind = 1
        cfg.wlen_bias.n = len(wlen_bias.values)
        cfg.wlen_bias.x0, cfg.wlen_bias.dx = wlen_bias.start_wlen, wlen_bias.wlen_step
        cfg.wlen_bias.v = _dptr(wlen_bias.values)
        keep.append(wlen_bias.values)

    m = cfg.medium
    m.num_layers = int(medium.layersNum)
    m.scat_kind = int(medium.scat_kind)
    m.layers_zstart, m.layers_height = float(medium.layersZStart), float(medium.layersHeight)
    m.kappa, m.A, m.B, m.D, m.E = (float(medium.kappa), float(medium.A), float(medium.B), float(medium.D), float(medium.E))
    m.alpha = float(medium.alpha)
    for name, attr in (("a_dust400", "aDust400"), ("delta_tau", "deltaTau"), ("b400", "b400")):
        arr = np.ascontiguousarray(getattr(medium, attr), dtype=np.float64)
        assert len(arr) == medium.layersNum
        keep.append(arr)
        setattr(m, name, _dptr(arr))
    m.n_phase = (C.c_double * 5)(*medium.n_phase)
    m.n_group = (C.c_double * 5)(*medium.n_group)
    m.f_sl, m.mean_cos = float(medium.fractionOfFirstDistribution), float(medium.meanCosine)
    if medium.tilt is not None:
        dist = np.ascontiguousarray(medium.tilt["distancesFromOriginAlongTilt"], dtype=np.float64)
        zc = np.ascontiguousarray(medium.tilt["zCoordinates"], dtype=np.float64)
        corr = np.ascontiguousarray(medium.tilt["zCorrections"], dtype=np.float64)
        assert corr.shape == (len(dist), len(zc))
        # I3CLSimScalarFieldIceTiltZShift.cxx:63-88: equal spacing enforced, mean spacing used
        spacing = np.diff(zc)
        mean_spacing = float(spacing.sum() / (len(zc) - 1))
        if np.any(spacing <= 0) or np.any(np.abs(spacing - mean_spacing) > 1e-5):
            raise ValueError("zCoordinates (dimension 2) are not equally spaced / ascending")
        if np.any(np.diff(dist) <= 0):
            raise ValueError("distancesFromOriginAlongTilt (dimension 1) is not in ascending order.")
        m.tilt_num_dist, m.tilt_num_z = len(dist), len(zc)
        m.tilt_dist, m.tilt_corr = _dptr(dist), _dptr(corr)
        m.tilt_z0, m.tilt_dz = float(zc[0]), mean_spacing
        m.tilt_azimuth = float(medium.tilt["directionOfTiltAzimuth"])
        keep += [dist, corr]
    if medium.anisotropy is not None:
        m.has_anisotropy = 1
        m.aniso_azimuth = float(medium.anisotropy["anisotropyDirAzimuth"])
        m.aniso_along = float(medium.anisotropy["magnitudeAlongDir"])
        m.aniso_perp = float(medium.anisotropy["magnitudePerpToDir"])
        m.pre_matrix = (C.c_double * 9)(*np.asarray(medium.preMatrix, dtype=np.float64).ravel())
        m.post_matrix = (C.c_double * 9)(*np.asarray(medium.postMatrix, dtype=np.float64).ravel())
        m.pre_renormalize, m.post_renormalize = int(medium.preRenormalize), int(medium.postRenormalize)

    if geometry is not None:
        g = cfg.geometry
        g.num_doms = geometry.size()
        g.string_id = geometry.stringIDs.ctypes.data_as(C.POINTER(C.c_int32))
        g.dom_id = geometry.domIDs.ctypes.data_as(C.POINTER(C.c_uint32))
        g.x, g.y, g.z = _dptr(geometry.posX), _dptr(geometry.posY), _dptr(geometry.posZ)
        g.subdetector = geometry.subdetectorIndex.ctypes.data_as(C.POINTER(C.c_int32))
        g.om_radius = geometry.OMRadius
        keep.append(geometry)

    cfg.rng_n = int(options.rng_n)
    if options.rng_a is not None:
        a = np.ascontiguousarray(options.rng_a, dtype=np.uint32)
        x = np.ascontiguousarray(options.rng_x, dtype=np.uint64)
        assert len(a) >= cfg.rng_n and len(x) >= cfg.rng_n
        cfg.rng_a = a.ctypes.data_as(C.POINTER(C.c_uint32))
        cfg.rng_x = x.ctypes.data_as(C.POINTER(C.c_uint64))
        keep += [a, x]
    cfg.rng_seed = int(options.rng_seed)
    cfg.rng_first_multiplier = int(options.rng_first_multiplier)
    keep.append(medium)
    return cfg, keep
