// I3CLSimStepToPhotonConverterCUDA.cxx -- see the header.  Host logic only: flatten the polymorphic
// description objects into the POD clsimcu_config (the job the reference's code generators do,
// private/opencl/I3CLSimHelperGenerate{Geometry,MediumProperties}Source*.cxx, minus the source text), and
// forward the hot calls to the C ABI.  Everything numeric happens behind include/clsimcuda.h on the GPU.
#include "I3CLSimStepToPhotonConverterCUDA.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <set>
#include <sstream>
#include <typeinfo>

#include "clsimcuda.h"

#ifdef CLSIM_CUDA_IN_ICETRAY
// the concrete description classes the flattening recognises (stand-alone they are all in clsim_compat.h)
#include "clsim/I3CLSimPhoton.h"
#include "clsim/I3CLSimStep.h"
#include "clsim/function/I3CLSimFunctionAbsLenIceCube.h"
#include "clsim/function/I3CLSimFunctionConstant.h"
#include "clsim/function/I3CLSimFunctionFromTable.h"
#include "clsim/function/I3CLSimFunctionRefIndexIceCube.h"
#include "clsim/function/I3CLSimFunctionScatLenIceCube.h"
#include "clsim/function/I3CLSimScalarFieldAnisotropyAbsLenScaling.h"
#include "clsim/function/I3CLSimScalarFieldConstant.h"
#include "clsim/function/I3CLSimScalarFieldIceTiltZShift.h"
#include "clsim/function/I3CLSimVectorTransformConstant.h"
#include "clsim/function/I3CLSimVectorTransformMatrix.h"
#include "clsim/random_value/I3CLSimRandomValueConstant.h"
#include "clsim/random_value/I3CLSimRandomValueHenyeyGreenstein.h"
#include "clsim/random_value/I3CLSimRandomValueInterpolatedDistribution.h"
#include "clsim/random_value/I3CLSimRandomValueMixed.h"
#include "clsim/random_value/I3CLSimRandomValueSimplifiedLiu.h"
#include "clsim/random_value/I3CLSimRandomValueWlenCherenkovNoDispersion.h"
#endif

static_assert(sizeof(I3CLSimStep) == sizeof(clsimcu_step), "step record layouts must agree");
static_assert(sizeof(I3CLSimPhoton) == sizeof(clsimcu_photon), "photon record layouts must agree");

const bool I3CLSimStepToPhotonConverterCUDA::default_useNativeMath = true;

namespace {
typedef I3CLSimStepToPhotonConverter_exception Err;

#ifdef CLSIM_CUDA_IN_ICETRAY
// (IceTray's pointer typedefs are boost::shared_ptr)
template <class T, class B> boost::shared_ptr<const T> as(const boost::shared_ptr<const B> &p) { return boost::dynamic_pointer_cast<const T>(p); }
template <class B> std::string class_name(const boost::shared_ptr<const B> &p)
#else
template <class T, class B> std::shared_ptr<const T> as(const std::shared_ptr<const B> &p) { return std::dynamic_pointer_cast<const T>(p); }
template <class B> std::string class_name(const std::shared_ptr<const B> &p)
#endif
{
    if (!p) return "(null)";
    const B &ref = *p;
    return typeid(ref).name();
}

void throw_capi(const char *what)
{
    throw Err(std::string(what) + ": " + clsimcu_last_error());
}
} // namespace

// The arrays the config points into.
struct I3CLSimStepToPhotonConverterCUDA::Flat {
    clsimcu_config config;
    std::vector<clsimcu_wlen_generator> generators;
    std::vector<std::vector<double> > generatorX, generatorY;
    std::vector<double> bias;
    std::vector<double> aDust400, deltaTau, b400;
    std::vector<double> tiltDist, tiltCorr;
    std::vector<int32_t> stringID, subdetector;
    std::vector<uint32_t> domID;
    std::vector<double> x, y, z;
};

I3CLSimStepToPhotonConverterCUDA::I3CLSimStepToPhotonConverterCUDA(uint64_t randomSeed, bool useNativeMath)
    : randomSeed_(randomSeed), useNativeMath_(useNativeMath), initialized_(false), compiled_(false), deviceIsSelected_(false), device_(0),
      // class defaults of the reference constructor, private/opencl/I3CLSimStepToPhotonConverterOpenCL.cxx:68-95
      enableDoubleBuffering_(false), stopDetectedPhotons_(false), saveAllPhotons_(false), saveAllPhotonsPrescale_(0.001),
      fixedNumberOfAbsorptionLengths_(std::numeric_limits<double>::quiet_NaN()), pancakeFactor_(1.0), photonHistoryEntries_(0), workgroupSize_(0),
      maxNumWorkitems_(10240), firstRNGMultiplierRow_(0), flat_(nullptr), engine_(nullptr)
{
}

I3CLSimStepToPhotonConverterCUDA::~I3CLSimStepToPhotonConverterCUDA()
{
    // interrupts and joins the worker like …OpenCL.cxx:110-145
    if (engine_) clsimcu_destroy(engine_);
    delete flat_;
}

void I3CLSimStepToPhotonConverterCUDA::ThrowIfInitialized() const
{
    if (initialized_) throw Err("I3CLSimStepToPhotonConverterCUDA already initialized!");
}
void I3CLSimStepToPhotonConverterCUDA::ThrowIfNotInitialized() const
{
    if (!initialized_) throw Err("I3CLSimStepToPhotonConverterCUDA is not initialized!");
}

// ---- setters (…OpenCL.cxx:1324-1523: each throws after Initialize and invalidates a previous Compile) ----
#define CLSIM_SETTER_PROLOGUE() \
    ThrowIfInitialized();       \
    compiled_ = false

void I3CLSimStepToPhotonConverterCUDA::SetDevice(int cudaOrdinal)
{
    CLSIM_SETTER_PROLOGUE();
    if (cudaOrdinal < 0) throw Err("Invalid CUDA device ordinal!");
    device_ = cudaOrdinal;
    deviceIsSelected_ = true;
}
void I3CLSimStepToPhotonConverterCUDA::SetEnableDoubleBuffering(bool value) { CLSIM_SETTER_PROLOGUE(); enableDoubleBuffering_ = value; }
bool I3CLSimStepToPhotonConverterCUDA::GetEnableDoubleBuffering() const { return enableDoubleBuffering_; }
void I3CLSimStepToPhotonConverterCUDA::SetDoublePrecision(bool value)
{
    CLSIM_SETTER_PROLOGUE();
    if (value) throw Err("DoublePrecision is not available in I3CLSimStepToPhotonConverterCUDA (single precision only)");
}
bool I3CLSimStepToPhotonConverterCUDA::GetDoublePrecision() const { return false; }
void I3CLSimStepToPhotonConverterCUDA::SetStopDetectedPhotons(bool value) { CLSIM_SETTER_PROLOGUE(); stopDetectedPhotons_ = value; }
bool I3CLSimStepToPhotonConverterCUDA::GetStopDetectedPhotons() const { return stopDetectedPhotons_; }
void I3CLSimStepToPhotonConverterCUDA::SetSaveAllPhotons(bool value) { CLSIM_SETTER_PROLOGUE(); saveAllPhotons_ = value; }
bool I3CLSimStepToPhotonConverterCUDA::GetSaveAllPhotons() const { return saveAllPhotons_; }
void I3CLSimStepToPhotonConverterCUDA::SetSaveAllPhotonsPrescale(double value) { CLSIM_SETTER_PROLOGUE(); saveAllPhotonsPrescale_ = value; }
double I3CLSimStepToPhotonConverterCUDA::GetSaveAllPhotonsPrescale() const { return saveAllPhotonsPrescale_; }
void I3CLSimStepToPhotonConverterCUDA::SetFixedNumberOfAbsorptionLengths(double value) { CLSIM_SETTER_PROLOGUE(); fixedNumberOfAbsorptionLengths_ = value; }
double I3CLSimStepToPhotonConverterCUDA::GetFixedNumberOfAbsorptionLengths() const { return fixedNumberOfAbsorptionLengths_; }
void I3CLSimStepToPhotonConverterCUDA::SetDOMPancakeFactor(double value) { CLSIM_SETTER_PROLOGUE(); pancakeFactor_ = value; }
double I3CLSimStepToPhotonConverterCUDA::GetDOMPancakeFactor() const { return pancakeFactor_; }
void I3CLSimStepToPhotonConverterCUDA::SetPhotonHistoryEntries(uint32_t value) { CLSIM_SETTER_PROLOGUE(); photonHistoryEntries_ = value; }
uint32_t I3CLSimStepToPhotonConverterCUDA::GetPhotonHistoryEntries() const { return photonHistoryEntries_; }
void I3CLSimStepToPhotonConverterCUDA::SetFirstRNGMultiplierRow(uint64_t row) { CLSIM_SETTER_PROLOGUE(); firstRNGMultiplierRow_ = row; }

void I3CLSimStepToPhotonConverterCUDA::SetWlenGenerators(const std::vector<I3CLSimRandomValueConstPtr> &wlenGenerators)
{
    CLSIM_SETTER_PROLOGUE();
    wlenGenerators_ = wlenGenerators;
}
void I3CLSimStepToPhotonConverterCUDA::SetWlenBias(I3CLSimFunctionConstPtr wlenBias) { CLSIM_SETTER_PROLOGUE(); wlenBias_ = wlenBias; }
void I3CLSimStepToPhotonConverterCUDA::SetMediumProperties(I3CLSimMediumPropertiesConstPtr mediumProperties)
{
    CLSIM_SETTER_PROLOGUE();
    mediumProperties_ = mediumProperties;
}
void I3CLSimStepToPhotonConverterCUDA::SetGeometry(I3CLSimSimpleGeometryConstPtr geometry) { CLSIM_SETTER_PROLOGUE(); geometry_ = geometry; }

std::size_t I3CLSimStepToPhotonConverterCUDA::GetMaxWorkgroupSize() const
{
    // …OpenCL.cxx:1345-1358 requires a compiled kernel
    if (!compiled_) throw Err("You need to compile the kernel first. Call Compile().");
    return 1024;
}
void I3CLSimStepToPhotonConverterCUDA::SetWorkgroupSize(std::size_t val)
{
    ThrowIfInitialized();
    if (!compiled_) throw Err("You need to compile the kernel first. Call Compile().");
    if (val > 1024) throw Err("Workgroup size too large!");
    workgroupSize_ = val; // 0 = "default", advertised as 1: bunch sizes are not tied to a thread block here
}
void I3CLSimStepToPhotonConverterCUDA::SetMaxNumWorkitems(std::size_t val)
{
    ThrowIfInitialized();
    if (val <= 0) throw Err("Invalid maximum number of work items!");
    maxNumWorkitems_ = val;
}
std::size_t I3CLSimStepToPhotonConverterCUDA::GetWorkgroupSize() const
{
    if (initialized_) {
        size_t v = 0;
        if (clsimcu_workgroup_size(engine_, &v) != CLSIMCU_OK) throw_capi("GetWorkgroupSize");
        return v;
    }
    return workgroupSize_ == 0 ? 1 : workgroupSize_;
}
std::size_t I3CLSimStepToPhotonConverterCUDA::GetMaxNumWorkitems() const { return maxNumWorkitems_; }

// ---- Compile = validate + flatten ---------------------------------------------------------------------
void I3CLSimStepToPhotonConverterCUDA::Compile()
{
    ThrowIfInitialized();
    if (compiled_) return; // silently, like the reference
    // same checks in the same order as …OpenCL.cxx:492-508
    if (wlenGenerators_.empty()) throw Err("WlenGenerators not set!");
    if (!wlenBias_) throw Err("WlenBias not set!");
    if (!mediumProperties_) throw Err("MediumProperties not set!");
    if (!geometry_) throw Err("Geometry not set!");
    if (!deviceIsSelected_) throw Err("Device not selected!");
    if (saveAllPhotons_ && stopDetectedPhotons_)
        throw Err("Internal error: both the saveAllPhotons and stopDetectedPhotons options are set at the same time.");
    Flatten();
    compiled_ = true;
}

void I3CLSimStepToPhotonConverterCUDA::Flatten()
{
    delete flat_;
    flat_ = new Flat();
    Flat &f = *flat_;
    clsimcu_config &c = f.config;
    std::memset(&c, 0, sizeof(c));
    c.struct_size = static_cast<int32_t>(sizeof(clsimcu_config));
    c.device = device_;
    // every option of the path runs on the fast kernel (photon history, non-stopping detection included); the
    // reference-order kernel is the precise-math twin (useNativeMath = false)
    c.kernel_mode = useNativeMath_ ? CLSIMCU_KERNEL_FAST : CLSIMCU_KERNEL_REFERENCE;
    c.enable_double_buffering = enableDoubleBuffering_ ? 1 : 0;
    c.stop_detected_photons = stopDetectedPhotons_ ? 1 : 0;
    c.save_all_photons = saveAllPhotons_ ? 1 : 0;
    c.photon_history_entries = static_cast<int32_t>(photonHistoryEntries_);
    c.save_all_photons_prescale = saveAllPhotonsPrescale_;
    c.fixed_number_of_absorption_lengths = fixedNumberOfAbsorptionLengths_;
    c.pancake_factor = pancakeFactor_;
    c.rng_seed = randomSeed_;
    c.rng_first_multiplier = firstRNGMultiplierRow_;

    // -- wavelength generators (R3a)
    const std::size_t ng = wlenGenerators_.size();
    f.generators.resize(ng);
    f.generatorX.resize(ng);
    f.generatorY.resize(ng);
    for (std::size_t i = 0; i < ng; ++i) {
        clsimcu_wlen_generator &g = f.generators[i];
        std::memset(&g, 0, sizeof(g));
        const I3CLSimRandomValueConstPtr &p = wlenGenerators_[i];
        if (auto d = as<I3CLSimRandomValueInterpolatedDistribution>(p)) {
            f.generatorY[i] = d->GetY();
            g.n = static_cast<int32_t>(f.generatorY[i].size());
            g.y = f.generatorY[i].data();
            if (d->GetConstantXSpacing()) {
                g.kind = CLSIMCU_WLEN_INTERP_EQUAL;
                g.x0 = d->GetFirstX();
                g.dx = d->GetXSpacing();
            } else {
                g.kind = CLSIMCU_WLEN_INTERP_UNEQUAL;
                f.generatorX[i] = d->GetX();
                g.x = f.generatorX[i].data();
            }
        } else if (auto d = as<I3CLSimRandomValueWlenCherenkovNoDispersion>(p)) {
            g.kind = CLSIMCU_WLEN_NO_DISPERSION;
            g.from_wlen = d->GetFromWlen();
            g.to_wlen = d->GetToWlen();
        } else if (auto d = as<I3CLSimRandomValueConstant>(p)) {
            g.kind = CLSIMCU_WLEN_CONSTANT;
            g.value = d->GetValue();
        } else {
            throw Err("wavelength generator #" + std::to_string(i) + " is of a class the CUDA converter does not know: " + class_name(p));
        }
    }
    c.wlen_generators = f.generators.data();
    c.num_wlen_generators = static_cast<int32_t>(ng);

    // -- wavelength bias (R3b)
    if (auto t = as<I3CLSimFunctionFromTable>(wlenBias_)) {
        if (!t->GetInEqualSpacingMode()) throw Err("WlenBias: I3CLSimFunctionFromTable must be in equal spacing mode for the CUDA converter");
        c.wlen_bias.kind = CLSIMCU_BIAS_TABLE;
        c.wlen_bias.n = static_cast<int32_t>(t->GetNumEntries());
        c.wlen_bias.x0 = t->GetFirstWavelength();
        c.wlen_bias.dx = t->GetWavelengthStepping();
        for (std::size_t i = 0; i < t->GetNumEntries(); ++i) f.bias.push_back(t->GetEntryValue(i));
        c.wlen_bias.v = f.bias.data();
    } else if (auto k = as<I3CLSimFunctionConstant>(wlenBias_)) {
        c.wlen_bias.kind = CLSIMCU_BIAS_CONSTANT;
        c.wlen_bias.value = k->GetValue(400e-9);
    } else {
        throw Err("WlenBias is of a class the CUDA converter does not know: " + class_name(wlenBias_));
    }

    // -- medium (R4, R4a, R4b, R9)
    const I3CLSimMediumProperties &mp = *mediumProperties_;
    if (!mp.IsReady()) throw Err("MediumProperties are not ready (a layer or the scattering angle distribution is missing)!");
    clsimcu_medium &m = c.medium;
    const uint32_t nl = mp.GetLayersNum();
    m.num_layers = static_cast<int32_t>(nl);
    m.layers_zstart = mp.GetLayersZStart();
    m.layers_height = mp.GetLayersHeight();
    for (uint32_t l = 0; l < nl; ++l) {
        auto a = as<I3CLSimFunctionAbsLenIceCube>(mp.GetAbsorptionLength(l));
        auto s = as<I3CLSimFunctionScatLenIceCube>(mp.GetScatteringLength(l));
        auto np = as<I3CLSimFunctionRefIndexIceCube>(mp.GetPhaseRefractiveIndex(l));
        auto ng_ = as<I3CLSimFunctionRefIndexIceCube>(mp.GetGroupRefractiveIndexOverride(l));
        if (!a) throw Err("absorption length of layer " + std::to_string(l) + " is not an I3CLSimFunctionAbsLenIceCube: " + class_name(mp.GetAbsorptionLength(l)));
        if (!s) throw Err("scattering length of layer " + std::to_string(l) + " is not an I3CLSimFunctionScatLenIceCube: " + class_name(mp.GetScatteringLength(l)));
        if (!np || np->GetMode() != "phase") throw Err("phase refractive index of layer " + std::to_string(l) + " is not an I3CLSimFunctionRefIndexIceCube(\"phase\")");
        if (!ng_ || ng_->GetMode() != "group")
            throw Err("group refractive index override of layer " + std::to_string(l) + " is not an I3CLSimFunctionRefIndexIceCube(\"group\")");
        if (l == 0) {
            m.kappa = a->GetKappa(); m.A = a->GetA(); m.B = a->GetB(); m.D = a->GetD(); m.E = a->GetE();
            m.alpha = s->GetAlpha();
            for (int i = 0; i < 5; ++i) {
                m.n_phase[i] = np->GetPhaseCoefficient(i);
                m.n_group[i] = ng_->GetGroupCoefficient(i);
            }
        } else {
            // the optimised generators of the reference demand the same (…MediumPropertiesSource_Optimizers.cxx:123-250)
            if (a->GetKappa() != m.kappa || a->GetA() != m.A || a->GetB() != m.B || a->GetD() != m.D || a->GetE() != m.E || s->GetAlpha() != m.alpha)
                throw Err("layer " + std::to_string(l) + ": kappa/A/B/D/E/alpha must be the same in all layers");
        }
        for (int i = 0; i < 5; ++i)
            if (np->GetPhaseCoefficient(i) != m.n_phase[i] || ng_->GetPhaseCoefficient(i) != m.n_phase[i] || ng_->GetGroupCoefficient(i) != m.n_group[i])
                throw Err("layer " + std::to_string(l) + ": refractive index coefficients must be the same in all layers");
        f.aDust400.push_back(a->GetADust400());
        f.deltaTau.push_back(a->GetDeltaTau());
        f.b400.push_back(s->GetB400());
    }
    m.a_dust400 = f.aDust400.data();
    m.delta_tau = f.deltaTau.data();
    m.b400 = f.b400.data();

    const I3CLSimRandomValueConstPtr cosAngle = mp.GetScatteringCosAngleDistribution();
    if (auto mix = as<I3CLSimRandomValueMixed>(cosAngle)) {
        auto sl = as<I3CLSimRandomValueSimplifiedLiu>(mix->GetFirstDistribution());
        auto hg = as<I3CLSimRandomValueHenyeyGreenstein>(mix->GetSecondDistribution());
        if (!sl || !hg || sl->GetMeanCosine() != hg->GetMeanCosine())
            throw Err("scattering angle: only Mixed(f, SimplifiedLiu(g), HenyeyGreenstein(g)) is known to the CUDA converter");
        m.scat_kind = CLSIMCU_SCAT_MIXED_SL_HG;
        m.f_sl = mix->GetFractionOfFirstDistribution();
        m.mean_cos = hg->GetMeanCosine();
    } else if (auto hg = as<I3CLSimRandomValueHenyeyGreenstein>(cosAngle)) {
        m.scat_kind = CLSIMCU_SCAT_HG;
        m.mean_cos = hg->GetMeanCosine();
    } else if (auto sl = as<I3CLSimRandomValueSimplifiedLiu>(cosAngle)) {
        m.scat_kind = CLSIMCU_SCAT_SL;
        m.mean_cos = sl->GetMeanCosine();
    } else {
        throw Err("scattering angle distribution is of a class the CUDA converter does not know: " + class_name(cosAngle));
    }

    const I3CLSimScalarFieldConstPtr tilt = mp.GetIceTiltZShift();
    if (auto t = as<I3CLSimScalarFieldIceTiltZShift>(tilt)) {
        f.tiltDist = t->GetDistancesFromOriginAlongTilt();
        const std::size_t nz = t->GetZCoordinates().size();
        for (std::size_t i = 0; i < f.tiltDist.size(); ++i)
#ifdef CLSIM_CUDA_IN_ICETRAY
            for (std::size_t k = 0; k < nz; ++k) f.tiltCorr.push_back(t->GetZCorrections()(i, k));   // I3Matrix (ublas)
#else
            for (std::size_t k = 0; k < nz; ++k) f.tiltCorr.push_back(t->GetZCorrections()[i][k]);
#endif
        m.tilt_num_dist = static_cast<int32_t>(f.tiltDist.size());
        m.tilt_num_z = static_cast<int32_t>(nz);
        m.tilt_dist = f.tiltDist.data();
        m.tilt_corr = f.tiltCorr.data();
        m.tilt_z0 = t->GetFirstZCoordinate();
        m.tilt_dz = t->GetZCoordinateSpacing();
        m.tilt_azimuth = t->GetDirectionOfTiltAzimuth();
    } else if (auto k = as<I3CLSimScalarFieldConstant>(tilt)) {
        if (k->GetValue(0, 0, 0) != 0.0) throw Err("a constant non-zero ice tilt shift is not supported by the CUDA converter (shift the layer table instead)");
    } else {
        throw Err("ice tilt is of a class the CUDA converter does not know: " + class_name(tilt));
    }

    const I3CLSimScalarFieldConstPtr absCorr = mp.GetDirectionalAbsorptionLengthCorrection();
    const I3CLSimVectorTransformConstPtr pre = mp.GetPreScatterDirectionTransform(), post = mp.GetPostScatterDirectionTransform();
    auto an = as<I3CLSimScalarFieldAnisotropyAbsLenScaling>(absCorr);
    auto preM = as<I3CLSimVectorTransformMatrix>(pre), postM = as<I3CLSimVectorTransformMatrix>(post);
    if (an || preM || postM) {
        if (!(an && preM && postM))
            throw Err("anisotropy needs all of: I3CLSimScalarFieldAnisotropyAbsLenScaling and two I3CLSimVectorTransformMatrix (pre, post)");
        m.has_anisotropy = 1;
        m.aniso_azimuth = an->GetAnisotropyDirAzimuth();
        m.aniso_along = an->GetMagnitudeAlongDir();
        m.aniso_perp = an->GetMagnitudePerpToDir();
        for (int r = 0; r < 3; ++r)
            for (int col = 0; col < 3; ++col) {
                m.pre_matrix[3 * r + col] = preM->GetMatrixElement(r, col);
                m.post_matrix[3 * r + col] = postM->GetMatrixElement(r, col);
            }
        m.pre_renormalize = preM->GetRenormalize() ? 1 : 0;
        m.post_renormalize = postM->GetRenormalize() ? 1 : 0;
    } else {
        auto k = as<I3CLSimScalarFieldConstant>(absCorr);
        if (!k || k->GetValue(0, 0, 0) != 1.0 || !as<I3CLSimVectorTransformConstant>(pre) || !as<I3CLSimVectorTransformConstant>(post))
            throw Err("directional absorption correction / direction transforms are of classes the CUDA converter does not know");
    }

    // -- geometry (R7).  Subdetector = rank of the name in sorted order (the reference keys a std::set<std::string>,
    //    I3CLSimHelperGenerateGeometrySource.cxx:737-760).
    const I3CLSimSimpleGeometry &geo = *geometry_;
    const std::size_t nd = geo.size();
    f.stringID = geo.GetStringIDVector();
    f.domID = geo.GetDomIDVector();
    f.x = geo.GetPosXVector();
    f.y = geo.GetPosYVector();
    f.z = geo.GetPosZVector();
    const std::vector<std::string> &sub = geo.GetSubdetectorVector();
    const std::set<std::string> names(sub.begin(), sub.end());
    f.subdetector.resize(nd);
    for (std::size_t i = 0; i < nd; ++i) f.subdetector[i] = static_cast<int32_t>(std::distance(names.begin(), names.find(sub[i])));
    c.geometry.num_doms = static_cast<int32_t>(nd);
    c.geometry.string_id = f.stringID.data();
    c.geometry.dom_id = f.domID.data();
    c.geometry.x = f.x.data();
    c.geometry.y = f.y.data();
    c.geometry.z = f.z.data();
    c.geometry.subdetector = f.subdetector.data();
    c.geometry.om_radius = geo.GetOMRadius();
}

std::string I3CLSimStepToPhotonConverterCUDA::DescribeTables() const
{
    if (!compiled_ && !initialized_) throw Err("You need to compile the kernel first. Call Compile().");
    clsimcu_config c = flat_->config;
    c.max_num_workitems = maxNumWorkitems_;
    c.workgroup_size = static_cast<uint32_t>(workgroupSize_);
    size_t needed = 0;
    clsimcu_describe_tables_from_config(&c, nullptr, 0, &needed);
    std::string out(needed + 1, '\0');
    if (clsimcu_describe_tables_from_config(&c, &out[0], out.size(), &needed) != CLSIMCU_OK) throw_capi("DescribeTables");
    out.resize(std::strlen(out.c_str()));
    return out;
}

// ---- life cycle -------------------------------------------------------------------------------------
void I3CLSimStepToPhotonConverterCUDA::Initialize()
{
    ThrowIfInitialized();
    Compile();
    clsimcu_config &c = flat_->config;
    c.max_num_workitems = maxNumWorkitems_;
    c.workgroup_size = static_cast<uint32_t>(workgroupSize_);
    // no CPU fallback: a missing device, a missing library symbol or an unsupported option is an exception
    if (clsimcu_create(&c, &engine_) != CLSIMCU_OK) {
        engine_ = nullptr;
        throw_capi("I3CLSimStepToPhotonConverterCUDA::Initialize");
    }
    initialized_ = true;
}

bool I3CLSimStepToPhotonConverterCUDA::IsInitialized() const { return initialized_; }

void I3CLSimStepToPhotonConverterCUDA::EnqueueSteps(I3CLSimStepSeriesConstPtr steps, uint32_t identifier)
{
    // checks and messages of …OpenCL.cxx:1525-1544
    ThrowIfNotInitialized();
    if (!steps) throw Err("Steps pointer is (null)!");
    if (steps->empty()) throw Err("Steps are empty!");
    if (steps->size() > maxNumWorkitems_) throw Err("Number of steps is greater than maximum number of work items!");
    if (steps->size() % GetWorkgroupSize() != 0) throw Err("The number of steps is not a multiple of the workgroup size!");
    // the library copies the records into pinned staging before returning, so `steps` need not be kept alive
    if (clsimcu_enqueue(engine_, reinterpret_cast<const clsimcu_step *>(steps->data()), steps->size(), identifier) != CLSIMCU_OK)
        throw_capi("EnqueueSteps");
}

std::size_t I3CLSimStepToPhotonConverterCUDA::QueueSize() const
{
    ThrowIfNotInitialized();
    size_t v = 0;
    if (clsimcu_queue_size(engine_, &v) != CLSIMCU_OK) throw_capi("QueueSize");
    return v;
}

bool I3CLSimStepToPhotonConverterCUDA::MorePhotonsAvailable() const
{
    ThrowIfNotInitialized();
    int v = 0;
    if (clsimcu_more_photons_available(engine_, &v) != CLSIMCU_OK) throw_capi("MorePhotonsAvailable");
    return v != 0;
}

I3CLSimStepToPhotonConverter::ConversionResult_t I3CLSimStepToPhotonConverterCUDA::GetConversionResult()
{
    std::vector<clsimcu_mcpe> unused;
    return GetConversionResultWithMCPEs(unused);
}

I3CLSimStepToPhotonConverter::ConversionResult_t I3CLSimStepToPhotonConverterCUDA::GetConversionResultWithMCPEs(std::vector<clsimcu_mcpe> &mcpes)
{
    ThrowIfNotInitialized();
    clsimcu_result r;
    std::memset(&r, 0, sizeof(r));
    if (clsimcu_get_result(engine_, &r) != CLSIMCU_OK) throw_capi("GetConversionResult");
    ConversionResult_t out(r.identifier);
    // photons is never NULL (I3CLSimClientModule.cxx:589); string/OM IDs are already real IDs
    // (the library rewrites indices on the device; the reference does it here, …OpenCL.cxx:1604-1619)
    const I3CLSimPhoton *first = reinterpret_cast<const I3CLSimPhoton *>(r.photons);
    out.photons = I3CLSimPhotonSeriesPtr(new I3CLSimPhotonSeries(first, first + r.num_photons));
    if (photonHistoryEntries_ > 0 && r.history) {
        // rows are in forward order, the unused ones NaN (…OpenCL.cxx:940-989 does the ring-buffer unrolling)
        I3CLSimPhotonHistorySeriesPtr hs(new I3CLSimPhotonHistorySeries(r.num_photons));
        for (std::size_t i = 0; i < r.num_photons; ++i) {
            const uint32_t n = std::min<uint32_t>((*out.photons)[i].GetNumScatters(), photonHistoryEntries_);
            const float *row = r.history + i * photonHistoryEntries_ * 4;
            for (uint32_t j = 0; j < n; ++j) (*hs)[i].push_back(row[4 * j], row[4 * j + 1], row[4 * j + 2], row[4 * j + 3]);
        }
        out.photonHistories = hs;
    }
    mcpes.assign(r.mcpes, r.mcpes + r.num_mcpes);
    clsimcu_release_result(engine_, &r);
    return out;
}

std::map<std::string, double> I3CLSimStepToPhotonConverterCUDA::GetStatistics() const
{
    std::map<std::string, double> summary;
    if (!initialized_) return summary;
    double v[8];
    if (clsimcu_get_statistics(engine_, v) != CLSIMCU_OK) throw_capi("GetStatistics");
    // keys of …OpenCL.cxx:1621-1640
    static const char *keys[8] = {"TotalDeviceTime", "TotalHostTime", "NumKernelCalls", "TotalNumPhotonsGenerated",
                                  "TotalNumPhotonsAtDOMs", "AverageDeviceTimePerPhoton", "AverageHostTimePerPhoton", "DeviceUtilization"};
    for (int i = 0; i < 8; ++i) summary[keys[i]] = v[i];
    return summary;
}

// ---- factory (I3CLSimModuleHelper.cxx:303-372) ------------------------------------------------------
std::shared_ptr<I3CLSimStepToPhotonConverterCUDA>
I3CLSimModuleHelper::initializeCUDA(const I3CLSimCUDADevice &device, uint64_t randomSeed, I3CLSimSimpleGeometryConstPtr geometry,
                                    I3CLSimMediumPropertiesConstPtr medium, I3CLSimFunctionConstPtr wavelengthGenerationBias,
                                    const std::vector<I3CLSimRandomValueConstPtr> &wavelengthGenerators, bool enableDoubleBuffering,
                                    bool doublePrecision, bool stopDetectedPhotons, bool saveAllPhotons, double saveAllPhotonsPrescale,
                                    double fixedNumberOfAbsorptionLengths, double pancakeFactor, uint32_t photonHistoryEntries,
                                    uint32_t limitWorkgroupSize, uint64_t firstRNGMultiplierRow)
{
    std::shared_ptr<I3CLSimStepToPhotonConverterCUDA> conv(new I3CLSimStepToPhotonConverterCUDA(randomSeed, device.useNativeMath));
    conv->SetDevice(device.ordinal);
    conv->SetWlenGenerators(wavelengthGenerators);
    conv->SetWlenBias(wavelengthGenerationBias);
    conv->SetMediumProperties(medium);
    conv->SetGeometry(geometry);
    conv->SetEnableDoubleBuffering(enableDoubleBuffering);
    conv->SetDoublePrecision(doublePrecision);
    conv->SetStopDetectedPhotons(stopDetectedPhotons);
    conv->SetSaveAllPhotons(saveAllPhotons);
    conv->SetSaveAllPhotonsPrescale(saveAllPhotonsPrescale);
    conv->SetFixedNumberOfAbsorptionLengths(fixedNumberOfAbsorptionLengths);
    conv->SetDOMPancakeFactor(pancakeFactor);
    conv->SetPhotonHistoryEntries(photonHistoryEntries);
    conv->SetFirstRNGMultiplierRow(firstRNGMultiplierRow);
    conv->Compile();
    std::size_t maxWorkgroupSize = conv->GetMaxWorkgroupSize();
    if (limitWorkgroupSize != 0) maxWorkgroupSize = std::min<std::size_t>(limitWorkgroupSize, maxWorkgroupSize);
    // bunch sizes are not tied to thread blocks on this backend: granularity 1 unless the caller limits it
    conv->SetWorkgroupSize(limitWorkgroupSize == 0 ? 1 : maxWorkgroupSize);
    const std::size_t workgroupSize = conv->GetWorkgroupSize();
    // use approximately the given number of work items, convert to a multiple of the workgroup size
    std::size_t maxNumWorkitems = (device.approximateNumberOfWorkItems / workgroupSize) * workgroupSize;
    if (maxNumWorkitems == 0) maxNumWorkitems = workgroupSize;
    conv->SetMaxNumWorkitems(maxNumWorkitems);
    conv->Initialize();
    return conv;
}
