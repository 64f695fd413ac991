// ref_medium.cpp -- the reference's own medium / wavelength source GENERATORS, compiled for the host.
// TEST INFRASTRUCTURE (oracle/_ref).  Nothing under clsim_b200/ links this.
//
// What this is: the reference builds the "generated" part of its OpenCL program at run time, as text, from a tree of
// parameter objects: private/clsim/function/*.cxx (wavelength-dependent functions, scalar fields, direction
// transforms), private/clsim/random_value/*.cxx (samplers), private/clsim/I3CLSimMediumProperties.cxx (the
// container) and private/opencl/I3CLSimHelperGenerateMediumPropertiesSource{,_Optimizers}.cxx (the text generator
// with its per-layer folding).  Those 33 files are compiled here UNMODIFIED, read where they lie (unity build:
// each is an #include below); the IceTray / boost headers they name are the stand-ins under oracle/ref_shim/host/.
//
// The extern "C" functions at the bottom put the objects together the way the reference's Python does
// (python/MakeIceCubeMediumProperties.py:185-244, python/util/__init__.py GetSpiceLeaAnisotropyTransforms /
// GetIceTiltZShift, private/clsim/I3CLSimModuleHelper.cxx:75-330 for the generators) from the oracle's plain parameter
// structs and return the text the reference would hand to its OpenCL compiler.  tests/test_ref_medium.py compiles
// that text for the host under the same OpenCL-C shim as the kernel text and holds the oracle's restatements against
// it value for value; tests/test_ref_program.py runs the reference's kernel on top of it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../clsim_oracle.h"

// ---- the reference's sources (unity build; order: bases first)
#include "clsim/function/I3CLSimFunction.cxx"
#include "clsim/function/I3CLSimFunctionAbsLenIceCube.cxx"
#include "clsim/function/I3CLSimFunctionConstant.cxx"
#include "clsim/function/I3CLSimFunctionDeltaPeak.cxx"
#include "clsim/function/I3CLSimFunctionFromTable.cxx"
#include "clsim/function/I3CLSimFunctionPolynomial.cxx"
#include "clsim/function/I3CLSimFunctionRefIndexIceCube.cxx"
#include "clsim/function/I3CLSimFunctionRefIndexQuanFry.cxx"
#include "clsim/function/I3CLSimFunctionScatLenIceCube.cxx"
#include "clsim/function/I3CLSimFunctionScatLenPartic.cxx"
#include "clsim/function/I3CLSimScalarField.cxx"
#include "clsim/function/I3CLSimScalarFieldAnisotropyAbsLenScaling.cxx"
#include "clsim/function/I3CLSimScalarFieldConstant.cxx"
#include "clsim/function/I3CLSimScalarFieldIceTiltZShift.cxx"
#include "clsim/function/I3CLSimVectorTransform.cxx"
#include "clsim/function/I3CLSimVectorTransformConstant.cxx"
#include "clsim/function/I3CLSimVectorTransformMatrix.cxx"
#include "clsim/random_value/I3CLSimRandomValue.cxx"
#include "clsim/random_value/I3CLSimRandomValueApplyFunction.cxx"
#include "clsim/random_value/I3CLSimRandomValueConstant.cxx"
#include "clsim/random_value/I3CLSimRandomValueFixParameter.cxx"
#include "clsim/random_value/I3CLSimRandomValueHenyeyGreenstein.cxx"
#include "clsim/random_value/I3CLSimRandomValueInterpolatedDistribution.cxx"
#include "clsim/random_value/I3CLSimRandomValueMixed.cxx"
#include "clsim/random_value/I3CLSimRandomValueNormalDistribution.cxx"
#include "clsim/random_value/I3CLSimRandomValueRayleighScatteringCosAngle.cxx"
#include "clsim/random_value/I3CLSimRandomValueSimplifiedLiu.cxx"
#include "clsim/random_value/I3CLSimRandomValueUniform.cxx"
#include "clsim/random_value/I3CLSimRandomValueWlenCherenkovNoDispersion.cxx"
// (two files of the unity build name a file-local serialization helper alike; serialization is not exercised here)
#define LoadFromArchiveIntoConstPtr LoadFromArchiveIntoConstPtr_of_medium_properties
#include "clsim/I3CLSimMediumProperties.cxx"
#undef LoadFromArchiveIntoConstPtr
#include "opencl/ieeehalfprecision.cxx"
#include "opencl/I3CLSimHelperGenerateMediumPropertiesSource_Optimizers.cxx"
#include "opencl/I3CLSimHelperGenerateMediumPropertiesSource.cxx"
// ... and the factories that MAKE the wavelength generators from a bias and a medium (private/clsim/I3CLSimModuleHelper.cxx:75-300;
// the converter / device classes its third function, initializeOpenCL, talks to are stand-ins: that function is not run)
#include "clsim/I3CLSimModuleHelper.cxx"
// ... and the table-maker's binning-code generators (they read the coordinate kernels from $I3_BUILD/clsim/resources/kernels/)
#include "clsim/tabulator/Axis.cxx"
#include "clsim/tabulator/Axes.cxx"

#include "ref_make_objects.h"

namespace {

thread_local std::string g_error;

int64_t give(const std::string &s, char *out, size_t cap)
{
    if (out && cap > 0) {
        const size_t n = std::min(s.size(), cap - 1);
        std::memcpy(out, s.data(), n);
        out[n] = 0;
    }
    return static_cast<int64_t>(s.size());
}

} // namespace

extern "C" {

const char *ref_medium_last_error() { return g_error.c_str(); }

// Text of I3CLSimHelper::GenerateMediumPropertiesSource for the medium `m` describes.  Returns the length of the text
// (call with out == NULL to size the buffer; tilt_z: the tilt table's z coordinates, tilt_num_z of them, or NULL), or -1 with ref_medium_last_error() set.
int64_t ref_medium_source(const oracle_medium *m, const double *tilt_z, char *out, size_t cap)
{
    try {
        I3CLSimMediumPropertiesPtr med = make_medium(*m, tilt_z);
        if (!med->IsReady()) throw std::runtime_error("medium is not ready");
        return give(I3CLSimHelper::GenerateMediumPropertiesSource(*med), out, cap);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Text of I3CLSimHelper::GenerateWavelengthGeneratorSource for n generators.
int64_t ref_wlen_generator_source(const oracle_wlen_generator *gens, int32_t n, char *out, size_t cap)
{
    try {
        std::vector<I3CLSimRandomValueConstPtr> v;
        for (int32_t i = 0; i < n; ++i) v.push_back(make_generator(gens[i]));
        return give(I3CLSimHelper::GenerateWavelengthGeneratorSource(v), out, cap);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Text of the wavelength bias function as I3CLSimStepToPhotonConverterOpenCL::GetWlenBiasSource asks for it
// (private/opencl/I3CLSimStepToPhotonConverterOpenCL.cxx:444-449: GetOpenCLFunction("getWavelengthBias")).
int64_t ref_wlen_bias_source(const oracle_wlen_bias *b, char *out, size_t cap)
{
    try {
        I3CLSimFunctionConstPtr f;
        if (b->kind == 0)
            f = I3CLSimFunctionConstPtr(new I3CLSimFunctionConstant(b->value));
        else
            f = I3CLSimFunctionConstPtr(new I3CLSimFunctionFromTable(b->x0, b->dx, std::vector<double>(b->v, b->v + b->n)));
        return give(f->GetOpenCLFunction("getWavelengthBias"), out, cap);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// The option #defines I3CLSimStepToPhotonConverterOpenCL::GetPreambleSource puts in front of the program
// (private/opencl/I3CLSimStepToPhotonConverterOpenCL.cxx:390-442, single precision; the six lines of
// I3CLSimHelper::GetMathPreamble, private/opencl/I3CLSimHelperMath.cxx:32-37, first).  That file needs an OpenCL
// runtime and cannot be compiled here, so the sequence is restated; the numbers are written by the reference's own
// ToFloatString (private/clsim/I3CLSimHelperToFloatString.h).  fixed_abs_lengths: NaN = option off.
int64_t ref_preamble_source(int32_t stop_detected_photons, int32_t save_all_photons, double save_all_prescale, int32_t history_entries,
                            double fixed_abs_lengths, double pancake_factor, char *out, size_t cap)
{
    using I3CLSimHelper::ToFloatString;
    std::string p = "typedef float floating_t;\n"
                    "typedef float2 floating2_t;\n"
                    "typedef float4 floating4_t;\n"
                    "#define convert_floating_t convert_float\n"
                    "#define ZERO 0.f\n"
                    "#define ONE 1.f\n"
                    "\n";
    if (stop_detected_photons) p += "#define STOP_PHOTONS_ON_DETECTION\n";
    if (save_all_photons) {
        p += "#define SAVE_ALL_PHOTONS\n";
        p += "#define SAVE_ALL_PHOTONS_PRESCALE " + ToFloatString(save_all_prescale) + "\n";
    }
    if (history_entries > 0) {
        p += "#define SAVE_PHOTON_HISTORY\n";
        p += "#define NUM_PHOTONS_IN_HISTORY " + std::to_string(history_entries) + "\n";
    }
    if (!std::isnan(fixed_abs_lengths)) p += "#define PROPAGATE_FOR_FIXED_NUMBER_OF_ABSORPTION_LENGTHS " + ToFloatString(fixed_abs_lengths) + "\n";
    if (pancake_factor != 1.) p += "#define PANCAKE_FACTOR " + ToFloatString(pancake_factor) + "\n";
    return give(p, out, cap);
}

// Text of the wavelength generator the reference's own factory makes for (bias, medium): makeCherenkovWavelengthGenerator
// (I3CLSimModuleHelper.cxx:176-300) when n_spectrum == 0, else makeWavelengthGenerator (:75-173) for the tabulated spectrum
// (spectrum_x, spectrum_y), e.g. a flasher LED.  The product's clsim_b200/ice.py restates both factories.
int64_t ref_made_wlen_generator_source(const oracle_medium *m, const double *tilt_z, const oracle_wlen_bias *b, int32_t without_dispersion,
                                       const double *spectrum_x, const double *spectrum_y, int32_t n_spectrum, char *out, size_t cap)
{
    try {
        I3CLSimMediumPropertiesPtr med = make_medium(*m, tilt_z);
        I3CLSimFunctionConstPtr bias;
        if (b->kind == 0) bias = I3CLSimFunctionConstPtr(new I3CLSimFunctionConstant(b->value));
        else bias = I3CLSimFunctionConstPtr(new I3CLSimFunctionFromTable(b->x0, b->dx, std::vector<double>(b->v, b->v + b->n)));
        I3CLSimRandomValueConstPtr gen;
        if (n_spectrum == 0) {
            gen = I3CLSimModuleHelper::makeCherenkovWavelengthGenerator(bias, without_dispersion != 0, med);
        } else {
            I3CLSimFunctionConstPtr spectrum(new I3CLSimFunctionFromTable(std::vector<double>(spectrum_x, spectrum_x + n_spectrum),
                                                                          std::vector<double>(spectrum_y, spectrum_y + n_spectrum)));
            gen = I3CLSimModuleHelper::makeWavelengthGenerator(spectrum, bias, med);
        }
        return give(I3CLSimHelper::GenerateWavelengthGeneratorSource(std::vector<I3CLSimRandomValueConstPtr>(1, gen)), out, cap);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// ---- table-maker (private/clsim/tabulator/): the parts of the program I3CLSimStepToTableConverter joins that differ
//      from the step-to-photon program

// Axes::GenerateBinningCode for the axes the oracle's table description names (kind 0 linear, 1 power).  The generator
// loads {spherical,cylindrical}_coordinates.c.cl from $I3_BUILD/clsim/resources/kernels/: the caller sets I3_BUILD to a
// directory in which `clsim` is a link to the reference tree.  n_bins_out: Axes::GetNBins() (with under/overflow).
int64_t ref_binning_source(int32_t geometry, int32_t num_axes, const int32_t *kind, const uint32_t *power, const uint32_t *bins, const double *lo,
                           const double *hi, uint64_t *n_bins_out, char *out, size_t cap)
{
    using namespace clsim::tabulator;
    try {
        std::vector<Axes::value_type> ax;
        for (int32_t i = 0; i < num_axes; ++i) {
            if (kind[i] == 0) ax.push_back(Axes::value_type(new LinearAxis(lo[i], hi[i], bins[i])));
            else ax.push_back(Axes::value_type(new PowerAxis(lo[i], hi[i], bins[i], power[i])));
        }
        std::string text;
        if (geometry == 0) {
            SphericalAxes axes(ax);
            if (n_bins_out) *n_bins_out = axes.GetNBins();
            text = axes.GenerateBinningCode();
        } else {
            CylindricalAxes axes(ax);
            if (n_bins_out) *n_bins_out = axes.GetNBins();
            text = axes.GenerateBinningCode();
        }
        return give(text, out, cap);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Axes::GetBinVolume (tabulator/Axes.cxx:121-166; what I3CLSimStepToTableConverter::Normalize divides the bin content by) for
// n index triples (the three spatial axes), and Axis::GetBinEdges of one axis: edges_out has bins[axis] + 1 entries.
int32_t ref_bin_volumes(int32_t geometry, int32_t num_axes, const int32_t *kind, const uint32_t *power, const uint32_t *bins, const double *lo,
                        const double *hi, const uint64_t *idx3, uint64_t n, double *volumes_out, int32_t edges_axis, double *edges_out)
{
    using namespace clsim::tabulator;
    try {
        std::vector<Axes::value_type> ax;
        for (int32_t i = 0; i < num_axes; ++i) {
            if (kind[i] == 0) ax.push_back(Axes::value_type(new LinearAxis(lo[i], hi[i], bins[i])));
            else ax.push_back(Axes::value_type(new PowerAxis(lo[i], hi[i], bins[i], power[i])));
        }
        boost::shared_ptr<Axes> axes;
        if (geometry == 0) axes.reset(new SphericalAxes(ax));
        else axes.reset(new CylindricalAxes(ax));
        for (uint64_t k = 0; k < n; ++k) {
            std::vector<size_t> idxs(idx3 + 3 * k, idx3 + 3 * k + 3);
            volumes_out[k] = axes->GetBinVolume(idxs);
        }
        if (edges_out && edges_axis >= 0 && edges_axis < num_axes) {
            const std::vector<double> e = axes->at(edges_axis)->GetBinEdges();
            for (size_t i = 0; i < e.size(); ++i) edges_out[i] = e[i];
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// angularAcceptance->GetOpenCLFunction("getAngularAcceptance") for the polynomial python/GetIceCubeDOMAngularSensitivity.py:45 returns
int64_t ref_angular_acceptance_source(const double *coefficients, int32_t n, char *out, size_t cap)
{
    try {
        I3CLSimFunctionPolynomial f(std::vector<double>(coefficients, coefficients + n));
        return give(f.GetOpenCLFunction("getAngularAcceptance"), out, cap);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// GetMinimumRefractiveIndex (private/clsim/tabulator/I3CLSimStepToTableConverter.cxx:95-119; that file needs an OpenCL
// runtime, so its fifteen lines are restated on the reference's own classes -- including the scan `wmin + i*(wmax-wmin)`
// that leaves the wavelength range after its first point).  out[0] = group index, out[1] = phase index.
int32_t ref_minimum_refractive_index(const oracle_medium *m, double *out)
{
    try {
        I3CLSimMediumPropertiesPtr medp = make_medium(*m, nullptr);
        const I3CLSimMediumProperties &med = *medp;
        std::pair<double, double> n_min(std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity());
        for (unsigned j = 0; j < med.GetLayersNum(); j++) {
            const I3CLSimFunction &groupIndex = *(med.GetGroupRefractiveIndexOverride(j));
            const I3CLSimFunction &phaseIndex = *(med.GetPhaseRefractiveIndex(j));
            double wmin = std::max(med.GetMinWavelength(), groupIndex.GetMinWlen());
            double wmax = std::min(med.GetMaxWavelength(), groupIndex.GetMaxWlen());
            for (unsigned i = 0; i < 1000; i++) {
                double n = groupIndex.GetValue(wmin + i * (wmax - wmin));
                if (n > 1 && n < n_min.first) n_min = std::make_pair(n, phaseIndex.GetValue(wmin + i * (wmax - wmin)));
            }
        }
        out[0] = n_min.first;
        out[1] = n_min.second;
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// The table-maker's preamble (…StepToTableConverter.cxx:178-199, after the six lines of GetMathPreamble).
int64_t ref_table_preamble_source(int32_t num_axes, uint64_t entries_per_stream, double step_length, double n_group, double n_phase, char *out, size_t cap)
{
    std::ostringstream preamble;
    preamble << "typedef float floating_t;\n"
                "typedef float2 floating2_t;\n"
                "typedef float4 floating4_t;\n"
                "#define convert_floating_t convert_float\n"
                "#define ZERO 0.f\n"
                "#define ONE 1.f\n"
                "\n";
    preamble << "#define SAVE_ALL_PHOTONS\n"
                "#define SAVE_ALL_PHOTONS_PRESCALE 1\n"
                "#define PROPAGATE_FOR_FIXED_NUMBER_OF_ABSORPTION_LENGTHS 42\n"
                "#define TABULATE\n"
                "//#define DOM_RADIUS " << I3CLSimHelper::ToFloatString(0.16510 * I3Units::m) << "\n"
                "//#define PRINTF_ENABLED\n";
    if (num_axes > 4) preamble << "#define TABULATE_IMPACT_ANGLE\n";
    preamble << "#define TABLE_ENTRIES_PER_STREAM " << entries_per_stream << "\n";
    preamble << "#define VOLUME_MODE_STEP " << I3CLSimHelper::ToFloatString(step_length) << "\n";
    preamble << "__constant floating_t min_invGroupVel = " << I3CLSimHelper::ToFloatString(n_group / I3Constants::c) << ";\n";
    preamble << "__constant floating_t tan_thetaC = " << I3CLSimHelper::ToFloatString(std::sqrt(n_phase * n_phase - 1.)) << ";\n";
    return give(preamble.str(), out, cap);
}

// The host-side twins the reference's classes carry (GetValue / ApplyTransform in double): evaluated for completeness
// of the pin -- the device text and the host method of one class are meant to agree.
double ref_medium_host_value(const oracle_medium *m, const double *tilt_z, int32_t what, int32_t layer, double a, double b, double c)
{
    try {
        I3CLSimMediumPropertiesPtr med = make_medium(*m, tilt_z);
        switch (what) {
        case 0: return med->GetPhaseRefractiveIndex(layer)->GetValue(a);
        case 1: return med->GetGroupRefractiveIndexOverride(layer)->GetValue(a);
        case 2: return med->GetScatteringLength(layer)->GetValue(a);
        case 3: return med->GetAbsorptionLength(layer)->GetValue(a);
        case 4: return med->GetIceTiltZShift()->GetValue(a, b, c);
        case 5: return med->GetDirectionalAbsorptionLengthCorrection()->GetValue(a, b, c);
        case 6: return med->GetMinWavelength();
        case 7: return med->GetMaxWavelength();
        }
        throw std::runtime_error("unknown quantity");
    } catch (const std::exception &e) {
        g_error = e.what();
        return NAN;
    }
}

// ... the same for n argument triples at once (one medium object): abc[3 * i ..], out[i]
int32_t ref_medium_host_values(const oracle_medium *m, const double *tilt_z, int32_t what, int32_t layer, const double *abc, double *out, uint64_t n)
{
    try {
        I3CLSimMediumPropertiesPtr med = make_medium(*m, tilt_z);
        for (uint64_t i = 0; i < n; ++i) {
            const double a = abc[3 * i], b = abc[3 * i + 1], c = abc[3 * i + 2];
            switch (what) {
            case 0: out[i] = med->GetPhaseRefractiveIndex(layer)->GetValue(a); break;
            case 1: out[i] = med->GetGroupRefractiveIndexOverride(layer)->GetValue(a); break;
            case 2: out[i] = med->GetScatteringLength(layer)->GetValue(a); break;
            case 3: out[i] = med->GetAbsorptionLength(layer)->GetValue(a); break;
            case 4: out[i] = med->GetIceTiltZShift()->GetValue(a, b, c); break;
            case 5: out[i] = med->GetDirectionalAbsorptionLengthCorrection()->GetValue(a, b, c); break;
            case 8: case 9: {
                // ApplyTransform of the pre (8) / post (9) scattering direction transform: out holds 3 values per triple
                const std::vector<double> v = (what == 8 ? med->GetPreScatterDirectionTransform() : med->GetPostScatterDirectionTransform())
                                                  ->ApplyTransform(std::vector<double>{a, b, c});
                out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
                break;
            }
            default: throw std::runtime_error("unknown quantity");
            }
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

} // extern "C"
