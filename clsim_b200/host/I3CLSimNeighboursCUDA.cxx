// I3CLSimNeighboursCUDA.cxx -- see the header.  Flattening of the reference's description objects to the POD
// blocks of include/clsimcuda.h; anything that is not the class the reference's tray segments build
// (python/traysegments/common.py:181-215) throws, naming the class.
#include "I3CLSimNeighboursCUDA.h"

#include <cmath>
#include <cstring>
#include <string>
#include <typeinfo>

typedef I3CLSimStepToPhotonConverter_exception Err;

namespace {
void throw_capi(const char *what) { throw Err(std::string(what) + ": " + clsimcu_last_error()); }
} // namespace

I3CLSimPhotonToMCPEConverterCUDA::I3CLSimPhotonToMCPEConverterCUDA(uint64_t randomSeed, const std::map<OMKey, I3CLSimFunctionConstPtr> &wavelengthAcceptance,
                                                                   I3CLSimFunctionConstPtr angularAcceptance, int device, uint64_t firstRNGMultiplierRow)
    : handle_(nullptr)
{
    if (wavelengthAcceptance.empty()) throw Err("The \"WavelengthAcceptance\" parameter must not be empty.");
    if (!angularAcceptance) throw Err("The \"AngularAcceptance\" parameter must not be empty.");
    const I3CLSimFunctionPolynomial *poly = dynamic_cast<const I3CLSimFunctionPolynomial *>(angularAcceptance.get());
    if (!poly) throw Err(std::string("angular acceptance must be an I3CLSimFunctionPolynomial, got ") + typeid(*angularAcceptance).name());
    // distinct acceptance functions -> tables; DOMs -> index of their function
    std::vector<const I3CLSimFunction *> distinct;
    std::vector<std::vector<double> > values;
    std::vector<clsimcu_wlen_bias> tables;
    std::vector<int32_t> strings;
    std::vector<uint32_t> oms;
    std::vector<uint8_t> which;
    for (std::map<OMKey, I3CLSimFunctionConstPtr>::const_iterator it = wavelengthAcceptance.begin(); it != wavelengthAcceptance.end(); ++it) {
        if (!it->second) throw Err("No wavelength acceptance configured for an OMKey in the map");
        std::size_t k = 0;
        while (k < distinct.size() && distinct[k] != it->second.get()) ++k;
        if (k == distinct.size()) {
            clsimcu_wlen_bias t;
            std::memset(&t, 0, sizeof t);
            if (const I3CLSimFunctionFromTable *ft = dynamic_cast<const I3CLSimFunctionFromTable *>(it->second.get())) {
                if (!ft->GetInEqualSpacingMode()) throw Err("wavelength acceptance tables must be in equal spacing mode");
                std::vector<double> v(ft->GetNumEntries());
                for (std::size_t i = 0; i < v.size(); ++i) v[i] = ft->GetEntryValue(i);
                values.push_back(v);
                t.kind = CLSIMCU_BIAS_TABLE;
                t.n = static_cast<int32_t>(v.size());
                t.x0 = ft->GetFirstWavelength();
                t.dx = ft->GetWavelengthStepping();
            } else if (dynamic_cast<const I3CLSimFunctionConstant *>(it->second.get())) {
                values.push_back(std::vector<double>());
                t.kind = CLSIMCU_BIAS_CONSTANT;
                t.value = it->second->GetValue(400e-9);
            } else {
                throw Err(std::string("wavelength acceptance must be I3CLSimFunctionFromTable or I3CLSimFunctionConstant, got ") + typeid(*it->second).name());
            }
            distinct.push_back(it->second.get());
            tables.push_back(t);
            if (distinct.size() > 255) throw Err("more than 255 distinct wavelength acceptance functions");
        }
        strings.push_back(it->first.GetString());
        oms.push_back(it->first.GetOM());
        which.push_back(static_cast<uint8_t>(k));
    }
    for (std::size_t k = 0; k < tables.size(); ++k)
        if (tables[k].kind == CLSIMCU_BIAS_TABLE) tables[k].v = values[k].data();
    clsimcu_mcpe_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = static_cast<int32_t>(sizeof cfg);
    cfg.device = device;
    cfg.flavour = CLSIMCU_MCPE_INLOOP;
    cfg.num_acceptances = static_cast<int32_t>(tables.size());
    cfg.acceptances = tables.data();
    cfg.num_doms = static_cast<int32_t>(strings.size());
    cfg.string_id = strings.data();
    cfg.dom_id = oms.data();
    cfg.acceptance_of_dom = which.data();
    cfg.num_angular_coefficients = static_cast<int32_t>(poly->GetCoefficients().size());
    cfg.angular_coefficients = poly->GetCoefficients().data();
    cfg.dom_dir[2] = -1.;
    cfg.dom_radius = 0.16510;
    cfg.oversize_factor = cfg.pancake_factor = 1.;
    cfg.rng_seed = randomSeed;
    cfg.rng_first_multiplier = firstRNGMultiplierRow;
    if (clsimcu_mcpe_create(&cfg, &handle_) != CLSIMCU_OK) throw_capi("I3CLSimPhotonToMCPEConverterCUDA");
}

I3CLSimPhotonToMCPEConverterCUDA::~I3CLSimPhotonToMCPEConverterCUDA()
{
    if (handle_) clsimcu_mcpe_destroy(handle_);
}

std::vector<clsimcu_mcpe> I3CLSimPhotonToMCPEConverterCUDA::Convert(const I3CLSimPhotonSeries &photons, const std::vector<float> *uniforms)
{
    if (uniforms && uniforms->size() != photons.size()) throw Err("one uniform per photon");
    std::vector<clsimcu_mcpe> out(photons.size());
    std::size_t n = 0;
    static_assert(sizeof(I3CLSimPhoton) == sizeof(clsimcu_photon), "photon record layout");
    if (clsimcu_mcpe_convert(handle_, reinterpret_cast<const clsimcu_photon *>(photons.data()), photons.size(), uniforms ? uniforms->data() : nullptr,
                             out.data(), out.size(), &n) != CLSIMCU_OK)
        throw_capi("Convert");
    out.resize(n);
    return out;
}

void I3CLSimPhotonToMCPEConverterCUDA::AttachTo(I3CLSimStepToPhotonConverterCUDA &converter, bool keepPhotons)
{
    if (!converter.IsInitialized()) throw Err("I3CLSimStepToPhotonConverterCUDA is not initialized!");
    if (clsimcu_attach_mcpe_converter(converter.GetEngine(), handle_, keepPhotons ? 1 : 0) != CLSIMCU_OK) throw_capi("AttachTo");
}

// ---- steps ---------------------------------------------------------------------------------------------
I3CLSimStepGeneratorCUDA::I3CLSimStepGeneratorCUDA(uint64_t randomSeed, int device, uint64_t firstRNGMultiplierRow, double angularDistA, double angularDistB)
    : handle_(nullptr)
{
    clsimcu_step_generator_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = static_cast<int32_t>(sizeof cfg);
    cfg.device = device;
    cfg.angular_a = angularDistA;
    cfg.angular_b = angularDistB;
    cfg.rng_seed = randomSeed;
    cfg.rng_first_multiplier = firstRNGMultiplierRow;
    if (clsimcu_stepgen_create(&cfg, &handle_) != CLSIMCU_OK) throw_capi("I3CLSimStepGeneratorCUDA");
}

I3CLSimStepGeneratorCUDA::~I3CLSimStepGeneratorCUDA()
{
    if (handle_) clsimcu_stepgen_destroy(handle_);
}

std::vector<clsimcu_step_source> I3CLSimStepGeneratorCUDA::Flatten(const std::vector<Source> &sources)
{
    std::vector<clsimcu_step_source> out(sources.size());
    for (std::size_t i = 0; i < sources.size(); ++i) {
        const Source &s = sources[i];
        clsimcu_step_source &o = out[i];
        std::memset(&o, 0, sizeof o);
        o.x = s.x; o.y = s.y; o.z = s.z; o.t = s.time;
        o.dir_x = s.dirX; o.dir_y = s.dirY; o.dir_z = s.dirZ;
        o.identifier = s.particleIdentifier;
        if (s.photonsPerStep > 0xffffffffull || s.numPhotonsInLastStep > 0xffffffffull) throw Err("photons per step do not fit 32 bits");
        o.photons_per_step = static_cast<uint32_t>(s.photonsPerStep);
        o.photons_in_last_step = static_cast<uint32_t>(s.numPhotonsInLastStep);
        o.num_steps = s.numSteps;
        if (s.isCascade) {
            o.kind = CLSIMCU_SOURCE_CASCADE;
            o.pa = s.pa;
            o.pb = s.pb;
        } else {
            o.kind = s.stepIsCascadeLike ? CLSIMCU_SOURCE_TRACK_CASCADE_LIKE : CLSIMCU_SOURCE_TRACK_MUON_LIKE;
            o.length = s.length;
        }
    }
    return out;
}

I3CLSimStepSeriesPtr I3CLSimStepGeneratorCUDA::MakeSteps(const std::vector<Source> &sources)
{
    const std::vector<clsimcu_step_source> flat = Flatten(sources);
    std::size_t total = 0;
    for (std::size_t i = 0; i < flat.size(); ++i) total += flat[i].num_steps + (flat[i].photons_in_last_step > 0 ? 1 : 0);
    I3CLSimStepSeriesPtr steps(new I3CLSimStepSeries(total));
    static_assert(sizeof(I3CLSimStep) == sizeof(clsimcu_step), "step record layout");
    std::size_t n = 0;
    if (clsimcu_stepgen_generate(handle_, flat.data(), flat.size(), reinterpret_cast<clsimcu_step *>(steps->data()), total, &n) != CLSIMCU_OK)
        throw_capi("MakeSteps");
    return steps;
}

std::size_t I3CLSimStepGeneratorCUDA::EnqueueInto(I3CLSimStepToPhotonConverterCUDA &converter, const std::vector<Source> &sources, uint32_t identifier)
{
    if (!converter.IsInitialized()) throw Err("I3CLSimStepToPhotonConverterCUDA is not initialized!");
    const std::vector<clsimcu_step_source> flat = Flatten(sources);
    std::size_t total = 0;
    for (std::size_t i = 0; i < flat.size(); ++i) total += flat[i].num_steps + (flat[i].photons_in_last_step > 0 ? 1 : 0);
    if (clsimcu_enqueue_sources(converter.GetEngine(), handle_, flat.data(), flat.size(), identifier) != CLSIMCU_OK) throw_capi("EnqueueSteps");
    return total;
}
