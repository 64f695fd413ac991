// ref_icetray_mode.cpp -- the PRODUCT's converter class as a maintainer of the reference would build it: compiled with
// -DCLSIM_CUDA_IN_ICETRAY against the reference's own public headers (public/clsim/I3CLSimStepToPhotonConverter.h, the
// description classes, the step / photon records), patched with exactly the getters INTEGRATION.md section 2 lists
// (tools/integration_getters.py), and linked with the reference's own description-class sources and with libclsimcuda.so.
// TEST INFRASTRUCTURE (oracle/_ref/libclsim_icetray_mode.so): proves the drop-in boundary on the CPU -- the class derives from
// the reference's abstract interface, takes the reference's objects, and flattens them to the tables the Python path makes.
// IceTray itself (logging, pointer typedefs, serialization, I3Vector ...) is the stand-ins under oracle/ref_shim/host*/.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "I3CLSimStepToPhotonConverterCUDA.h"   // clsim_b200/host, IceTray mode
#include "clsim/I3CLSimSimpleGeometryUserConfigurable.h"
#include "clsimcuda.h"
#include "ref_make_objects.h"

namespace {
thread_local std::string g_error, g_text;
}

extern "C" {

const char *icetray_mode_error() { return g_error.c_str(); }

// SetWlenGenerators / SetWlenBias / SetMediumProperties / SetGeometry / the option setters / Compile() on an
// I3CLSimStepToPhotonConverterCUDA -- through a pointer to the REFERENCE's abstract base where the base has the method --
// then DescribeTables(): the JSON text of the tables the library would put on the device.  Returns the text's length (-1: error);
// *out points to it (valid until the next call on this thread).
int64_t icetray_mode_describe_tables(const oracle_config *cfg, const double *tilt_z, const char *const *subdetector_names, const char **out)
{
    try {
        boost::shared_ptr<I3CLSimStepToPhotonConverterCUDA> cuda(new I3CLSimStepToPhotonConverterCUDA(cfg->rng_seed, cfg->kernel_mode == 0));
        I3CLSimStepToPhotonConverter *base = cuda.get();          // the interface I3CLSimModule / I3CLSimServer see

        cuda->SetDevice(cfg->device);
        std::vector<I3CLSimRandomValueConstPtr> gens;
        for (int32_t i = 0; i < cfg->num_wlen_generators; ++i) gens.push_back(make_generator(cfg->wlen_generators[i]));
        base->SetWlenGenerators(gens);
        base->SetWlenBias(make_bias(cfg->wlen_bias));
        base->SetMediumProperties(make_medium(cfg->medium, tilt_z));
        const oracle_geometry &g = cfg->geometry;
        boost::shared_ptr<I3CLSimSimpleGeometryUserConfigurable> geo(new I3CLSimSimpleGeometryUserConfigurable(g.om_radius, g.num_doms));
        for (int32_t i = 0; i < g.num_doms; ++i) {
            geo->SetStringID(i, g.string_id[i]);
            geo->SetDomID(i, g.dom_id[i]);
            geo->SetPosX(i, g.x[i]);
            geo->SetPosY(i, g.y[i]);
            geo->SetPosZ(i, g.z[i]);
            geo->SetSubdetector(i, subdetector_names[g.subdetector[i]]);
        }
        base->SetGeometry(geo);

        cuda->SetStopDetectedPhotons(cfg->stop_detected_photons != 0);
        cuda->SetSaveAllPhotons(cfg->save_all_photons != 0);
        cuda->SetSaveAllPhotonsPrescale(cfg->save_all_photons_prescale);
        cuda->SetFixedNumberOfAbsorptionLengths(cfg->fixed_number_of_absorption_lengths);
        cuda->SetDOMPancakeFactor(cfg->pancake_factor);
        cuda->SetPhotonHistoryEntries(cfg->photon_history_entries);
        cuda->Compile();
        g_text = cuda->DescribeTables();
        *out = g_text.c_str();
        return static_cast<int64_t>(g_text.size());
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// What the class says to a description object it does not know (an ANTARES scattering model): the reference's exception type.
int32_t icetray_mode_unknown_class_is_refused(const oracle_config *cfg)
{
    try {
        boost::shared_ptr<I3CLSimStepToPhotonConverterCUDA> cuda(new I3CLSimStepToPhotonConverterCUDA(1, true));
        cuda->SetDevice(0);
        std::vector<I3CLSimRandomValueConstPtr> gens(1, make_generator(cfg->wlen_generators[0]));
        cuda->SetWlenGenerators(gens);
        cuda->SetWlenBias(make_bias(cfg->wlen_bias));
        I3CLSimMediumPropertiesPtr med = make_medium(cfg->medium, nullptr);
        med->SetScatteringCosAngleDistribution(I3CLSimRandomValueConstPtr(new I3CLSimRandomValueConstant(0.9)));
        cuda->SetMediumProperties(med);
        cuda->SetGeometry(I3CLSimSimpleGeometryConstPtr(new I3CLSimSimpleGeometryUserConfigurable(0.2, 0)));
        cuda->Compile();
        return 0;
    } catch (const I3CLSimStepToPhotonConverter_exception &e) {
        g_error = e.what();
        return 1;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

} // extern "C"
