// ref_mcpe.cpp -- the reference's photon -> photo-electron converters, compiled for the host.
// TEST INFRASTRUCTURE (oracle/_ref/libclsim_ref_mcpe.so).  Nothing under clsim_b200/ links this.
//
// private/clsim/dom/I3PhotonToMCPEConverter.cxx -- BOTH converters of row f3: I3CLSimPhotonToMCPEConverterForDOMs::Convert
// (the one clsim runs as photons come back from the device) and the I3PhotonToMCPEConverter module (oversize / pancake time
// correction, relative DOM efficiency from the calibration) -- is compiled unmodified, together with the acceptance
// function classes it evaluates (I3CLSimFunctionFromTable / Polynomial / Constant).  IceTray's module protocol and data
// classes are the stand-ins under oracle/ref_shim/host_mcpe/: plain data holders, plus an I3Module whose parameters the
// driver sets by name and an I3Frame that is a map of named objects.  The random service hands out the uniforms the caller
// supplies, in the order the reference asks for them.  Hit merging (sim-services, un-vendored) is not pinned.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "clsim/dom/I3PhotonToMCPEConverter.h"
#include "clsim/function/I3CLSimFunctionConstant.h"
#include "clsim/function/I3CLSimFunctionFromTable.h"
#include "clsim/function/I3CLSimFunctionPolynomial.h"
#include "simclasses/I3Photon.h"

namespace {

thread_local std::string g_error;

class SuppliedUniforms : public I3RandomService {
public:
    SuppliedUniforms(const double *u, size_t n) : u_(u), n_(n), at_(0) {}
    unsigned int Integer(unsigned int) override { throw std::runtime_error("not used"); }
    double Uniform(double) override
    {
        if (at_ >= n_) throw std::runtime_error("out of supplied uniforms");
        return u_[at_++];
    }
    size_t used() const { return at_; }
private:
    const double *u_;
    size_t n_, at_;
};

// the 80-byte record of the path (include/clsimcuda.h: clsimcu_photon == I3CLSimPhoton)
struct Record {
    float x, y, z, t, theta, phi, wavelength, cherenkov_dist;
    uint32_t num_scatters;
    float weight;
    uint32_t identifier;
    int16_t string_id;
    uint16_t om_id;
    float start_x, start_y, start_z, start_t, start_theta, start_phi, group_velocity, dist_in_abs_lens;
};
static_assert(sizeof(Record) == 80, "record layout");

I3CLSimFunctionConstPtr acceptance(const double *values, int32_t n, double x0, double dx, double constant)
{
    if (n > 0) return I3CLSimFunctionConstPtr(new I3CLSimFunctionFromTable(x0, dx, std::vector<double>(values, values + n)));
    return I3CLSimFunctionConstPtr(new I3CLSimFunctionConstant(constant));
}

// Emit<I3CompressedPhoton> of private/clsim/I3CLSimClientModule.cxx:326-349 (time shift 0)
template <class P> void fill(P &p, const Record &r)
{
    p.SetTime(r.t);
    p.SetWeight(r.weight);
    p.SetParticleID(I3ParticleID(r.identifier, 0));
    p.SetWavelength(r.wavelength);
    p.SetGroupVelocity(r.group_velocity);
    p.SetPos(I3Position(r.x, r.y, r.z));
    I3Direction d;
    d.SetThetaPhi(r.theta, r.phi);
    p.SetDir(d);
}

} // namespace

extern "C" {

const char *ref_mcpe_error() { return g_error.c_str(); }

// I3CLSimPhotonToMCPEConverterForDOMs::Convert for each of n photons (positions relative to their DOM), photon i drawing
// uniforms[i] if it gets that far.  survive[i], time[i] out.  Returns the number of survivors, -1 on log_fatal (message kept).
int64_t ref_mcpe_convert_inloop(const void *photons, uint64_t n, const double *acc_values, int32_t acc_n, double acc_x0, double acc_dx,
                                double acc_constant, const double *angular, int32_t n_angular, const double *uniforms, uint8_t *survive,
                                double *time_out, uint64_t *uniforms_used)
{
    try {
        const Record *rec = static_cast<const Record *>(photons);
        boost::shared_ptr<std::map<OMKey, I3CLSimFunctionConstPtr> > acc(new std::map<OMKey, I3CLSimFunctionConstPtr>());
        I3CLSimFunctionConstPtr f = acceptance(acc_values, acc_n, acc_x0, acc_dx, acc_constant);
        for (uint64_t i = 0; i < n; ++i) (*acc)[OMKey(rec[i].string_id, rec[i].om_id)] = f;
        I3CLSimFunctionConstPtr ang(new I3CLSimFunctionPolynomial(std::vector<double>(angular, angular + n_angular)));
        int64_t count = 0;
        uint64_t used = 0;
        for (uint64_t i = 0; i < n; ++i) {
            boost::shared_ptr<SuppliedUniforms> rng(new SuppliedUniforms(uniforms + i, 1));
            I3CLSimPhotonToMCPEConverterForDOMs conv(rng, acc, ang);
            I3CompressedPhoton p;
            fill(p, rec[i]);
            const std::tuple<OMKey, I3MCPE, bool> out = conv.Convert(ModuleKey(rec[i].string_id, rec[i].om_id), p);
            survive[i] = std::get<2>(out) ? 1 : 0;
            time_out[i] = std::get<2>(out) ? std::get<1>(out).time : 0.;
            if (std::get<2>(out)) {
                ++count;
                if (!(std::get<0>(out) == OMKey(rec[i].string_id, rec[i].om_id)) || std::get<1>(out).npe != 1 ||
                    std::get<1>(out).ID.majorID != rec[i].identifier)
                    throw std::runtime_error("unexpected photo-electron record");
            }
            used += rng->used();
        }
        if (uniforms_used) *uniforms_used = used;
        return count;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// The I3PhotonToMCPEConverter MODULE on one frame: n photons (positions relative to their DOM; the DOM sits at dom_xyz[3*i..],
// the frame carries absolute positions), one wavelength acceptance, a relative efficiency per photon's DOM through an
// I3Calibration (NaN = no entry: the module's default applies).  Uniforms are handed out in the order the module asks:
// DOMs in key order, photons of a DOM in input order, photons of weight 0 skipped.  Output: per photo-electron
// (string, om, time, index of the input photon) in the module's output order (per DOM sorted by time).
// Returns the number of photo-electrons (at most cap are written), -1 on log_fatal.
int64_t ref_mcpe_convert_module(const void *photons, uint64_t n, const double *dom_xyz, const double *acc_values, int32_t acc_n, double acc_x0,
                                double acc_dx, double acc_constant, const double *angular, int32_t n_angular, const double *efficiency,
                                double default_efficiency, int32_t replace_with_default, double oversize, double pancake, double dom_radius,
                                int32_t only_warn, const double *uniforms, uint64_t n_uniforms, int32_t *out_string, uint32_t *out_om,
                                double *out_time, int64_t *out_photon, uint64_t cap, uint64_t *uniforms_used)
{
    try {
        const Record *rec = static_cast<const Record *>(photons);
        I3Context context;
        I3PhotonToMCPEConverter module(context);
        boost::shared_ptr<SuppliedUniforms> rng(new SuppliedUniforms(uniforms, n_uniforms));
        module.Set<I3RandomServicePtr>("RandomService", rng);
        module.Set<I3CLSimFunctionConstPtr>("WavelengthAcceptance", acceptance(acc_values, acc_n, acc_x0, acc_dx, acc_constant));
        module.Set<I3CLSimFunctionConstPtr>("AngularAcceptance",
                                            I3CLSimFunctionConstPtr(new I3CLSimFunctionPolynomial(std::vector<double>(angular, angular + n_angular))));
        module.Set<double>("DOMOversizeFactor", oversize);
        module.Set<double>("DOMPancakeFactor", pancake);
        module.Set<double>("DOMRadiusWithoutOversize", dom_radius);
        module.Set<double>("DefaultRelativeDOMEfficiency", default_efficiency);
        module.Set<bool>("ReplaceRelativeDOMEfficiencyWithDefault", replace_with_default != 0);
        module.Set<bool>("OnlyWarnAboutInvalidPhotonPositions", only_warn != 0);
        module.Configure();

        boost::shared_ptr<I3Calibration> calibration(new I3Calibration());
        boost::shared_ptr<I3OMGeoMap> omgeo(new I3OMGeoMap());
        boost::shared_ptr<I3ModuleGeoMap> modulegeo(new I3ModuleGeoMap());
        boost::shared_ptr<I3PhotonSeriesMap> series(new I3PhotonSeriesMap());
        std::map<ModuleKey, std::vector<int64_t> > source;   // which input photon each series entry is
        const I3Direction down(0., 0., -1.);                 // IceCube PMTs look down
        for (uint64_t i = 0; i < n; ++i) {
            const OMKey key(rec[i].string_id, rec[i].om_id);
            const ModuleKey mkey(rec[i].string_id, rec[i].om_id);
            const I3Position at(dom_xyz[3 * i], dom_xyz[3 * i + 1], dom_xyz[3 * i + 2]);
            I3OMGeo g;
            g.position = at;
            g.direction = down;
            (*omgeo)[key] = g;
            (*modulegeo)[mkey] = I3ModuleGeo(at, down, dom_radius);
            if (efficiency[i] == efficiency[i]) {
                I3DOMCalibration c;
                c.SetRelativeDomEff(efficiency[i]);
                calibration->domCal[key] = c;
            }
            I3Photon p;
            fill(p, rec[i]);
            p.SetPos(I3Position(at.GetX() + rec[i].x, at.GetY() + rec[i].y, at.GetZ() + rec[i].z));
            p.SetStartPos(I3Position(at.GetX() + rec[i].start_x, at.GetY() + rec[i].start_y, at.GetZ() + rec[i].start_z));
            p.SetNumScattered(rec[i].num_scatters);
            p.SetParticleID(I3ParticleID(i, 0));   // (the photon's index rides along as the particle ID)
            (*series)[mkey].push_back(p);
        }
        I3FramePtr cal(new I3Frame());
        cal->Put("I3Calibration", calibration);
        module.Calibration(cal);
        I3FramePtr frame(new I3Frame());
        frame->Put("I3OMGeoMap", omgeo);
        frame->Put("I3ModuleGeoMap", modulegeo);
        frame->Put("PropagatedPhotons", series);
        module.DAQ(frame);
        I3MCPESeriesMapConstPtr out = frame->Get<I3MCPESeriesMapConstPtr>("MCPESeriesMap");
        if (!out) throw std::runtime_error("the module wrote no MCPESeriesMap");
        uint64_t k = 0;
        for (I3MCPESeriesMap::const_iterator it = out->begin(); it != out->end(); ++it)
            for (const I3MCPE &pe : it->second) {
                if (k < cap) {
                    out_string[k] = it->first.GetString();
                    out_om[k] = it->first.GetOM();
                    out_time[k] = pe.time;
                    out_photon[k] = static_cast<int64_t>(pe.ID.majorID);
                }
                ++k;
            }
        if (uniforms_used) *uniforms_used = rng->used();
        return static_cast<int64_t>(k);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

} // extern "C"
