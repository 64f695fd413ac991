// test_neighbours_cuda.cxx -- C++ test of the two neighbours on the device (I3CLSimNeighboursCUDA.h).
//   --no-gpu   argument checks with the reference's messages; construction THROWS without a CUDA device
//   --gpu      photon -> MCPE against a host restatement of I3CLSimPhotonToMCPEConverterForDOMs::Convert
//              (private/clsim/dom/I3PhotonToMCPEConverter.cxx:602-669) with explicit uniforms, the conversion attached
//              to a converter, steps made on the device (MakeSteps) and bunches generated in place (EnqueueInto)
// Exit code 0 = all checks passed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <random>
#include <set>
#include <tuple>

#include "I3CLSimNeighboursCUDA.h"
#include "test_models.h"

static int g_failed = 0, g_checked = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        ++g_checked;                                                                 \
        if (!(cond)) { ++g_failed; std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

static bool throws_with(const std::function<void()> &f, const char *needle)
{
    try {
        f();
    } catch (const I3CLSimStepToPhotonConverter_exception &e) {
        if (std::strstr(e.what(), needle)) return true;
        std::fprintf(stderr, "  exception text was: %s (wanted: %s)\n", e.what(), needle);
        return false;
    }
    std::fprintf(stderr, "  no exception (wanted: %s)\n", needle);
    return false;
}

static I3CLSimFunctionConstPtr hole_ice()
{
    // resources/ice/ppc_aha_0.80/as.holeice rows 1.. (tests/golden/angular_acceptance.json)
    const double c[11] = {0.32813, 0.63899, 0.20049, -1.2250, -0.14470, 4.1695, 0.76898, -5.8690, -2.0939, 2.3834, 1.0435};
    return I3CLSimFunctionConstPtr(new I3CLSimFunctionPolynomial(std::vector<double>(c, c + 11)));
}

static std::map<OMKey, I3CLSimFunctionConstPtr> acceptance_map(const I3CLSimFunctionConstPtr &plain, const I3CLSimFunctionConstPtr &high_qe)
{
    std::map<OMKey, I3CLSimFunctionConstPtr> m;
    for (int s = 1; s <= 8; ++s)
        for (unsigned d = 1; d <= 3; ++d) m[OMKey(s, d)] = (s > 6) ? high_qe : plain;
    return m;
}

static I3CLSimFunctionConstPtr scaled(const I3CLSimFunctionConstPtr &f, double factor)
{
    auto t = std::dynamic_pointer_cast<const I3CLSimFunctionFromTable>(f);
    std::vector<double> v;
    for (std::size_t i = 0; i < t->GetNumEntries(); ++i) v.push_back(t->GetEntryValue(i) * factor);
    return I3CLSimFunctionConstPtr(new I3CLSimFunctionFromTable(t->GetFirstWavelength(), t->GetWavelengthStepping(), v));
}

static void test_arguments(bool have_gpu)
{
    I3CLSimFunctionConstPtr bias = make_bias();
    CHECK(throws_with([&] { I3CLSimPhotonToMCPEConverterCUDA c(1, std::map<OMKey, I3CLSimFunctionConstPtr>(), hole_ice()); }, "\"WavelengthAcceptance\" parameter must not be empty"));
    CHECK(throws_with([&] { I3CLSimPhotonToMCPEConverterCUDA c(1, acceptance_map(bias, bias), I3CLSimFunctionConstPtr()); }, "\"AngularAcceptance\" parameter must not be empty"));
    CHECK(throws_with([&] { I3CLSimPhotonToMCPEConverterCUDA c(1, acceptance_map(bias, bias), bias); }, "must be an I3CLSimFunctionPolynomial"));
    if (!have_gpu) {
        CHECK(throws_with([&] { I3CLSimPhotonToMCPEConverterCUDA c(1, acceptance_map(bias, bias), hole_ice()); }, "no CPU fallback"));
        CHECK(throws_with([&] { I3CLSimStepGeneratorCUDA g(1); }, "no CPU fallback"));
    }
    CHECK(throws_with([&] { I3CLSimStepGeneratorCUDA g(1, 0, 0, -1.0, 2.61); }, "angular smearing parameters must be positive"));
}

// the reference's Convert, in double, on one photon: survival probability
static double reference_probability(const I3CLSimPhoton &p, const I3CLSimFunction &acceptance, const I3CLSimFunction &angular)
{
    double prob = p.GetWeight();
    const double cos_angle = std::max(-1., std::min(1., -std::cos(static_cast<double>(p.GetDirTheta()))));
    prob *= acceptance.GetValue(p.GetWavelength());
    prob *= angular.GetValue(cos_angle);
    return prob;
}

static void test_on_device()
{
    const std::size_t bunch = test_sized(16384, 8192);   // (the host check, tests/test_hostcheck.py, takes smaller bunches)
    const std::size_t enough = test_sized(150, 60);
    I3CLSimCUDADevice dev = {0, bunch, test_native_math()};
    I3CLSimMediumPropertiesConstPtr medium = make_medium(false);
    // generation bias = envelope of the two DOM classes (python/traysegments/common.py:186-191)
    I3CLSimFunctionConstPtr plain = scaled(make_bias(), 0.9 * 0.75), high_qe = scaled(make_bias(), 0.9 * 0.75 * 1.35);
    I3CLSimFunctionConstPtr bias = high_qe;
    std::vector<I3CLSimRandomValueConstPtr> gens(1, make_generator(bias, medium));
    auto make_conv = [&](uint64_t seed) {
        return I3CLSimModuleHelper::initializeCUDA(dev, seed, make_ring_geometry(5.0), medium, bias, gens, true, false, true, false, 0.01, NAN, 5.0, 0, 0);
    };
    I3CLSimFunctionConstPtr angular = hole_ice();
    const std::map<OMKey, I3CLSimFunctionConstPtr> acc = acceptance_map(plain, high_qe);

    // ---- photons from a real run, converted with explicit uniforms, against the host restatement
    auto conv = make_conv(11);
    conv->EnqueueSteps(make_steps(bunch, 200, 5, 5), 5);
    I3CLSimPhotonSeriesPtr photons = conv->GetConversionResult().photons;
    CHECK(photons && photons->size() > enough);
    I3CLSimPhotonToMCPEConverterCUDA mcpe(21, acc, angular, 0, 2700000);
    std::mt19937 rng(3);
    std::uniform_real_distribution<float> uni(0.f, 1.f);
    std::vector<float> u(photons->size());
    for (float &v : u) v = std::min(uni(rng), 0.99999994f);
    std::vector<clsimcu_mcpe> got = mcpe.Convert(*photons, &u);
    typedef std::tuple<int, unsigned, float, uint32_t> Key;
    std::multiset<Key> want_set, got_set;
    double max_prob = 0.;
    for (std::size_t i = 0; i < photons->size(); ++i) {
        const I3CLSimPhoton &p = (*photons)[i];
        const double r = std::sqrt(double(p.GetPosX()) * p.GetPosX() + double(p.GetPosY()) * p.GetPosY() + double(p.GetPosZ()) * p.GetPosZ());
        CHECK(std::fabs(r - 0.1651) < 0.005); // on the real-size DOM: the pancake is undone on the device
        const double prob = reference_probability(p, *acc.at(OMKey(p.GetStringID(), p.GetOMID())), *angular);
        max_prob = std::max(max_prob, prob);
        if (!(prob <= static_cast<double>(u[i]))) want_set.insert(Key(p.GetStringID(), p.GetOMID(), p.GetTime(), p.GetID()));
    }
    for (const clsimcu_mcpe &m : got) {
        got_set.insert(Key(m.string_id, m.om_id, m.time, m.identifier));
        CHECK(m.npe == 1);
    }
    CHECK(max_prob <= 1.0 && max_prob > 0.3);
    CHECK(!want_set.empty() && want_set.size() < photons->size());
    CHECK(got_set == want_set); // every survivor, exact times
    CHECK(throws_with([&] { std::vector<float> few(3); mcpe.Convert(*photons, &few); }, "one uniform per photon"));
    I3CLSimPhotonSeries bad(*photons);
    bad[0].SetWeight(-1.f);
    CHECK(throws_with([&] { mcpe.Convert(bad, &u); }, "negative weight"));

    // ---- attached: results carry photo-electrons, each one belongs to a photon of the same result
    auto conv2 = make_conv(12);
    I3CLSimPhotonToMCPEConverterCUDA attached(22, acc, angular, 0, 2750000);
    attached.AttachTo(*conv2, /*keepPhotons=*/true);
    CHECK(throws_with([&] { attached.AttachTo(*conv2, true); }, "attached once"));
    conv2->EnqueueSteps(make_steps(bunch, 200, 6, 6), 6);
    std::vector<clsimcu_mcpe> pes;
    I3CLSimStepToPhotonConverter::ConversionResult_t res = conv2->GetConversionResultWithMCPEs(pes);
    CHECK(res.identifier == 6 && res.photons && res.photons->size() > enough);
    CHECK(!pes.empty() && pes.size() < res.photons->size());
    std::multiset<Key> photon_keys;
    for (const I3CLSimPhoton &p : *res.photons) photon_keys.insert(Key(p.GetStringID(), p.GetOMID(), p.GetTime(), p.GetID()));
    bool all_found = true;
    for (const clsimcu_mcpe &m : pes) all_found = all_found && photon_keys.count(Key(m.string_id, m.om_id, m.time, m.identifier)) > 0;
    CHECK(all_found);
    const double frac = pes.size() / double(res.photons->size()), frac_explicit = got.size() / double(photons->size());
    CHECK(std::fabs(frac - frac_explicit) < 6.0 * std::sqrt(0.25 / photons->size()) + 0.02);

    // ---- steps made on the device
    I3CLSimStepGeneratorCUDA gen(31, 0, 2800000);
    I3CLSimStepGeneratorCUDA::Source track;
    std::memset(&track, 0, sizeof track);
    track.x = -20; track.y = 5; track.z = -40; track.time = 10;
    track.dirX = 0.6; track.dirY = 0.0; track.dirZ = 0.8;
    track.particleIdentifier = 77;
    track.photonsPerStep = 200; track.numSteps = 3000; track.numPhotonsInLastStep = 42;
    track.isCascade = false; track.stepIsCascadeLike = false; track.length = 80.0;
    I3CLSimStepGeneratorCUDA::Source smear = track;
    smear.stepIsCascadeLike = true; smear.numSteps = 1000; smear.numPhotonsInLastStep = 0; smear.particleIdentifier = 78;
    I3CLSimStepGeneratorCUDA::Source casc = track;
    casc.isCascade = true; casc.pa = 4.2; casc.pb = 0.62; casc.numSteps = 500; casc.numPhotonsInLastStep = 7; casc.particleIdentifier = 79;
    std::vector<I3CLSimStepGeneratorCUDA::Source> sources = {track, smear, casc};
    I3CLSimStepSeriesPtr steps = gen.MakeSteps(sources);
    CHECK(steps->size() == 3001 + 1000 + 501);
    uint64_t n_photons = 0;
    bool muon_like_ok = true, on_axis = true;
    for (std::size_t i = 0; i < steps->size(); ++i) {
        const I3CLSimStep &s = (*steps)[i];
        n_photons += s.GetNumPhotons();
        if (i < 3001) muon_like_ok = muon_like_ok && s.GetLength() == 80.f && s.GetPosX() == -20.f && s.GetID() == 77 && std::fabs(s.GetDirTheta() - std::acos(0.8)) < 1e-6;
        else {
            // cascade-like steps sit on the particle axis, downstream, in time with c
            const double along = (s.GetPosX() + 20) * 0.6 + (s.GetPosZ() + 40) * 0.8;
            const double ox = s.GetPosX() + 20 - along * 0.6, oy = s.GetPosY() - 5, oz = s.GetPosZ() + 40 - along * 0.8;
            on_axis = on_axis && along >= -1e-4 && std::sqrt(ox * ox + oy * oy + oz * oz) < 1e-4 && std::fabs(s.GetTime() - 10 - along / 0.299792458) < 1e-3 &&
                      s.GetLength() == 0.001f && (i < 4001 ? along <= 80.0 + 1e-4 : true);
        }
    }
    CHECK(n_photons == 3000 * 200 + 42 + 1000 * 200 + 500 * 200 + 7);
    CHECK(muon_like_ok);
    CHECK(on_axis);
    CHECK((*steps)[3000].GetNumPhotons() == 42 && (*steps)[4501].GetNumPhotons() == 7);

    // ---- the same entries as a bunch that only ever exists on the device
    auto conv3 = make_conv(13);
    const std::size_t sent = gen.EnqueueInto(*conv3, sources, 9);
    CHECK(sent == steps->size());
    I3CLSimStepToPhotonConverter::ConversionResult_t r3 = conv3->GetConversionResult();
    CHECK(r3.identifier == 9 && r3.photons && !r3.photons->empty());
    std::set<uint32_t> ids;
    for (const I3CLSimPhoton &p : *r3.photons) ids.insert(p.GetID());
    CHECK(*ids.begin() >= 77 && *ids.rbegin() <= 79);
    CHECK(conv3->GetStatistics()["TotalNumPhotonsGenerated"] == static_cast<double>(n_photons));
    sources[0].numSteps = bunch;
    CHECK(throws_with([&] { gen.EnqueueInto(*conv3, sources, 10); }, "greater than maximum number of work items"));
    sources[0].numSteps = 10; sources[0].length = 0.0;
    CHECK(throws_with([&] { gen.MakeSteps(sources); }, "cascade segment with length"));
}

int main(int argc, char **argv)
{
    const bool gpu = argc > 1 && std::strcmp(argv[1], "--gpu") == 0;
    const bool have_device = gpu || (argc > 1 && std::strcmp(argv[1], "--gpu-args-only") == 0);
    try {
        test_arguments(have_device);
        if (gpu) test_on_device();
    } catch (const std::exception &e) {
        std::fprintf(stderr, "unexpected exception: %s\n", e.what());
        return 2;
    }
    std::printf("%d checks, %d failed\n", g_checked, g_failed);
    return g_failed == 0 ? 0 : 1;
}
