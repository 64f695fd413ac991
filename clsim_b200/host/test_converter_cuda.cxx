// test_converter_cuda.cxx -- C++ test of I3CLSimStepToPhotonConverterCUDA, in the spirit of the reference's
// own interface tests (resources/tests/testCLSimServer.py: many producers, results matched by identifier;
// private/test/ has none for the converter).  Modes:
//   --no-gpu     contract checks that need no device: setter/Compile/EnqueueSteps error behaviour and messages
//                (…ConverterOpenCL.cxx:492-508, 1324-1544), flattening, and that Initialize() THROWS without a
//                CUDA device (there is no CPU fallback)
//   --gpu        the above (minus the no-device throw) plus a real run: 5 threads x (EnqueueSteps ;
//                GetConversionResult) against one converter, any-order results, conservation through the
//                statistics, real string/OM IDs, photon history, destructor with work in flight
//   --describe   print the flattened device tables as JSON (tests compare it with the Python flattening)
// Exit code 0 = all checks passed.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <random>
#include <set>
#include <thread>

#include "I3CLSimStepToPhotonConverterCUDA.h"

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

static const double nm = 1e-9, deg = M_PI / 180.0;

// 24-DOM ring (resources/scripts/benchmark.py:63-114)
static I3CLSimSimpleGeometryConstPtr make_ring_geometry(double oversize)
{
    const double radius = 120.0;
    const double dirs[8][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}};
    std::shared_ptr<I3CLSimSimpleGeometryUserConfigurable> g(new I3CLSimSimpleGeometryUserConfigurable(0.16510 * oversize, 24));
    std::size_t at = 0;
    for (int s = 0; s < 8; ++s) {
        const double len = std::sqrt(dirs[s][0] * dirs[s][0] + dirs[s][1] * dirs[s][1]);
        const double dz[3] = {radius, 0.0, -radius};
        for (int d = 0; d < 3; ++d, ++at) {
            g->SetStringID(at, s + 1);
            g->SetDomID(at, d + 1);
            g->SetPosX(at, dirs[s][0] / len * radius);
            g->SetPosY(at, dirs[s][1] / len * radius);
            g->SetPosZ(at, dz[d]);
            g->SetSubdetector(at, "Unknown");
        }
    }
    return g;
}

// A 12-layer ice in the shape MakeIceCubeMediumProperties.py:166-230 builds (values of the order of SpiceMie's)
static I3CLSimMediumPropertiesConstPtr make_medium(bool with_tilt_and_anisotropy)
{
    const uint32_t layers = 12;
    std::shared_ptr<I3CLSimMediumProperties> m(new I3CLSimMediumProperties(0.9216, layers, -60.0, 10.0, -870.0, 1940.0));
    m->SetForcedMinWlen(265 * nm);
    m->SetForcedMaxWlen(675 * nm);
    const double kappa = 1.08410680294, A = 6954.09033203, B = 6617.75439453, alpha = 0.898608505726;
    I3CLSimFunctionConstPtr phase(new I3CLSimFunctionRefIndexIceCube("phase"));
    I3CLSimFunctionConstPtr group(new I3CLSimFunctionRefIndexIceCube("group"));
    for (uint32_t l = 0; l < layers; ++l) {
        const double be400 = 0.020 + 0.004 * std::sin(0.9 * l), adust = 0.006 + 0.002 * std::cos(0.7 * l), dtau = 5.0 + 0.3 * l;
        const double g = 0.9;
        m->SetAbsorptionLength(l, I3CLSimFunctionConstPtr(new I3CLSimFunctionAbsLenIceCube(kappa, A, B, std::pow(400.0, kappa), 0.0, adust, dtau)));
        m->SetScatteringLength(l, I3CLSimFunctionConstPtr(new I3CLSimFunctionScatLenIceCube(alpha, be400 / (1.0 - g))));
        m->SetPhaseRefractiveIndex(l, phase);
        m->SetGroupRefractiveIndexOverride(l, group);
    }
    I3CLSimRandomValueConstPtr sl(new I3CLSimRandomValueSimplifiedLiu(0.9)), hg(new I3CLSimRandomValueHenyeyGreenstein(0.9));
    m->SetScatteringCosAngleDistribution(I3CLSimRandomValueConstPtr(new I3CLSimRandomValueMixed(0.45, sl, hg)));
    if (with_tilt_and_anisotropy) {
        std::vector<double> dist = {-500.0, -100.0, 0.0, 150.0, 400.0}, zc;
        std::vector<std::vector<double> > corr(dist.size());
        for (int k = 0; k < 20; ++k) zc.push_back(-70.0 + 7.5 * k);
        for (std::size_t i = 0; i < dist.size(); ++i)
            for (int k = 0; k < 20; ++k) corr[i].push_back(0.004 * dist[i] * std::cos(0.2 * k));
        m->SetIceTiltZShift(I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldIceTiltZShift(dist, zc, corr, 225.0 * deg)));
        m->SetDirectionalAbsorptionLengthCorrection(I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldAnisotropyAbsLenScaling(216.0 * deg, 0.04, -0.08)));
        // GetSpiceLeaAnisotropyTransforms.py:39-101: T^T diag(k1,k2,1/(k1 k2))^{+-1} T with T the rotation by the azimuth
        const double k1 = std::exp(0.04), k2 = std::exp(-0.08), kz = 1.0 / (k1 * k2), ca = std::cos(216.0 * deg), sa = std::sin(216.0 * deg);
        double pre[9], post[9];
        const double T[9] = {ca, sa, 0, -sa, ca, 0, 0, 0, 1};
        const double kpre[3] = {k1, k2, kz}, kpost[3] = {1 / k1, 1 / k2, 1 / kz};
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                pre[3 * r + c] = post[3 * r + c] = 0;
                for (int j = 0; j < 3; ++j) {
                    pre[3 * r + c] += T[3 * j + r] * kpre[j] * T[3 * j + c];
                    post[3 * r + c] += T[3 * j + r] * kpost[j] * T[3 * j + c];
                }
            }
        m->SetPreScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformMatrix(pre, true)));
        m->SetPostScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformMatrix(post, true)));
    }
    return m;
}

static I3CLSimFunctionConstPtr make_bias()
{
    // shape of GetIceCubeDOMAcceptance (43 values, 260 nm + 10 nm * i), numbers synthetic
    std::vector<double> v;
    for (int i = 0; i < 43; ++i) v.push_back(0.02 + 0.11 * std::exp(-0.5 * std::pow((i - 14) / 7.0, 2)));
    return I3CLSimFunctionConstPtr(new I3CLSimFunctionFromTable(260 * nm, 10 * nm, v));
}

// makeCherenkovWavelengthGenerator with a table bias (I3CLSimModuleHelper.cxx:224-256): spectrum on the bias grid
static I3CLSimRandomValueConstPtr make_generator(const I3CLSimFunctionConstPtr &bias, const I3CLSimMediumPropertiesConstPtr &medium)
{
    auto t = std::dynamic_pointer_cast<const I3CLSimFunctionFromTable>(bias);
    std::vector<double> y;
    for (std::size_t i = 0; i < t->GetNumEntries(); ++i) {
        const double w = t->GetEntryWavelength(i), n = medium->GetPhaseRefractiveIndex(0)->GetValue(w);
        y.push_back(t->GetEntryValue(i) * (2.0 * M_PI / 137.0) / (w * w) * (1.0 - 1.0 / (n * n)));
    }
    return I3CLSimRandomValueConstPtr(new I3CLSimRandomValueInterpolatedDistribution(t->GetFirstWavelength(), t->GetWavelengthStepping(), y));
}

static I3CLSimStepSeriesPtr make_steps(std::size_t n, uint32_t photons, uint32_t id, unsigned seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    I3CLSimStepSeriesPtr s(new I3CLSimStepSeries(n));
    for (std::size_t i = 0; i < n; ++i) {
        I3CLSimStep &st = (*s)[i];
        std::memset(&st, 0, sizeof(st));
        st.SetPosX(static_cast<float>(30.0 * u(rng)));
        st.SetPosY(static_cast<float>(30.0 * u(rng)));
        st.SetPosZ(static_cast<float>(30.0 * u(rng)));
        st.SetTime(static_cast<float>(100.0 + 50.0 * u(rng)));
        double x, y, z, r2;
        do { x = u(rng); y = u(rng); z = u(rng); r2 = x * x + y * y + z * z; } while (r2 > 1.0 || r2 < 1e-4);
        st.SetDir(x, y, z);
        st.SetLength(1.0f);
        st.SetBeta(1.0f);
        st.SetNumPhotons(photons);
        st.SetWeight(1.0f);
        st.SetID(id);
        st.SetSourceType(0);
    }
    return s;
}

static std::shared_ptr<I3CLSimStepToPhotonConverterCUDA> configured(bool aniso, bool native = true)
{
    std::shared_ptr<I3CLSimStepToPhotonConverterCUDA> c(new I3CLSimStepToPhotonConverterCUDA(12345, native));
    I3CLSimMediumPropertiesConstPtr medium = make_medium(aniso);
    I3CLSimFunctionConstPtr bias = make_bias();
    c->SetDevice(0);
    c->SetWlenGenerators(std::vector<I3CLSimRandomValueConstPtr>(1, make_generator(bias, medium)));
    c->SetWlenBias(bias);
    c->SetMediumProperties(medium);
    c->SetGeometry(make_ring_geometry(5.0));
    c->SetStopDetectedPhotons(true);
    c->SetDOMPancakeFactor(5.0);
    return c;
}

static void test_contract_without_device()
{
    // Compile checks, in the reference's order (…OpenCL.cxx:492-508)
    I3CLSimMediumPropertiesConstPtr medium = make_medium(false);
    I3CLSimFunctionConstPtr bias = make_bias();
    {
        I3CLSimStepToPhotonConverterCUDA c;
        CHECK(!c.IsInitialized());
        CHECK(throws_with([&] { c.Compile(); }, "WlenGenerators not set!"));
        c.SetWlenGenerators(std::vector<I3CLSimRandomValueConstPtr>(1, make_generator(bias, medium)));
        CHECK(throws_with([&] { c.Compile(); }, "WlenBias not set!"));
        c.SetWlenBias(bias);
        CHECK(throws_with([&] { c.Compile(); }, "MediumProperties not set!"));
        c.SetMediumProperties(medium);
        CHECK(throws_with([&] { c.Compile(); }, "Geometry not set!"));
        c.SetGeometry(make_ring_geometry(5.0));
        CHECK(throws_with([&] { c.Compile(); }, "Device not selected!"));
        c.SetDevice(0);
        c.SetSaveAllPhotons(true);
        c.SetStopDetectedPhotons(true);
        CHECK(throws_with([&] { c.Compile(); }, "both the saveAllPhotons and stopDetectedPhotons"));
        c.SetSaveAllPhotons(false);
        CHECK(throws_with([&] { c.GetMaxWorkgroupSize(); }, "compile the kernel first"));
        CHECK(throws_with([&] { c.SetWorkgroupSize(64); }, "compile the kernel first"));
        c.Compile();
        c.Compile(); // silently
        CHECK(c.GetMaxWorkgroupSize() == 1024);
        CHECK(throws_with([&] { c.SetWorkgroupSize(4096); }, "Workgroup size too large!"));
        CHECK(throws_with([&] { c.SetMaxNumWorkitems(0); }, "Invalid maximum number of work items!"));
        CHECK(throws_with([&] { c.SetDoublePrecision(true); }, "DoublePrecision"));
        // not initialized: the hot calls and the queue queries throw (…OpenCL.cxx:1525-1562, 1604-1607)
        CHECK(throws_with([&] { c.EnqueueSteps(make_steps(8, 10, 0, 1), 0); }, "is not initialized!"));
        CHECK(throws_with([&] { c.GetConversionResult(); }, "is not initialized!"));
        CHECK(throws_with([&] { c.QueueSize(); }, "is not initialized!"));
        CHECK(throws_with([&] { c.MorePhotonsAvailable(); }, "is not initialized!"));
        CHECK(c.GetStatistics().empty());
        CHECK(c.GetWorkgroupSize() == 1);
        CHECK(c.GetMaxNumWorkitems() == 10240); // class default …OpenCL.cxx:86
    }
    // classes outside the hot path are refused by name, not silently approximated
    {
        auto c = configured(false);
        struct Unknown : public I3CLSimRandomValue {};
        c->SetWlenGenerators(std::vector<I3CLSimRandomValueConstPtr>(1, I3CLSimRandomValueConstPtr(new Unknown())));
        CHECK(throws_with([&] { c->Compile(); }, "does not know"));
    }
    // the flattening: table description comes back and names the pieces
    {
        auto c = configured(true);
        c->Compile();
        const std::string js = c->DescribeTables();
        CHECK(js.find("\"medium\"") != std::string::npos || js.find("medium") != std::string::npos);
        CHECK(js.size() > 1000);
    }
}

static void test_no_fallback()
{
    auto c = configured(false);
    c->SetMaxNumWorkitems(1024);
    bool threw = false;
    try {
        c->Initialize();
    } catch (const I3CLSimStepToPhotonConverter_exception &e) {
        threw = true;
        std::printf("Initialize without a device throws: %s\n", e.what());
    }
    CHECK(threw);
    CHECK(!c->IsInitialized());
}

static void test_on_device()
{
    const std::size_t bunch = 4096;
    const uint32_t photons = 100;
    I3CLSimCUDADevice dev = {0, bunch, true};
    I3CLSimMediumPropertiesConstPtr medium = make_medium(true);
    I3CLSimFunctionConstPtr bias = make_bias();
    std::vector<I3CLSimRandomValueConstPtr> gens(1, make_generator(bias, medium));
    auto conv = I3CLSimModuleHelper::initializeCUDA(dev, 777, make_ring_geometry(5.0), medium, bias, gens, /*doubleBuffering*/ true,
                                                    /*doublePrecision*/ false, /*stop*/ true, /*saveAll*/ false, 0.01, NAN, 5.0, 0, 0);
    CHECK(conv->IsInitialized());
    CHECK(conv->GetWorkgroupSize() == 1);
    CHECK(conv->GetMaxNumWorkitems() == bunch);
    // every setter throws once initialized (…OpenCL.cxx:1324-1523), Initialize twice too (:219-220)
    CHECK(throws_with([&] { conv->SetDevice(0); }, "already initialized!"));
    CHECK(throws_with([&] { conv->SetDOMPancakeFactor(1.0); }, "already initialized!"));
    CHECK(throws_with([&] { conv->SetGeometry(make_ring_geometry(1.0)); }, "already initialized!"));
    CHECK(throws_with([&] { conv->SetMaxNumWorkitems(5); }, "already initialized!"));
    CHECK(throws_with([&] { conv->Initialize(); }, "already initialized!"));
    CHECK(throws_with([&] { conv->Compile(); }, "already initialized!"));
    // EnqueueSteps argument errors (…OpenCL.cxx:1530-1540)
    CHECK(throws_with([&] { conv->EnqueueSteps(I3CLSimStepSeriesConstPtr(), 0); }, "Steps pointer is (null)!"));
    CHECK(throws_with([&] { conv->EnqueueSteps(I3CLSimStepSeriesPtr(new I3CLSimStepSeries()), 0); }, "Steps are empty!"));
    CHECK(throws_with([&] { conv->EnqueueSteps(make_steps(bunch + 1, 1, 0, 1), 0); }, "greater than maximum number of work items"));
    CHECK(!conv->MorePhotonsAvailable());
    CHECK(conv->QueueSize() == 0);

    // 5 producer threads per converter, each EnqueueSteps then GetConversionResult, results matched by identifier
    // afterwards (I3CLSimServer.cxx:126-135, 310-343: any thread may receive any bunch)
    const int threads = 5, rounds = 4;
    std::mutex mu;
    std::map<uint32_t, std::size_t> got; // identifier -> photons at DOMs
    std::atomic<uint64_t> generated(0);
    std::set<int> string_ids, om_ids;
    bool all_ids_match = true, never_null = true;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            for (int r = 0; r < rounds; ++r) {
                const uint32_t id = 1000u * (t + 1) + r;
                const std::size_t n = (r % 2) ? bunch : bunch / 2 + 7 * t; // ragged bunch sizes
                conv->EnqueueSteps(make_steps(n, photons, id, id), id);
                generated += static_cast<uint64_t>(n) * photons;
                I3CLSimStepToPhotonConverter::ConversionResult_t res = conv->GetConversionResult();
                std::lock_guard<std::mutex> lock(mu);
                if (!res.photons) { never_null = false; continue; }
                got[res.identifier] += res.photons->size();
                for (const I3CLSimPhoton &p : *res.photons) {
                    if (p.GetID() != res.identifier) all_ids_match = false;
                    string_ids.insert(p.GetStringID());
                    om_ids.insert(p.GetOMID());
                }
            }
        });
    for (std::thread &th : pool) th.join();
    CHECK(never_null);
    CHECK(all_ids_match);
    CHECK(got.size() == static_cast<std::size_t>(threads * rounds)); // every identifier came back exactly once
    CHECK(!conv->MorePhotonsAvailable());
    std::size_t hits = 0;
    for (auto &kv : got) hits += kv.second;
    CHECK(hits > 100);
    CHECK(*string_ids.begin() >= 1 && *string_ids.rbegin() <= 8); // real IDs, not table indices
    CHECK(*om_ids.begin() >= 1 && *om_ids.rbegin() <= 3);
    std::map<std::string, double> st = conv->GetStatistics();
    for (const char *k : {"TotalDeviceTime", "TotalHostTime", "NumKernelCalls", "TotalNumPhotonsGenerated", "TotalNumPhotonsAtDOMs",
                          "AverageDeviceTimePerPhoton", "AverageHostTimePerPhoton", "DeviceUtilization"})
        CHECK(st.count(k) == 1);
    CHECK(st["NumKernelCalls"] == threads * rounds);
    CHECK(st["TotalNumPhotonsGenerated"] == static_cast<double>(generated.load())); // conservation
    CHECK(st["TotalNumPhotonsAtDOMs"] == static_cast<double>(hits));
    std::printf("device run: %d bunches, %llu photons generated, %zu at DOMs, device time %.3f ms\n", threads * rounds,
                static_cast<unsigned long long>(generated.load()), hits, st["TotalDeviceTime"] * 1e-6);

    // hits sit on the true DOM sphere after the pancake is undone (propagation_kernel.c.cl:340-355)
    conv->EnqueueSteps(make_steps(bunch, photons, 42, 4242), 42);
    I3CLSimStepToPhotonConverter::ConversionResult_t res = conv->GetConversionResult();
    CHECK(res.identifier == 42 && res.photons && !res.photonHistories);
    bool on_sphere = true;
    for (const I3CLSimPhoton &p : *res.photons) {
        const double r = std::sqrt(double(p.GetPosX()) * p.GetPosX() + double(p.GetPosY()) * p.GetPosY() + double(p.GetPosZ()) * p.GetPosZ());
        if (std::fabs(r - 0.16510) > 1e-3) on_sphere = false;
    }
    CHECK(on_sphere);

    // photon history through the precise (reference-order) kernel
    {
        I3CLSimCUDADevice precise = {0, 1024, true}; // native math requested: history still routes to the reference-order kernel
        auto hc = I3CLSimModuleHelper::initializeCUDA(precise, 5, make_ring_geometry(5.0), medium, bias, gens, false, false, true, false, 0.01, NAN,
                                                      5.0, /*history*/ 4, 0);
        hc->EnqueueSteps(make_steps(1024, 200, 7, 77), 7);
        I3CLSimStepToPhotonConverter::ConversionResult_t hr = hc->GetConversionResult();
        CHECK(hr.photons && hr.photonHistories && hr.photonHistories->size() == hr.photons->size());
        bool sizes_ok = true;
        for (std::size_t i = 0; i < hr.photons->size(); ++i)
            if ((*hr.photonHistories)[i].size() != std::min<std::size_t>((*hr.photons)[i].GetNumScatters(), 4)) sizes_ok = false;
        CHECK(sizes_ok);
    }
    // destructor with work still queued must interrupt and join (…OpenCL.cxx:110-145), not hang
    for (int i = 0; i < 3; ++i) conv->EnqueueSteps(make_steps(bunch, photons, 900 + i, 900 + i), 900 + i);
    conv.reset();
}

int main(int argc, char **argv)
{
    const std::string mode = argc > 1 ? argv[1] : "--no-gpu";
    try {
        if (mode == "--describe") {
            auto c = configured(argc > 2 && std::string(argv[2]) == "aniso");
            c->Compile();
            std::printf("%s\n", c->DescribeTables().c_str());
            return 0;
        }
        test_contract_without_device();
        if (mode == "--gpu") test_on_device();
        else test_no_fallback();
    } catch (const std::exception &e) {
        std::fprintf(stderr, "unexpected exception: %s\n", e.what());
        return 2;
    }
    std::printf("%d checks, %d failed\n", g_checked, g_failed);
    return g_failed ? 1 : 0;
}
