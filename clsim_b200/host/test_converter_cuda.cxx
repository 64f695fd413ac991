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
    I3CLSimCUDADevice dev = {0, bunch, test_native_math()};
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

    // photon history (on the fast kernel: native math requested)
    {
        I3CLSimCUDADevice precise = {0, 1024, test_native_math()};
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
