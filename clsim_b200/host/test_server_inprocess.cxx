// test_server_inprocess.cxx -- the reference's own interface test of the server seam, restated in C++:
// resources/tests/testCLSimServer.py (DummyConverter; a client sends 10 ragged bunches and checks every result
// against its input by identifier; three clients at once through one server).  Modes:
//   (no argument)  DummyConverter only, no device needed
//   --gpu [N]      additionally N real I3CLSimStepToPhotonConverterCUDA behind one server (one per CUDA ordinal;
//                  with fewer devices than N the ordinals wrap), three clients, conservation by statistics
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>

#include "I3CLSimServerInProcess.h"
#include "clsimcuda.h"
#include "test_models.h"

static int g_failed = 0, g_checked = 0;
static std::mutex g_check_mutex;
#define CHECK(cond)                                                                                          \
    do {                                                                                                     \
        std::lock_guard<std::mutex> check_lock(g_check_mutex);                                               \
        ++g_checked;                                                                                         \
        if (!(cond)) { ++g_failed; std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

// testCLSimServer.py:11-43
static I3CLSimPhoton dummy_photon(const I3CLSimStep &step)
{
    I3CLSimPhoton p;
    std::memset(&p, 0, sizeof(p));
    p.posAndTime[0] = step.GetPosX(); p.posAndTime[1] = step.GetPosY(); p.posAndTime[2] = step.GetPosZ(); p.posAndTime[3] = step.GetTime();
    p.dir[0] = step.GetDirTheta(); p.dir[1] = step.GetDirPhi();
    p.weight = step.GetWeight();
    p.numScatters = 3;
    p.omID = 52;
    p.stringID = 23;
    return p;
}
static I3CLSimPhotonHistory dummy_photon_history(const I3CLSimPhoton &photon)
{
    I3CLSimPhotonHistory h;
    for (uint32_t i = 0; i < photon.GetNumScatters(); ++i) h.push_back(float(i), float(i + 0.5), float(i + 3.14), float(i));
    return h;
}

class DummyConverter : public I3CLSimStepToPhotonConverter {
public:
    explicit DummyConverter(std::size_t workgroup = 1, std::size_t maxItems = 64) : workgroup_(workgroup), maxItems_(maxItems) {}
    void SetWlenGenerators(const std::vector<I3CLSimRandomValueConstPtr> &) override {}
    void SetWlenBias(I3CLSimFunctionConstPtr) override {}
    void SetMediumProperties(I3CLSimMediumPropertiesConstPtr) override {}
    void SetGeometry(I3CLSimSimpleGeometryConstPtr) override {}
    void Initialize() override {}
    bool IsInitialized() const override { return true; }
    std::size_t GetWorkgroupSize() const override { return workgroup_; }
    std::size_t GetMaxNumWorkitems() const override { return maxItems_; }
    std::size_t QueueSize() const override { return 0; }
    bool MorePhotonsAvailable() const override { return false; }
    void EnqueueSteps(I3CLSimStepSeriesConstPtr steps, uint32_t id) override
    {
        std::lock_guard<std::mutex> lock(mutex_);
        input_queue_.push_back(std::make_pair(steps, id));
    }
    ConversionResult_t GetConversionResult() override
    {
        std::pair<I3CLSimStepSeriesConstPtr, uint32_t> job;
        {
            std::lock_guard<std::mutex> lock(mutex_);
            job = input_queue_.front();
            input_queue_.pop_front();
        }
        I3CLSimPhotonSeriesPtr photons(new I3CLSimPhotonSeries());
        I3CLSimPhotonHistorySeriesPtr history(new I3CLSimPhotonHistorySeries());
        for (const I3CLSimStep &s : *job.first) {
            photons->push_back(dummy_photon(s));
            history->push_back(dummy_photon_history(photons->back()));
        }
        return ConversionResult_t(job.second, photons, history);
    }
    std::map<std::string, double> GetStatistics() const override { return {{"NumKernelCalls", 1.0}}; }

private:
    std::size_t workgroup_, maxItems_;
    std::mutex mutex_;
    std::deque<std::pair<I3CLSimStepSeriesConstPtr, uint32_t> > input_queue_;
};

// testCLSimServer.py:45-78, for anything with EnqueueSteps / GetConversionResult
template <class Client> static void test_client(Client &client, int num_bunches, unsigned seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    std::vector<I3CLSimStepSeriesPtr> input_steps;
    for (int i = 0; i < num_bunches; ++i) {
        I3CLSimStepSeriesPtr steps(new I3CLSimStepSeries(int(u(rng) * 10) + 1));
        for (I3CLSimStep &s : *steps) {
            std::memset(&s, 0, sizeof(s));
            s.SetPosX(u(rng)); s.SetPosY(u(rng)); s.SetPosZ(u(rng));
            s.SetDir(u(rng) + 0.01, u(rng), u(rng));
            s.SetTime(u(rng));
            s.SetWeight(u(rng));
        }
        input_steps.push_back(steps);
        client.EnqueueSteps(steps, uint32_t(i));
    }
    std::vector<int> seen(num_bunches, 0);
    for (int i = 0; i < num_bunches; ++i) {
        I3CLSimStepToPhotonConverter::ConversionResult_t result = client.GetConversionResult();
        CHECK(result.identifier < uint32_t(num_bunches));
        if (result.identifier >= uint32_t(num_bunches)) continue;
        seen[result.identifier]++;
        const I3CLSimStepSeries &input = *input_steps[result.identifier];
        CHECK(result.photons && result.photons->size() == input.size());
        CHECK(result.photonHistories && result.photonHistories->size() == input.size());
        if (!result.photons || result.photons->size() != input.size()) continue;
        bool same = true;
        for (std::size_t k = 0; k < input.size(); ++k) {
            const I3CLSimPhoton &p = (*result.photons)[k];
            const I3CLSimStep &s = input[k];
            same = same && p.GetNumScatters() == 3 && p.GetOMID() == 52 && p.GetStringID() == 23 && p.GetPosX() == s.GetPosX() &&
                   p.GetPosY() == s.GetPosY() && p.GetPosZ() == s.GetPosZ() && p.GetDirTheta() == s.GetDirTheta() &&
                   p.GetDirPhi() == s.GetDirPhi() && p.GetTime() == s.GetTime() && p.GetWeight() == s.GetWeight();
            const I3CLSimPhotonHistory &h = (*result.photonHistories)[k];
            same = same && h.size() == 3 && h.GetX(2) == 2.f && h.GetY(1) == 1.5f && h.GetDistanceInAbsorptionLengths(2) == 2.f;
        }
        CHECK(same);
    }
    bool once = true;
    for (int v : seen) once = once && (v == 1);
    CHECK(once);
}

static void test_with_dummy_converters()
{
    // first, the test passes when the converter is called directly (testCLSimServer.py:80-81)
    {
        DummyConverter direct;
        test_client(direct, 10, 1);
    }
    // now through the server, three clients at once (testCLSimServer.py:83-102; threads instead of processes)
    {
        std::vector<I3CLSimStepToPhotonConverterPtr> converters(1, I3CLSimStepToPhotonConverterPtr(new DummyConverter()));
        I3CLSimServerInProcess server(converters);
        std::vector<std::thread> pool;
        for (int c = 0; c < 3; ++c)
            pool.emplace_back([&server, c] {
                std::shared_ptr<I3CLSimClientInProcess> client = server.Connect();
                CHECK(client->GetWorkgroupSize() == 1);
                CHECK(client->GetMaxNumWorkitems() == 64);
                test_client(*client, 10, 100 + c);
                // nothing pending: an empty result comes back at once (I3CLSimServer.cxx:394-397)
                I3CLSimStepToPhotonConverter::ConversionResult_t none = client->GetConversionResult();
                CHECK(!none.photons && none.identifier == 0);
            });
        for (std::thread &t : pool) t.join();
        CHECK(server.GetStatistics().count("NumKernelCalls") == 1); // one converter: no suffix
    }
    // bunch-size harmonisation (I3CLSimServer.cxx:95-113)
    {
        std::vector<I3CLSimStepToPhotonConverterPtr> converters;
        converters.push_back(I3CLSimStepToPhotonConverterPtr(new DummyConverter(4, 100)));
        converters.push_back(I3CLSimStepToPhotonConverterPtr(new DummyConverter(6, 64)));
        I3CLSimServerInProcess server(converters);
        CHECK(server.GetWorkgroupSize() == 12);
        CHECK(server.GetMaxNumWorkitems() == 60);
        std::map<std::string, double> st = server.GetStatistics();
        CHECK(st.count("NumKernelCalls_0") == 1 && st.count("NumKernelCalls_1") == 1);
        bool threw = false;
        try {
            std::vector<I3CLSimStepToPhotonConverterPtr> bad;
            bad.push_back(I3CLSimStepToPhotonConverterPtr(new DummyConverter(64, 64)));
            bad.push_back(I3CLSimStepToPhotonConverterPtr(new DummyConverter(48, 100)));
            I3CLSimServerInProcess incompatible(bad);
        } catch (const std::runtime_error &e) {
            threw = std::strstr(e.what(), "incompatible") != nullptr;
        }
        CHECK(threw);
        threw = false;
        try {
            I3CLSimServerInProcess empty((std::vector<I3CLSimStepToPhotonConverterPtr>()));
        } catch (const std::runtime_error &) { threw = true; }
        CHECK(threw);
    }
}

// A converter that fails must fail its clients, not hand them empty photon series (light would be lost silently).
class FailingConverter : public DummyConverter {
public:
    explicit FailingConverter(int good) : DummyConverter(1, 64), good_(good) {}
    void EnqueueSteps(I3CLSimStepSeriesConstPtr steps, uint32_t id) override
    {
        if (good_.fetch_sub(1) <= 0) throw I3CLSimStepToPhotonConverter_exception("device lost");
        DummyConverter::EnqueueSteps(steps, id);
    }

private:
    std::atomic<int> good_;
};

static void test_failing_converter_fails_the_clients()
{
    std::vector<I3CLSimStepToPhotonConverterPtr> converters(1, I3CLSimStepToPhotonConverterPtr(new FailingConverter(2)));
    I3CLSimServerInProcess server(converters);
    std::shared_ptr<I3CLSimClientInProcess> client = server.Connect();
    int results = 0;
    bool threw = false;
    try {
        for (uint32_t i = 0; i < 6; ++i) client->EnqueueSteps(make_steps(8, 10, 1, i), i);
        for (uint32_t i = 0; i < 6; ++i) {
            I3CLSimStepToPhotonConverter::ConversionResult_t res = client->GetConversionResult();
            CHECK(res.photons && res.photons->size() == 8);   // a result that does arrive is a real one
            ++results;
        }
    } catch (const std::runtime_error &e) {
        threw = std::strstr(e.what(), "device lost") != nullptr;
    }
    CHECK(threw);
    CHECK(results <= 2);
    CHECK(server.Failure().find("device lost") != std::string::npos);
    threw = false;
    try {
        client->EnqueueSteps(make_steps(8, 10, 1, 0), 99);
    } catch (const std::runtime_error &) { threw = true; }
    CHECK(threw);
}

static void test_with_cuda_converters(int want)
{
    int devices = 0;
    CHECK(clsimcu_device_count(&devices) == CLSIMCU_OK && devices >= 1);
    if (devices < 1) return;
    const std::size_t bunch = test_sized(8192, 1024);
    I3CLSimMediumPropertiesConstPtr medium = make_medium(false);
    I3CLSimFunctionConstPtr bias = make_bias();
    std::vector<I3CLSimRandomValueConstPtr> gens(1, make_generator(bias, medium));
    std::vector<I3CLSimStepToPhotonConverterPtr> converters;
    for (int i = 0; i < want; ++i) {
        I3CLSimCUDADevice dev = {i % devices, bunch, test_native_math()};
        // each converter its own slice of the safe-prime multiplier table: independent RNG streams
        converters.push_back(I3CLSimModuleHelper::initializeCUDA(dev, 1000 + i, make_ring_geometry(5.0), medium, bias, gens, true, false, true, false,
                                                                0.01, NAN, 5.0, 0, 0, uint64_t(i) * 2u * 160u * 1024u));
    }
    I3CLSimServerInProcess server(converters);
    CHECK(server.GetWorkgroupSize() == 1 && server.GetMaxNumWorkitems() == bunch);
    const int clients = 3, rounds = 6;
    const uint32_t photons = 100;
    std::atomic<uint64_t> generated(0), at_doms(0);
    std::vector<std::thread> pool;
    for (int c = 0; c < clients; ++c)
        pool.emplace_back([&, c] {
            std::shared_ptr<I3CLSimClientInProcess> client = server.Connect();
            std::vector<std::size_t> sizes(rounds);
            for (int r = 0; r < rounds; ++r) {
                sizes[r] = bunch - 13 * r - c;
                client->EnqueueSteps(make_steps(sizes[r], photons, 7000 + 10 * c + r, 50 * c + r), uint32_t(r));
                generated += uint64_t(sizes[r]) * photons;
            }
            std::vector<int> seen(rounds, 0);
            for (int r = 0; r < rounds; ++r) {
                I3CLSimStepToPhotonConverter::ConversionResult_t res = client->GetConversionResult();
                CHECK(res.photons && res.identifier < uint32_t(rounds));
                if (!res.photons || res.identifier >= uint32_t(rounds)) continue;
                seen[res.identifier]++;
                at_doms += res.photons->size();
                bool mine = true; // the steps carried this client's tag in their own identifier field
                for (const I3CLSimPhoton &p : *res.photons) mine = mine && p.GetID() == 7000u + 10 * c + res.identifier;
                CHECK(mine);
            }
            bool once = true;
            for (int v : seen) once = once && v == 1;
            CHECK(once);
        });
    for (std::thread &t : pool) t.join();
    std::map<std::string, double> st = server.GetStatistics();
    double gen = 0, hits = 0, calls = 0;
    for (int i = 0; i < want; ++i) {
        const std::string post = want == 1 ? "" : "_" + std::to_string(i);
        gen += st["TotalNumPhotonsGenerated" + post];
        hits += st["TotalNumPhotonsAtDOMs" + post];
        calls += st["NumKernelCalls" + post];
        if (want > 1) CHECK(st["NumKernelCalls" + post] >= 1); // every device got work
    }
    CHECK(gen == double(generated.load()));
    CHECK(hits == double(at_doms.load()));
    CHECK(calls == clients * rounds);
    std::printf("%d converter(s) on %d device(s): %.0f photons generated, %.0f at DOMs, %.0f kernel calls\n", want, devices, gen, hits, calls);
}

int main(int argc, char **argv)
{
    try {
        test_with_dummy_converters();
        test_failing_converter_fails_the_clients();
        if (argc > 1 && std::string(argv[1]) == "--gpu") test_with_cuda_converters(argc > 2 ? std::atoi(argv[2]) : 2);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "unexpected exception: %s\n", e.what());
        return 2;
    }
    std::printf("%d checks, %d failed\n", g_checked, g_failed);
    return g_failed ? 1 : 0;
}
