// stepgen.cu -- light source -> steps on the device.
//
// Restates MakeSteps of the reference's ppc-style parameterisation
// (private/clsim/I3CLSimLightSourceToStepConverterPPC.cxx:523-607, 785-842 and
// private/clsim/I3CLSimLightSourceToStepConverterUtils.h:63-198), one step per thread, in the double
// precision the reference uses before it rounds a step to its 48-byte float record.  What stays with the caller
// is what the reference does once per particle (EnqueueLightSource: light yield, Poisson draw, shower
// parameters); what runs here is the work per STEP -- millions of them for a bright event, 24 GB/s of records if
// they were made on the host at the rate the propagation kernel eats them.
//
// Random numbers: one MWC stream per generator thread (multipliers from the safe-prime table).  With T threads,
// step j is made by thread j % T and a thread makes its steps in ascending order, so a checker can replay every
// stream.  The reference draws from one converter stream plus four racing feeder threads (…PPC.cxx:689-711), so
// its own step sequence is not reproducible from a seed either; the distributions are what is specified.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/clsimcuda.h"
#include "mcpe.h"
#include "stepgen.h"
#include "tables.h"

namespace clsimcu {
namespace {

constexpr int kThreadsPerBlock = 256;
constexpr double kSpeedOfLight = 0.299792458;   // I3Constants::c [m/ns]
constexpr double kPi = 3.14159265358979323846;

struct DevStepGen {
    double one_over_a, b, big_i;   // angular smearing (…PPC.cxx:684-686)
};

#define CUDA_OK(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct Mwc64 {
    uint64_t x;
    uint32_t a;
    // mwcRngRandomNumber_co / _oc (…Utils.h:63-72)
    __device__ double co()
    {
        x = (x & 0xffffffffull) * a + (x >> 32);
        return static_cast<double>(static_cast<uint32_t>(x & 0xffffffffull)) / 4294967296.0;
    }
    __device__ double oc() { return 1.0 - co(); }
};

// gammaDistributedNumber (…Utils.h:78-110), "stolen from PPC": Weibull below shape 1, Cheng above, with the
// reference's float temporaries
__device__ double gamma_distributed(double shape, Mwc64 &rng)
{
    double x;
    if (shape < 1.) {
        const double c = 1. / shape;
        const double d = (1. - shape) * pow(shape, shape / (1. - shape));
        double z, e;
        do {
            z = -log(rng.oc());
            e = -log(rng.oc());
            x = pow(z, c);
        } while (z + e < d + x);
    } else {
        const double b = shape - log(4.0);
        const double l = sqrt(2. * shape - 1.0);
        const double cheng = 1.0 + log(4.5);
        float y, z, r;
        do {
            const double rx = rng.oc();
            const double ry = rng.oc();
            y = static_cast<float>(log(ry / (1. - ry)) / l);
            // y and z are floats in the reference, so its std::exp(y) and std::log(z) are the single-precision overloads
            // (pinned against the reference's own function: tests/test_stepgen_oracle.py).  Taken here as the double
            // function rounded to float, i.e. the correctly rounded value, which glibc's expf / logf return in all but
            // about one call in a thousand (CUDA's expf / logf are 1-2 ulp functions).
            x = shape * static_cast<double>(static_cast<float>(exp(static_cast<double>(y))));
            z = static_cast<float>(rx * ry * ry);
            r = static_cast<float>(b + (shape + l) * static_cast<double>(y) - x);
        } while (static_cast<double>(r) < 4.5 * static_cast<double>(z) - cheng &&
                 static_cast<double>(r) < static_cast<double>(static_cast<float>(log(static_cast<double>(z)))));
    }
    return x;
}

// scatterDirectionByAngle (…Utils.h:153-196)
__device__ void scatter_direction(double cosa, double sina, double &x, double &y, double &z, double random_value)
{
    const double b = 2.0 * kPi * random_value;
    const double cosb = cos(b), sinb = sin(b);
    const double sinth = sqrt(fmax(0., 1. - z * z));
    if (sinth > 0.) {
        const double ox = x, oy = y, oz = z;
        x = ox * cosa - (oy * cosb + oz * ox * sinb) * sina / sinth;
        y = oy * cosa + (ox * cosb - oz * oy * sinb) * sina / sinth;
        z = oz * cosa + sina * sinb * sinth;
    } else {
        x = sina * cosb;
        y = sina * sinb;
        z = (z >= 0.) ? cosa : -cosa;
    }
    const double recip = 1. / sqrt(x * x + y * y + z * z);
    x *= recip; y *= recip; z *= recip;
}

// I3CLSimStep::SetDir(x, y, z) (public/clsim/I3CLSimStep.h:128-133) goes through I3Direction (dataclasses,
// un-vendored): zenith/azimuth of where the particle comes FROM, theta = pi - zenith, phi = pi + azimuth
__device__ void set_dir(clsimcu_step &s, double x, double y, double z)
{
    const double r = sqrt(x * x + y * y + z * z);
    const double zenith = acos(fmax(-1., fmin(1., -z / r)));
    double azimuth = atan2(-y / r, -x / r);
    if (azimuth < 0.) azimuth += 2. * kPi;
    double phi = kPi + azimuth;
    if (phi >= 2. * kPi) phi -= 2. * kPi;
    s.theta = static_cast<float>(kPi - zenith);
    s.phi = static_cast<float>(phi);
}

__global__ void __launch_bounds__(kThreadsPerBlock)
make_steps(const __grid_constant__ DevStepGen cfg, const __grid_constant__ StepGenLaunch l, uint64_t *rng_x, const uint32_t *rng_a)
{
    const uint32_t threads = gridDim.x * blockDim.x;
    const uint32_t me = blockIdx.x * blockDim.x + threadIdx.x;
    Mwc64 rng{rng_x[me], rng_a[me]};
    for (uint64_t j = me; j < l.total; j += threads) {
        // the queue entry this step belongs to: the last one whose first step is <= j
        uint32_t lo = 0, hi = l.num_sources - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (l.first_step[mid] <= j) lo = mid;
            else hi = mid - 1;
        }
        const clsimcu_step_source src = l.sources[lo];
        const uint64_t local = j - l.first_step[lo];
        clsimcu_step s;
        s.num_photons = (local < src.num_steps) ? src.photons_per_step : src.photons_in_last_step;   // …PPC.cxx:586-597
        s.weight = 1.f;
        s.beta = 1.f;
        s.identifier = src.identifier;
        s.source_type = 0;   // Cherenkov emission
        s.dummy1 = 0;
        s.dummy2 = 0;
        double dx = src.dir_x, dy = src.dir_y, dz = src.dir_z;
        if (src.kind == CLSIMCU_SOURCE_TRACK_MUON_LIKE) {
            // GenerateStepForMuon (…PPC.cxx:821-842)
            s.x = static_cast<float>(src.x); s.y = static_cast<float>(src.y); s.z = static_cast<float>(src.z);
            s.t = static_cast<float>(src.t);
            s.length = static_cast<float>(src.length);
        } else {
            // FillStep (…PPC.cxx:523-556): where along the axis
            const double along = (src.kind == CLSIMCU_SOURCE_CASCADE) ? src.pb * gamma_distributed(src.pa, rng) : rng.co() * src.length;
            // FeederThread (…PPC.cxx:752-757): how far off the axis
            const double cos_val = fmax(1. - pow(-log(1. - rng.co() * cfg.big_i) / cfg.b, cfg.one_over_a), -1.);
            const double sin_val = sqrt(1. - cos_val * cos_val);
            const double random_value = rng.co();
            // GenerateStep (…PPC.cxx:785-819)
            s.x = static_cast<float>(src.x + along * dx);
            s.y = static_cast<float>(src.y + along * dy);
            s.z = static_cast<float>(src.z + along * dz);
            s.t = static_cast<float>(src.t + along / kSpeedOfLight);
            s.length = static_cast<float>(0.001);
            scatter_direction(cos_val, sin_val, dx, dy, dz, random_value);
        }
        set_dir(s, dx, dy, dz);
        // 48 bytes as three 16-byte stores
        const uint4 *w = reinterpret_cast<const uint4 *>(&s);
        uint4 *dst = reinterpret_cast<uint4 *>(l.out + j);
        dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2];
    }
    rng_x[me] = rng.x;
}

} // namespace
} // namespace clsimcu

using namespace clsimcu;

struct clsimcu_step_generator {
    int device = 0;
    DevStepGen dev{};
    uint64_t *d_rng_x = nullptr;
    uint32_t *d_rng_a = nullptr;
    uint32_t streams = 0;
    int blocks = 0;
    cudaStream_t stream = nullptr;
    std::mutex mutex;
    // every launch reads and writes the object's MWC rows: launches from different streams (the object's own, and the
    // compute stream of the engine it is attached to) are chained through this event, under launch_mutex
    std::mutex launch_mutex;
    cudaEvent_t last_use = nullptr;
};

namespace clsimcu {

int stepgen_device(const clsimcu_step_generator *g) { return g->device; }

void stepgen_enqueue(clsimcu_step_generator *g, const StepGenLaunch &l, cudaStream_t stream)
{
    if (l.total == 0) return;
    std::lock_guard<std::mutex> lk(g->launch_mutex);
    if (g->last_use) CUDA_OK(cudaStreamWaitEvent(stream, g->last_use, 0));
    make_steps<<<g->blocks, kThreadsPerBlock, 0, stream>>>(g->dev, l, g->d_rng_x, g->d_rng_a);
    CUDA_OK(cudaGetLastError());
    if (g->last_use) CUDA_OK(cudaEventRecord(g->last_use, stream));
}

// Validates n queue entries and lays their steps out: first_step[i] = index of entry i's first step,
// first_step[n] = total; *photons = sum of all photons.  Returns an error text, or empty.
std::string stepgen_layout(const clsimcu_step_source *sources, size_t n, std::vector<uint64_t> &first_step, uint64_t *photons)
{
    first_step.assign(n + 1, 0);
    uint64_t total = 0, ph = 0;
    for (size_t i = 0; i < n; ++i) {
        const clsimcu_step_source &s = sources[i];
        if (s.kind != CLSIMCU_SOURCE_CASCADE && s.kind != CLSIMCU_SOURCE_TRACK_CASCADE_LIKE && s.kind != CLSIMCU_SOURCE_TRACK_MUON_LIKE)
            return "unknown step source kind in entry " + std::to_string(i);
        if (s.kind == CLSIMCU_SOURCE_CASCADE && !(s.pa > 0.) ) return "cascade entry " + std::to_string(i) + " needs a positive gamma shape";
        if (s.kind != CLSIMCU_SOURCE_CASCADE && !(s.length > 0.))
            return "Found a cascade segment with length " + std::to_string(s.length) + ". This should not be.";   // …PPC.cxx:338-340
        const double norm = s.dir_x * s.dir_x + s.dir_y * s.dir_y + s.dir_z * s.dir_z;
        if (!(std::fabs(norm - 1.) < 1e-6)) return "direction of entry " + std::to_string(i) + " is not a unit vector";
        if (s.num_steps > 0 && s.photons_per_step == 0) return "entry " + std::to_string(i) + " has steps of zero photons";
        first_step[i] = total;
        total += s.num_steps + (s.photons_in_last_step > 0 ? 1u : 0u);
        ph += s.num_steps * s.photons_per_step + s.photons_in_last_step;
        if (total > 0xffffffffull) return "more than 2^32 steps in one call";
    }
    first_step[n] = total;
    if (photons) *photons = ph;
    return std::string();
}

} // namespace clsimcu

namespace {

void free_generator(clsimcu_step_generator *g)
{
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamDestroy(g->stream);
    if (g->last_use) cudaEventDestroy(g->last_use);
    cudaFree(g->d_rng_x);
    cudaFree(g->d_rng_a);
    delete g;
}

} // namespace

extern "C" {

int clsimcu_stepgen_create(const clsimcu_step_generator_config *cfg, clsimcu_step_generator **out)
{
    if (!cfg || !out) return report_error(CLSIMCU_ERR_INVALID, "config or output pointer is NULL");
    *out = nullptr;
    if (cfg->struct_size != static_cast<int32_t>(sizeof(clsimcu_step_generator_config)))
        return report_error(CLSIMCU_ERR_INVALID, "clsimcu_step_generator_config.struct_size does not match this library");
    if (!(cfg->angular_a > 0.) || !(cfg->angular_b > 0.)) return report_error(CLSIMCU_ERR_INVALID, "angular smearing parameters must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device || cfg->device < 0) {
        cudaGetLastError();
        return report_error(CLSIMCU_ERR_CUDA, "no usable CUDA device " + std::to_string(cfg->device) + " for the step generator (there is no CPU fallback)");
    }
    clsimcu_step_generator *g = new clsimcu_step_generator;
    try {
        g->device = cfg->device;
        CUDA_OK(cudaSetDevice(g->device));
        CUDA_OK(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
        g->dev.one_over_a = 1. / cfg->angular_a;
        g->dev.b = cfg->angular_b;
        g->dev.big_i = 1. - std::exp(-cfg->angular_b * std::pow(2., cfg->angular_a));
        int sms = 0;
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device));
        g->blocks = 2 * sms;
        g->streams = static_cast<uint32_t>(g->blocks) * kThreadsPerBlock;
        std::vector<uint32_t> a(g->streams);
        std::vector<uint64_t> x(g->streams);
        safeprime_multipliers(cfg->rng_first_multiplier, g->streams, a.data(), prime_cache_file());
        seed_rng_states(cfg->rng_seed, a.data(), x.data(), g->streams);
        CUDA_OK(cudaMalloc(&g->d_rng_x, g->streams * sizeof(uint64_t)));
        CUDA_OK(cudaMalloc(&g->d_rng_a, g->streams * sizeof(uint32_t)));
        CUDA_OK(cudaMemcpy(g->d_rng_x, x.data(), g->streams * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaEventCreateWithFlags(&g->last_use, cudaEventDisableTiming));
        CUDA_OK(cudaMemcpy(g->d_rng_a, a.data(), g->streams * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaDeviceSynchronize());   // (pageable sources on the default stream; the object's launches run on non-blocking streams)
    } catch (const std::exception &ex) {
        free_generator(g);
        return report_error(CLSIMCU_ERR_CUDA, ex.what());
    }
    *out = g;
    return CLSIMCU_OK;
}

int clsimcu_stepgen_destroy(clsimcu_step_generator *g)
{
    if (!g) return report_error(CLSIMCU_ERR_INVALID, "generator is NULL");
    free_generator(g);
    return CLSIMCU_OK;
}

int clsimcu_stepgen_generate(clsimcu_step_generator *g, const clsimcu_step_source *sources, size_t n, clsimcu_step *out, size_t cap, size_t *n_out)
{
    if (!g) return report_error(CLSIMCU_ERR_STATE, "I3CLSimLightSourceToStepConverterPPC is not initialized!");
    if (!n_out || (n > 0 && !sources) || (cap > 0 && !out)) return report_error(CLSIMCU_ERR_INVALID, "NULL argument");
    *n_out = 0;
    if (n == 0) return CLSIMCU_OK;
    if (n > 0x7fffffffull) return report_error(CLSIMCU_ERR_INVALID, "too many queue entries");
    std::vector<uint64_t> first;
    const std::string bad = stepgen_layout(sources, n, first, nullptr);
    if (!bad.empty()) return report_error(CLSIMCU_ERR_INVALID, bad);
    const uint64_t total = first[n];
    *n_out = total;
    if (total == 0) return CLSIMCU_OK;
    std::lock_guard<std::mutex> lk(g->mutex);
    clsimcu_step_source *d_src = nullptr;
    uint64_t *d_first = nullptr;
    clsimcu_step *d_out = nullptr;
    int rc = CLSIMCU_OK;
    try {
        CUDA_OK(cudaSetDevice(g->device));
        CUDA_OK(cudaMalloc(&d_src, n * sizeof(clsimcu_step_source)));
        CUDA_OK(cudaMalloc(&d_first, (n + 1) * sizeof(uint64_t)));
        CUDA_OK(cudaMalloc(&d_out, total * sizeof(clsimcu_step)));
        CUDA_OK(cudaMemcpyAsync(d_src, sources, n * sizeof(clsimcu_step_source), cudaMemcpyHostToDevice, g->stream));
        CUDA_OK(cudaMemcpyAsync(d_first, first.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, g->stream));
        StepGenLaunch l{d_src, d_first, static_cast<uint32_t>(n), total, d_out};
        stepgen_enqueue(g, l, g->stream);
        CUDA_OK(cudaStreamSynchronize(g->stream));
        const size_t m = std::min<size_t>(total, cap);
        if (m > 0) CUDA_OK(cudaMemcpy(out, d_out, m * sizeof(clsimcu_step), cudaMemcpyDeviceToHost));
    } catch (const std::exception &ex) {
        rc = report_error(CLSIMCU_ERR_CUDA, ex.what());
    }
    cudaFree(d_src);
    cudaFree(d_first);
    cudaFree(d_out);
    return rc;
}

int clsimcu_stepgen_rng_get(clsimcu_step_generator *g, uint64_t *x, uint32_t *a, size_t cap, size_t *streams)
{
    if (!g) return report_error(CLSIMCU_ERR_STATE, "I3CLSimLightSourceToStepConverterPPC is not initialized!");
    if (!streams) return report_error(CLSIMCU_ERR_INVALID, "NULL argument");
    *streams = g->streams;
    const size_t m = std::min<size_t>(cap, g->streams);
    if (m == 0) return CLSIMCU_OK;
    std::lock_guard<std::mutex> lk(g->mutex);
    try {
        CUDA_OK(cudaSetDevice(g->device));
        CUDA_OK(cudaDeviceSynchronize());
        if (x) CUDA_OK(cudaMemcpy(x, g->d_rng_x, m * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (a) CUDA_OK(cudaMemcpy(a, g->d_rng_a, m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    } catch (const std::exception &ex) {
        return report_error(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

} // extern "C"
