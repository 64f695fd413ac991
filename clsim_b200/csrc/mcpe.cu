// mcpe.cu -- detected photon -> photo-electron (MCPE) on the device.
//
// Restates the innermost loops of the reference's two converters in the arithmetic they use (double precision
// on float photon records), one photon per thread:
//   * I3CLSimPhotonToMCPEConverterForDOMs::Convert  (private/clsim/dom/I3PhotonToMCPEConverter.cxx:602-669),
//     the in-loop converter of I3CLSimClientModule (…ClientModule.cxx:424-433);
//   * I3PhotonToMCPEConverter::Convert, per-photon part (…cxx:395-523), the stand-alone module.
// The photons it reads are the hit records the propagation kernel has just written, still in HBM: only the
// surviving photo-electrons (16 bytes each, about one photon in ten) cross PCIe.  The work is one pass over at
// most a few hundred thousand 80-byte records per bunch -- bandwidth-trivial, so the kernel is written for
// exactness (double precision, the reference's operation order), not for speed.
//
// Thinning draws come from MWC streams of their own (one per thread, multipliers from the safe-prime table):
// with T threads, photon j takes draw number j / T of stream j % T, so a checker can reproduce every decision.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/clsimcuda.h"
#include "mcpe.h"
#include "tables.h"

namespace clsimcu {
namespace {

constexpr int kMaxAcceptances = 8;
constexpr int kMaxAngularCoefficients = 24;
constexpr int kThreadsPerBlock = 256;

struct DevAcceptance {
    int kind, n;
    double x0, dx, value;
    const double *v;
};

struct DevMcpe {
    int flavour, num_acceptances, num_doms, num_angular, only_warn;
    DevAcceptance acceptances[kMaxAcceptances];
    double angular[kMaxAngularCoefficients];
    const uint32_t *dom_keys;      // sorted: (uint16) string id << 16 | om id
    const uint8_t *dom_acceptance; // same order
    const double *dom_efficiency;  // same order
    double dom_dir[3];
    double dom_radius, oversize, pancake;
};

#define CUDA_OK(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// I3CLSimFunctionFromTable::GetValue, equal spacing mode (private/clsim/function/I3CLSimFunctionFromTable.cxx:106-124),
// or I3CLSimFunctionConstant
__device__ double acceptance_at(const DevAcceptance &f, double wlen)
{
    if (f.kind == CLSIMCU_BIAS_CONSTANT) return f.value;
    double whole;
    double fraction = modf((wlen - f.x0) / f.dx, &whole);
    int bin = static_cast<int>(whole);
    if ((bin < 0) || ((bin == 0) && (fraction < 0))) {
        bin = 0;
        fraction = 0.;
    } else if (bin >= f.n - 1) {
        bin = f.n - 2;
        fraction = 1.;
    }
    const double lo = f.v[bin], hi = f.v[bin + 1];
    return lo + (hi - lo) * fraction;
}

// I3CLSimFunctionPolynomial::GetValue without a range (private/clsim/function/I3CLSimFunctionPolynomial.cxx:86-102)
__device__ double angular_at(const DevMcpe &c, double x)
{
    if (c.num_angular == 0) return 0.;
    double sum = c.angular[0], multiplier = 1.;
    for (int i = 1; i < c.num_angular; ++i) {
        multiplier *= x;
        sum += c.angular[i] * multiplier;
    }
    return sum;
}

__device__ int find_dom(const DevMcpe &c, uint32_t key)
{
    int lo = 0, hi = c.num_doms - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t k = c.dom_keys[mid];
        if (k == key) return mid;
        if (k < key) lo = mid + 1;
        else hi = mid - 1;
    }
    return -1;
}

__global__ void __launch_bounds__(kThreadsPerBlock)
photons_to_mcpe(const __grid_constant__ DevMcpe cfg, const __grid_constant__ McpeLaunch l, uint64_t *rng_x, const uint32_t *rng_a)
{
    const uint32_t n = l.count ? min(*l.count, l.max_count) : l.max_count;
    const uint32_t threads = gridDim.x * blockDim.x;
    const uint32_t me = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint64_t x = rng_x[me];
    const uint32_t a = rng_a[me];
    // whole warps walk the series together so that the output slots can be reserved per warp
    for (uint32_t base = me - lane; base < n; base += threads) {
        const uint32_t j = base + lane;
        bool survives = false;
        clsimcu_mcpe pe{};
        if (j < n) {
            const clsimcu_photon ph = l.photons[j];
            float u;
            if (l.uniforms) {
                u = l.uniforms[j];
            } else {
                // mwcrng_kernel.cl:12-28
                x = static_cast<uint64_t>(static_cast<uint32_t>(x)) * a + (x >> 32);
                u = __uint2float_rz(static_cast<uint32_t>(x)) * 2.3283064365386963e-10f;
            }
            double p = static_cast<double>(ph.weight);
            if (p < 0.) atomicAdd(l.counters + kMcpeNegativeWeight, 1u);
            if (p > 0.) {
                const int dom = find_dom(cfg, (static_cast<uint32_t>(static_cast<uint16_t>(ph.string_id)) << 16) | ph.om_id);
                if (dom < 0) {
                    atomicAdd(l.counters + kMcpeUnknownDom, 1u);
                } else {
                    const double rx = ph.x, ry = ph.y, rz = ph.z;   // photon position relative to the DOM centre
                    const double dist = sqrt(rx * rx + ry * ry + rz * rz);
                    double sin_t, cos_t, sin_p, cos_p;
                    sincos(static_cast<double>(ph.theta), &sin_t, &cos_t);
                    double cos_angle, time = static_cast<double>(ph.t);
                    if (cfg.flavour == CLSIMCU_MCPE_INLOOP) {
                        if (fabs(dist - 0.1651) > 0.03) atomicAdd(l.counters + kMcpeBadPosition, 1u);   // …cxx:610-622
                        cos_angle = -cos_t;                                                            // …cxx:624
                    } else {
                        sincos(static_cast<double>(ph.phi), &sin_p, &cos_p);
                        const double dx = sin_t * cos_p, dy = sin_t * sin_p, dz = cos_t;
                        cos_angle = -(dx * cfg.dom_dir[0] + dy * cfg.dom_dir[1] + dz * cfg.dom_dir[2]); // …cxx:408-410
                        if (cfg.pancake == 1. && fabs(dist - cfg.oversize * cfg.dom_radius) > 0.03 && !cfg.only_warn)
                            atomicAdd(l.counters + kMcpeBadPosition, 1u);                              // …cxx:416-451
                        // px = om.position - photon.pos = -r (…cxx:402-404); …cxx:512-517
                        const double dot = -(rx * dx + ry * dy + rz * dz);
                        time += dot * (1. - cfg.pancake / cfg.oversize) / static_cast<double>(ph.group_velocity);
                    }
                    cos_angle = fmax(-1., fmin(1., cos_angle));
                    p *= acceptance_at(cfg.acceptances[cfg.dom_acceptance[dom]], static_cast<double>(ph.wavelength));
                    p *= angular_at(cfg, cos_angle);
                    if (cfg.flavour == CLSIMCU_MCPE_MODULE) p *= cfg.dom_efficiency[dom];
                    if (p > 1.) atomicAdd(l.counters + kMcpeProbabilityAboveOne, 1u);
                    survives = !(p <= static_cast<double>(u));                                          // …cxx:502, 664
                    pe.string_id = ph.string_id;
                    pe.om_id = ph.om_id;
                    pe.time = static_cast<float>(time);
                    pe.npe = 1u;
                    pe.identifier = ph.identifier;
                }
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, survives);
        if (mask != 0u) {
            uint32_t first = 0;
            if (lane == __ffs(mask) - 1) first = atomicAdd(l.counters + kMcpeSurvivors, static_cast<uint32_t>(__popc(mask)));
            first = __shfl_sync(0xffffffffu, first, __ffs(mask) - 1);
            const uint32_t slot = first + __popc(mask & ((1u << lane) - 1u));
            if (survives && slot < l.cap) reinterpret_cast<uint4 *>(l.out)[slot] = *reinterpret_cast<const uint4 *>(&pe);
        }
    }
    if (!l.uniforms) rng_x[me] = x;
}

} // namespace
} // namespace clsimcu

using namespace clsimcu;

struct clsimcu_mcpe_converter {
    int device = 0;
    DevMcpe dev{};
    uint8_t *d_tables = nullptr;
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

int mcpe_device(const clsimcu_mcpe_converter *c) { return c->device; }

void mcpe_enqueue(clsimcu_mcpe_converter *c, const McpeLaunch &l, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lk(c->launch_mutex);
    if (c->last_use) CUDA_OK(cudaStreamWaitEvent(stream, c->last_use, 0));
    photons_to_mcpe<<<c->blocks, kThreadsPerBlock, 0, stream>>>(c->dev, l, c->d_rng_x, c->d_rng_a);
    CUDA_OK(cudaGetLastError());
    if (c->last_use) CUDA_OK(cudaEventRecord(c->last_use, stream));
}

std::string mcpe_error_text(const clsimcu_mcpe_converter *c, const uint32_t *k)
{
    // the reference's log_fatal texts (…cxx:396, 474-498, 606, 613, 630, 640)
    if (k[kMcpeNegativeWeight]) return "Photon with negative weight found. (" + std::to_string(k[kMcpeNegativeWeight]) + " photons)";
    if (k[kMcpeUnknownDom]) return "No wavelength acceptance configured for OMKey (" + std::to_string(k[kMcpeUnknownDom]) + " photons on DOMs outside the converter's DOM list)";
    if (k[kMcpeBadPosition])
        return "distance not " + std::to_string(c->dev.flavour == CLSIMCU_MCPE_INLOOP ? 165.1 : c->dev.oversize * c->dev.dom_radius * 1e3) +
               "mm (" + std::to_string(k[kMcpeBadPosition]) + " photons are more than 3 cm off the DOM surface)";
    if (k[kMcpeProbabilityAboveOne]) return "hitProbability > 1: your hit weights are too high. (" + std::to_string(k[kMcpeProbabilityAboveOne]) + " photons) cannot continue.";
    return std::string();
}

} // namespace clsimcu

namespace {

void free_converter(clsimcu_mcpe_converter *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaFree(c->d_tables);
    if (c->last_use) cudaEventDestroy(c->last_use);
    cudaFree(c->d_rng_x);
    cudaFree(c->d_rng_a);
    delete c;
}

} // namespace

extern "C" {

int clsimcu_mcpe_create(const clsimcu_mcpe_config *cfg, clsimcu_mcpe_converter **out)
{
    if (!cfg || !out) return report_error(CLSIMCU_ERR_INVALID, "config or output pointer is NULL");
    *out = nullptr;
    if (cfg->struct_size != static_cast<int32_t>(sizeof(clsimcu_mcpe_config)))
        return report_error(CLSIMCU_ERR_INVALID, "clsimcu_mcpe_config.struct_size does not match this library");
    if (cfg->flavour != CLSIMCU_MCPE_INLOOP && cfg->flavour != CLSIMCU_MCPE_MODULE) return report_error(CLSIMCU_ERR_INVALID, "unknown MCPE converter flavour");
    // Configure() of the module (…cxx:146-185)
    if (cfg->num_acceptances < 1 || !cfg->acceptances) return report_error(CLSIMCU_ERR_INVALID, "The \"WavelengthAcceptance\" parameter must not be empty.");
    if (cfg->num_acceptances > kMaxAcceptances) return report_error(CLSIMCU_ERR_UNSUPPORTED, "more than 8 wavelength acceptance curves are not supported");
    if (cfg->num_angular_coefficients < 1 || !cfg->angular_coefficients) return report_error(CLSIMCU_ERR_INVALID, "The \"AngularAcceptance\" parameter must not be empty.");
    if (cfg->num_angular_coefficients > kMaxAngularCoefficients) return report_error(CLSIMCU_ERR_UNSUPPORTED, "more than 24 polynomial coefficients are not supported");
    if (cfg->num_doms < 1 || !cfg->string_id || !cfg->dom_id) return report_error(CLSIMCU_ERR_INVALID, "the DOM list is empty");
    if (cfg->flavour == CLSIMCU_MCPE_MODULE && !(cfg->oversize_factor > 0. && cfg->pancake_factor > 0. && cfg->dom_radius > 0.))
        return report_error(CLSIMCU_ERR_INVALID, "DOMOversizeFactor, DOMPancakeFactor and DOMRadiusWithoutOversize must be positive");
    for (int i = 0; i < cfg->num_acceptances; ++i) {
        const clsimcu_wlen_bias &f = cfg->acceptances[i];
        if (f.kind == CLSIMCU_BIAS_TABLE && (f.n < 2 || !f.v || !(f.dx > 0.))) return report_error(CLSIMCU_ERR_INVALID, "wavelength acceptance table needs >= 2 entries and a positive spacing");
        if (f.kind != CLSIMCU_BIAS_TABLE && f.kind != CLSIMCU_BIAS_CONSTANT) return report_error(CLSIMCU_ERR_UNSUPPORTED, "wavelength acceptance must be I3CLSimFunctionFromTable (equal spacing) or I3CLSimFunctionConstant");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device || cfg->device < 0) {
        cudaGetLastError();
        return report_error(CLSIMCU_ERR_CUDA, "no usable CUDA device " + std::to_string(cfg->device) + " for the photon -> MCPE converter (there is no CPU fallback)");
    }
    clsimcu_mcpe_converter *c = new clsimcu_mcpe_converter;
    try {
        c->device = cfg->device;
        CUDA_OK(cudaSetDevice(c->device));
        CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        DevMcpe &d = c->dev;
        d.flavour = cfg->flavour;
        d.num_acceptances = cfg->num_acceptances;
        d.num_doms = cfg->num_doms;
        d.num_angular = cfg->num_angular_coefficients;
        d.only_warn = cfg->only_warn_about_positions;
        for (int i = 0; i < d.num_angular; ++i) d.angular[i] = cfg->angular_coefficients[i];
        for (int i = 0; i < 3; ++i) d.dom_dir[i] = cfg->dom_dir[i];
        d.dom_radius = cfg->dom_radius;
        d.oversize = cfg->oversize_factor;
        d.pancake = cfg->pancake_factor;
        // DOM list sorted by key
        std::vector<int> order(cfg->num_doms);
        for (int i = 0; i < cfg->num_doms; ++i) order[i] = i;
        auto key_of = [&](int i) {
            return (static_cast<uint32_t>(static_cast<uint16_t>(static_cast<int16_t>(cfg->string_id[i]))) << 16) | static_cast<uint32_t>(static_cast<uint16_t>(cfg->dom_id[i]));
        };
        std::sort(order.begin(), order.end(), [&](int a, int b) { return key_of(a) < key_of(b); });
        std::vector<uint32_t> keys(cfg->num_doms);
        std::vector<uint8_t> acc(cfg->num_doms);
        std::vector<double> eff(cfg->num_doms);
        for (int k = 0; k < cfg->num_doms; ++k) {
            const int i = order[k];
            if (cfg->string_id[i] < -32768 || cfg->string_id[i] > 32767 || cfg->dom_id[i] > 65535u) throw std::invalid_argument("string / DOM IDs do not fit the photon record (int16 / uint16)");
            keys[k] = key_of(i);
            if (k > 0 && keys[k] == keys[k - 1]) throw std::invalid_argument("duplicate OMKey in the DOM list");
            acc[k] = cfg->acceptance_of_dom ? cfg->acceptance_of_dom[i] : 0;
            if (acc[k] >= cfg->num_acceptances) throw std::invalid_argument("acceptance_of_dom entry out of range");
            eff[k] = cfg->efficiency_of_dom ? cfg->efficiency_of_dom[i] : 1.;
            if (!(eff[k] >= 0.)) throw std::invalid_argument("The relative DOM efficiency must not be < 0 or NaN (set a default)");
        }
        // one arena: doubles first (efficiencies, acceptance tables), then keys, then bytes
        size_t doubles = eff.size();
        for (int i = 0; i < cfg->num_acceptances; ++i)
            if (cfg->acceptances[i].kind == CLSIMCU_BIAS_TABLE) doubles += cfg->acceptances[i].n;
        const size_t off_keys = doubles * sizeof(double);
        const size_t off_acc = off_keys + keys.size() * sizeof(uint32_t);
        std::vector<uint8_t> arena(off_acc + acc.size());
        double *dd = reinterpret_cast<double *>(arena.data());
        std::copy(eff.begin(), eff.end(), dd);
        CUDA_OK(cudaMalloc(&c->d_tables, arena.size()));
        size_t at = eff.size();
        for (int i = 0; i < cfg->num_acceptances; ++i) {
            const clsimcu_wlen_bias &f = cfg->acceptances[i];
            DevAcceptance &a = d.acceptances[i];
            a.kind = f.kind;
            a.n = f.n;
            a.x0 = f.x0;
            a.dx = f.dx;
            a.value = f.value;
            a.v = nullptr;
            if (f.kind == CLSIMCU_BIAS_TABLE) {
                std::copy(f.v, f.v + f.n, dd + at);
                a.v = reinterpret_cast<const double *>(c->d_tables) + at;
                at += f.n;
            }
        }
        std::memcpy(arena.data() + off_keys, keys.data(), keys.size() * sizeof(uint32_t));
        std::memcpy(arena.data() + off_acc, acc.data(), acc.size());
        CUDA_OK(cudaMemcpy(c->d_tables, arena.data(), arena.size(), cudaMemcpyHostToDevice));
        d.dom_efficiency = reinterpret_cast<const double *>(c->d_tables);
        d.dom_keys = reinterpret_cast<const uint32_t *>(c->d_tables + off_keys);
        d.dom_acceptance = c->d_tables + off_acc;
        // MWC streams, one per thread of a one-wave grid
        int sms = 0;
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        c->blocks = sms;
        c->streams = static_cast<uint32_t>(sms) * kThreadsPerBlock;
        std::vector<uint32_t> a(c->streams);
        std::vector<uint64_t> x(c->streams);
        safeprime_multipliers(cfg->rng_first_multiplier, c->streams, a.data(), prime_cache_file());
        seed_rng_states(cfg->rng_seed, a.data(), x.data(), c->streams);
        CUDA_OK(cudaMalloc(&c->d_rng_x, c->streams * sizeof(uint64_t)));
        CUDA_OK(cudaMalloc(&c->d_rng_a, c->streams * sizeof(uint32_t)));
        CUDA_OK(cudaMemcpy(c->d_rng_x, x.data(), c->streams * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaEventCreateWithFlags(&c->last_use, cudaEventDisableTiming));
        CUDA_OK(cudaMemcpy(c->d_rng_a, a.data(), c->streams * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaDeviceSynchronize());   // (pageable sources on the default stream; the object's launches run on non-blocking streams)
    } catch (const std::invalid_argument &ex) {
        free_converter(c);
        return report_error(CLSIMCU_ERR_INVALID, ex.what());
    } catch (const std::exception &ex) {
        free_converter(c);
        return report_error(CLSIMCU_ERR_CUDA, ex.what());
    }
    *out = c;
    return CLSIMCU_OK;
}

int clsimcu_mcpe_destroy(clsimcu_mcpe_converter *c)
{
    if (!c) return report_error(CLSIMCU_ERR_INVALID, "converter is NULL");
    free_converter(c);
    return CLSIMCU_OK;
}

int clsimcu_mcpe_convert(clsimcu_mcpe_converter *c, const clsimcu_photon *photons, size_t n, const float *uniforms, clsimcu_mcpe *out, size_t cap,
                         size_t *n_out)
{
    if (!c) return report_error(CLSIMCU_ERR_STATE, "photon -> MCPE converter is not initialized!");
    if (!n_out || (n > 0 && !photons) || (cap > 0 && !out)) return report_error(CLSIMCU_ERR_INVALID, "NULL argument");
    if (n > 0xffffffffull || cap > 0xffffffffull) return report_error(CLSIMCU_ERR_INVALID, "series too long");
    *n_out = 0;
    if (n == 0) return CLSIMCU_OK;
    std::lock_guard<std::mutex> lk(c->mutex);
    clsimcu_photon *d_ph = nullptr;
    float *d_u = nullptr;
    clsimcu_mcpe *d_out = nullptr;
    uint32_t *d_k = nullptr;
    uint32_t k[kMcpeCounters] = {0};
    int rc = CLSIMCU_OK;
    try {
        CUDA_OK(cudaSetDevice(c->device));
        CUDA_OK(cudaMalloc(&d_ph, n * sizeof(clsimcu_photon)));
        CUDA_OK(cudaMalloc(&d_out, std::max<size_t>(cap, 1) * sizeof(clsimcu_mcpe)));
        CUDA_OK(cudaMalloc(&d_k, sizeof k));
        CUDA_OK(cudaMemcpyAsync(d_ph, photons, n * sizeof(clsimcu_photon), cudaMemcpyHostToDevice, c->stream));
        if (uniforms) {
            CUDA_OK(cudaMalloc(&d_u, n * sizeof(float)));
            CUDA_OK(cudaMemcpyAsync(d_u, uniforms, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        }
        CUDA_OK(cudaMemsetAsync(d_k, 0, sizeof k, c->stream));
        McpeLaunch l{d_ph, nullptr, static_cast<uint32_t>(n), d_u, d_out, static_cast<uint32_t>(cap), d_k};
        mcpe_enqueue(c, l, c->stream);
        CUDA_OK(cudaMemcpyAsync(k, d_k, sizeof k, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        const std::string bad = mcpe_error_text(c, k);
        if (!bad.empty()) {
            rc = report_error(CLSIMCU_ERR_INVALID, bad);
        } else {
            *n_out = k[kMcpeSurvivors];
            const size_t m = std::min<size_t>(k[kMcpeSurvivors], cap);
            if (m > 0) CUDA_OK(cudaMemcpy(out, d_out, m * sizeof(clsimcu_mcpe), cudaMemcpyDeviceToHost));
        }
    } catch (const std::exception &ex) {
        rc = report_error(CLSIMCU_ERR_CUDA, ex.what());
    }
    cudaFree(d_ph);
    cudaFree(d_u);
    cudaFree(d_out);
    cudaFree(d_k);
    return rc;
}

int clsimcu_mcpe_rng_get(clsimcu_mcpe_converter *c, uint64_t *x, uint32_t *a, size_t cap, size_t *streams)
{
    if (!c) return report_error(CLSIMCU_ERR_STATE, "photon -> MCPE converter is not initialized!");
    if (!streams) return report_error(CLSIMCU_ERR_INVALID, "NULL argument");
    *streams = c->streams;
    const size_t m = std::min<size_t>(cap, c->streams);
    if (m == 0) return CLSIMCU_OK;
    std::lock_guard<std::mutex> lk(c->mutex);
    try {
        CUDA_OK(cudaSetDevice(c->device));
        CUDA_OK(cudaDeviceSynchronize());
        if (x) CUDA_OK(cudaMemcpy(x, c->d_rng_x, m * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (a) CUDA_OK(cudaMemcpy(a, c->d_rng_a, m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    } catch (const std::exception &ex) {
        return report_error(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

} // extern "C"
