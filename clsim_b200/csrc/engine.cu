// engine.cu -- host runtime of libclsimcuda and the C ABI of include/clsimcuda.h.
//
// Replaces the host driver of the reference (private/opencl/I3CLSimStepToPhotonConverterOpenCL.cxx:
// Initialize :217-388, worker thread :1142-1315, upload :778-902, download :994-1086,
// statistics :1088-1140, 1621-1640) with a CUDA-runtime design for one B200:
//
//   caller -> bounded queue(5) -> submit thread: stage into pinned memory, async H2D on the
//   slot's transfer stream, kernel on the single compute stream (kernels serialise, so the
//   per-thread RNG streams are never shared by two launches), async D2H of the counters
//   -> in-flight queue -> drain thread: waits for the counters, copies exactly the hits
//   that were produced, hands a result to the unbounded output queue.
//
// With double buffering three slots rotate, so bunch k+1 uploads and bunch k-1 downloads
// while bunch k computes; without it one slot gives the reference's strictly serial order.
// There is no CPU fallback: without a usable device clsimcu_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <limits>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/clsimcuda.h"
#include "device_scene.h"
#include "mcpe.h"
#include "stepgen.h"
#include "tabulate.h"
#include "tables.h"

namespace clsimcu {
namespace {

thread_local std::string t_last_error;

int fail(int code, const std::string &msg)
{
    t_last_error = msg;
    return code;
}

// Large host-to-host copies (a 48 MiB bunch into pinned staging, 18 MB of hits out of the pinned mirror) are on the
// latency path of the first and the last bunch of a run: split them over a few threads.
void parallel_copy(void *dst, const void *src, size_t bytes)
{
    constexpr size_t kPerThread = size_t(4) << 20;
    const unsigned parts = static_cast<unsigned>(std::min<size_t>(4, bytes / kPerThread));
    if (parts < 2) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t chunk = ((bytes / parts) + 63) & ~size_t(63);
    std::vector<std::thread> pool;
    for (unsigned p = 1; p < parts; ++p) {
        const size_t at = p * chunk, len = (p + 1 == parts) ? bytes - at : chunk;
        pool.emplace_back([=] { std::memcpy(static_cast<char *>(dst) + at, static_cast<const char *>(src) + at, len); });
    }
    std::memcpy(dst, src, chunk);
    for (std::thread &t : pool) t.join();
}

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

// launch_*_kernel: 0 ok, -1 bunch too large for the fast kernel's step index, -2 unsupported option, -3 CUDA error
// (left in place by the launcher: cudaPeekAtLastError, so that its text can be read here)
static std::string launch_error_text(int rc)
{
    if (rc == -1) return "kernel launch failed: more steps in one bunch than the fast kernel's birth tag can index";
    if (rc == -2) return "kernel launch failed: option not supported by this kernel (photon history longer than 32 entries)";
    return std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError());
}

#define CUDA_OK(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e__));                            \
    } while (0)

// Blocking queue with optional bound, the role of I3CLSimQueue (public/clsim/I3CLSimQueue.h:58-171).
template <class T> class BlockingQueue {
public:
    explicit BlockingQueue(size_t bound) : bound_(bound) {}
    bool put(T v)
    {
        std::unique_lock<std::mutex> lk(m_);
        not_full_.wait(lk, [&] { return closed_ || bound_ == 0 || q_.size() < bound_; });
        if (closed_) return false;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
        return true;
    }
    bool try_get(T &out)
    {
        std::lock_guard<std::mutex> lk(m_);
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    bool get(T &out)
    {
        std::unique_lock<std::mutex> lk(m_);
        not_empty_.wait(lk, [&] { return closed_ || !q_.empty(); });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    size_t size() const
    {
        std::lock_guard<std::mutex> lk(m_);
        return q_.size();
    }
    bool empty() const { return size() == 0; }
    void close()
    {
        std::lock_guard<std::mutex> lk(m_);
        closed_ = true;
        not_empty_.notify_all();
        not_full_.notify_all();
    }

private:
    mutable std::mutex m_;
    std::condition_variable not_empty_, not_full_;
    std::deque<T> q_;
    size_t bound_;
    bool closed_ = false;
};

// A bunch on its way to the device: the caller's steps were copied ONCE, straight into a pinned staging
// buffer (index `staging` of the engine's pool), from where the copy engine takes them.
struct Bunch {
    uint32_t identifier = 0;
    int staging = -1;
    size_t num_steps = 0;
    uint64_t generated = 0;
    // steps still to be made on the device (clsimcu_enqueue_sources): the staging buffer holds num_sources queue
    // entries followed by num_sources + 1 first-step indices
    clsimcu_step_generator *generator = nullptr;
    size_t num_sources = 0;
};

struct HostResult {
    uint32_t identifier = 0;
    std::unique_ptr<clsimcu_photon[]> photons;   // not a vector: no zero-fill of 18 MB before it is overwritten
    size_t num_photons = 0;
    std::vector<float> history;
    std::vector<clsimcu_mcpe> mcpes;
    uint64_t generated = 0, counted = 0;
};

struct Slot {
    int index = 0;
    cudaStream_t xfer = nullptr;
    cudaEvent_t uploaded = nullptr, k_start = nullptr, k_stop = nullptr, counted = nullptr, converted = nullptr;
    clsimcu_step *d_steps = nullptr;
    uint8_t *d_sources = nullptr;                                   // queue entries + first-step indices of a generated bunch
    int staging = -1;                                               // pinned staging buffer the bunch came in
    clsimcu_photon *d_photons = nullptr, *h_photons = nullptr;
    float *d_history = nullptr, *h_history = nullptr;
    uint32_t *d_counters = nullptr, *h_counters = nullptr;           // [0] hits, [1] work
    unsigned long long *d_stats = nullptr, *h_stats = nullptr;      // [0] photons, [1] segments
    clsimcu_mcpe *d_mcpes = nullptr, *h_mcpes = nullptr;            // photon -> MCPE conversion attached
    uint32_t *d_mcpe_counters = nullptr, *h_mcpe_counters = nullptr;
    uint32_t identifier = 0;
    size_t num_steps = 0;
    uint64_t generated = 0;
};

// Collects table arrays into one arena; pointers are fixed up after the upload.
class Arena {
public:
    template <class T> size_t add(const std::vector<T> &v)
    {
        const size_t at = (bytes_.size() + 15) & ~size_t(15);
        bytes_.resize(at + std::max<size_t>(16, v.size() * sizeof(T)), 0);
        if (!v.empty()) std::memcpy(bytes_.data() + at, v.data(), v.size() * sizeof(T));
        return at;
    }
    const std::vector<uint8_t> &bytes() const { return bytes_; }

private:
    std::vector<uint8_t> bytes_;
};

constexpr size_t kL2FlushBytes = size_t(256) << 20;

template <class T> const T *at(const uint8_t *base, size_t off) { return reinterpret_cast<const T *>(base + off); }

} // namespace
} // namespace clsimcu

using namespace clsimcu;

struct clsimcu_engine {
    int device = 0;
    int kernel_mode = CLSIMCU_KERNEL_FAST;
    size_t max_items = 0, granularity = 1, max_hits = 0;
    int history_entries = 0;
    bool save_all = false;
    SceneTables tables;
    DevScene scene;
    uint8_t *d_arena = nullptr;
    DevScene *d_scene = nullptr;   // copy of `scene` in global memory for out-of-line device code
    uint64_t *d_rng_x = nullptr;
    uint32_t *d_rng_a = nullptr;
    size_t rng_n = 0;
    int fast_blocks = 0, fast_threads = 0;
    cudaStream_t compute = nullptr;
    std::vector<Slot> slots;
    BlockingQueue<Bunch> inbox{5};                 // queueToOpenCL_ depth 5 (…OpenCL.cxx:77)
    BlockingQueue<int> free_slots{0}, in_flight{0};
    // pinned staging buffers for incoming bunches: as many as can be waiting (5) or in a slot
    clsimcu_step *staging[16] = {nullptr};
    size_t staging_count = 0, staging_max = 0;
    BlockingQueue<int> free_staging{0};
    BlockingQueue<std::shared_ptr<HostResult>> outbox{0};
    std::thread submit_thread, drain_thread;
    std::atomic<bool> stopping{false};
    std::atomic<bool> enqueued_any{false};
    clsimcu_mcpe_converter *mcpe = nullptr;        // clsimcu_attach_mcpe_converter
    bool mcpe_keep_photons = false;
    std::mutex error_mutex;
    std::string async_error;
    // statistics (…OpenCL.h:377-381)
    std::mutex stats_mutex;
    double device_ns = 0., host_ns = 0.;
    uint64_t kernel_calls = 0, photons_generated = 0, photons_at_doms = 0;
    std::chrono::steady_clock::time_point last_stamp;
    // resident path
    std::mutex compute_mutex;
    clsimcu_step *d_res_steps = nullptr;
    clsimcu_photon *d_res_photons = nullptr;
    uint32_t *d_res_counters = nullptr, *h_res_counters = nullptr;
    unsigned long long *d_res_stats = nullptr, *h_res_stats = nullptr;
    float *d_res_history = nullptr;                // resident path: raw history rings of the hits, [res_cap][entries][4]
    float *d_history_ring = nullptr;               // fast kernel with photon history: the lanes' scatter-point rings
    uint64_t *d_tag_x = nullptr;
    uint32_t *d_tag_a = nullptr;
    uint8_t *d_l2_flush = nullptr;   // written between timed launches (larger than the 126 MB L2)
    clsimcu_step *d_tab_steps = nullptr;   // table-maker variant: the bunch being tabulated
    size_t res_steps = 0, res_cap = 0;
    uint64_t res_generated_per_run = 0;
    uint32_t res_last_hits = 0;

    void set_async_error(const std::string &m)
    {
        std::lock_guard<std::mutex> lk(error_mutex);
        if (async_error.empty()) async_error = m;
    }
    bool check_async_error(std::string &m)
    {
        std::lock_guard<std::mutex> lk(error_mutex);
        m = async_error;
        return !async_error.empty();
    }
    // An error in one of the two worker threads ends the engine for good (the reference's threads log_fatal).  Everybody
    // who waits on one of the engine's queues is woken up with a failure: a caller blocked in EnqueueSteps for a staging
    // buffer, the submit thread waiting for a slot the dead drain thread would have freed, the drain thread waiting for a
    // launch the dead submit thread would have made.  (Results that were finished before can still be fetched.)
    void abort_pipeline(const std::string &m)
    {
        set_async_error(m);
        inbox.close();
        free_slots.close();
        free_staging.close();
        in_flight.close();
    }
};

namespace clsimcu {
namespace {

void upload_tables(clsimcu_engine &e, int near_pixel_budget)
{
    const SceneTables &t = e.tables;
    Arena arena;
    DevScene &s = e.scene;
    std::memset(&s, 0, sizeof s);

    // ---- medium
    const MediumTables &m = t.medium;
    DevMedium &dm = s.medium;
    dm.num_layers = m.num_layers;
    dm.z0 = m.z0; dm.h = m.h; dm.inv_h = 1.f / m.h;
    dm.kappa = m.kappa; dm.A = m.A; dm.B = m.B; dm.D = m.D; dm.E = m.E; dm.alpha = m.alpha;
    dm.inv_ref_wlen = m.inv_ref_wlen;
    for (int i = 0; i < 5; ++i) { dm.n_phase[i] = m.n_phase[i]; dm.n_group[i] = m.n_group[i]; }
    dm.c_light = m.c_light;
    dm.scat_kind = m.scat_kind;
    dm.f_sl = m.f_sl; dm.one_minus_f_sl = m.one_minus_f_sl; dm.g = m.g; dm.g2 = m.g2; dm.sl_beta = m.sl_beta;
    dm.inv_f_sl = (m.f_sl > 0.f) ? 1.f / m.f_sl : 0.f;
    dm.inv_one_minus_f_sl = (m.one_minus_f_sl > 0.f) ? 1.f / m.one_minus_f_sl : 0.f;
    dm.inv_2g = (m.g != 0.f) ? 1.f / (2.f * m.g) : 0.f;
    if (m.scat_kind == CLSIMCU_SCAT_MIXED_SL_HG && m.f_sl > 0.f && m.one_minus_f_sl > 0.f && m.g != 0.f) {
        const double f = m.f_sl, omf = m.one_minus_f_sl, g = m.g;
        dm.sl_off = static_cast<float>(static_cast<double>(m.sl_beta) * (std::log2(1.0 / f) - 32.0) + 1.0);
        dm.hg_h0 = static_cast<float>(1.0 + g * (2.0 / omf - 1.0));
        dm.hg_h1 = static_cast<float>(-2.0 * g / omf / 4294967296.0);
        dm.mix_split = m.f_sl * 4294967296.f;
        dm.hg_c = static_cast<float>((1.0 + g * g) / (2.0 * g));
        dm.hg_w = static_cast<float>((1.0 - g * g) * (1.0 - g * g) / (2.0 * g));
        dm.mix_folded = 1;
    }
    dm.tilt_nd = m.tilt_nd; dm.tilt_nz = m.tilt_nz;
    dm.tilt_z0 = m.tilt_z0; dm.tilt_dz = m.tilt_dz; dm.tilt_lnx = m.tilt_lnx; dm.tilt_lny = m.tilt_lny;
    dm.tilt_inv_dz = (m.tilt_nd > 0 && m.tilt_dz != 0.f) ? 1.f / m.tilt_dz : 0.f;
    dm.tilt_zr_offset = (m.tilt_nd > 0 && m.tilt_dz != 0.f) ? static_cast<float>(-static_cast<double>(m.tilt_z0) / static_cast<double>(m.tilt_dz)) : 0.f;
    dm.tilt_lut_n = 0; dm.tilt_lut_scale = dm.tilt_lut_offset = 0.f;
    if (m.tilt_nd >= 4) {
        // interior nodes dist[1 .. nd-2]; cell 0 is everything below dist[1], cell c >= 1 starts at dist[1] + (c - 1) * width
        double gap = std::numeric_limits<double>::infinity();
        for (int i = 2; i <= m.tilt_nd - 2; ++i) gap = std::min(gap, static_cast<double>(m.tilt_dist[i]) - m.tilt_dist[i - 1]);
        const double span = static_cast<double>(m.tilt_dist[m.tilt_nd - 2]) - m.tilt_dist[1];
        if (gap > 0. && std::isfinite(gap)) {
            const double width = gap * 0.99;   // below the smallest gap, with room for the margin the kernel widens a cell by: never two nodes in a cell
            const int cells = static_cast<int>(std::ceil(span / width)) + 2;
            if (cells <= kTiltLutMaxCells) {
                dm.tilt_lut_n = cells;
                dm.tilt_lut_scale = static_cast<float>(1. / width);
                dm.tilt_lut_offset = static_cast<float>(1. - m.tilt_dist[1] / width);
            }
        }
    }
    dm.anisotropy = m.anisotropy; dm.pre_renorm = m.pre_renorm; dm.post_renorm = m.post_renorm;
    for (int i = 0; i < 3; ++i) { dm.l[i] = m.l[i]; dm.rl[i] = m.rl[i]; }
    dm.azx = m.azx; dm.azy = m.azy; dm.neg_azy = m.neg_azy; dm.B2 = m.B2;
    for (int i = 0; i < 9; ++i) { dm.pre[i] = m.pre[i]; dm.post[i] = m.post[i]; }
    std::vector<float> abs_dust(m.num_layers), abs_tau(m.num_layers);
    for (int i = 0; i < m.num_layers; ++i) {
        abs_dust[i] = m.D * m.a_dust400[i] + m.E;
        abs_tau[i] = 1.f + 0.01f * m.delta_tau[i];
    }
    const size_t o_adust = arena.add(m.a_dust400), o_dtau = arena.add(m.delta_tau), o_b400 = arena.add(m.b400);
    const size_t o_absd = arena.add(abs_dust), o_abst = arena.add(abs_tau);
    const size_t o_tdist = arena.add(m.tilt_dist), o_tcorr = arena.add(m.tilt_corr);

    // ---- wavelength generators and bias
    if (t.generators.size() > static_cast<size_t>(kMaxWlenGenerators))
        throw std::runtime_error("at most " + std::to_string(kMaxWlenGenerators) + " wavelength generators are supported");
    s.num_generators = static_cast<int>(t.generators.size());
    size_t o_gx[kMaxWlenGenerators], o_gd[kMaxWlenGenerators], o_gc[kMaxWlenGenerators];
    for (int i = 0; i < s.num_generators; ++i) {
        const WlenGeneratorTable &g = t.generators[i];
        DevWlenGenerator &dg = s.generators[i];
        dg.kind = g.kind; dg.n = g.n; dg.x0 = g.x0; dg.dx = g.dx;
        dg.min_val = g.min_val; dg.range = g.range; dg.value = g.value;
        o_gx[i] = arena.add(g.xs);
        o_gd[i] = arena.add(g.density);
        o_gc[i] = arena.add(g.cumulative);
    }
    s.bias.kind = t.bias.kind; s.bias.n = t.bias.n; s.bias.x0 = t.bias.x0; s.bias.dx = t.bias.dx; s.bias.value = t.bias.value;
    const size_t o_bias = arena.add(t.bias.v);

    // ---- geometry
    const GeometryTables &g = t.geometry;
    DevGeometry &dg = s.geo;
    size_t o_grid[kMaxSubdetectors] = {0};
    size_t o_sx = 0, o_sy = 0, o_smin = 0, o_smax = 0, o_sset = 0, o_lcount = 0, o_lstart = 0, o_lheight = 0, o_l2d = 0;
    size_t o_tdx = 0, o_tdy = 0, o_tz = 0, o_tstart = 0, o_mx = 0, o_my = 0, o_sid = 0, o_doff = 0, o_dids = 0;
    size_t o_near_info = 0;
    if (t.has_geometry) {
        dg.num_strings = g.num_strings; dg.num_sets = g.num_sets; dg.max_layers = g.max_layers;
        dg.num_grids = static_cast<int>(g.grids.size());
        dg.layer_table_size = static_cast<int>(g.layer_to_dom.size());
        dg.om_radius = g.om_radius; dg.string_max_radius = g.string_max_radius;
        dg.tmpl_scale_x = g.tmpl_scale_x; dg.tmpl_scale_y = g.tmpl_scale_y;
        for (int i = 0; i < dg.num_grids; ++i) {
            const CellGridTable &c = g.grids[i];
            DevCellGrid &dc = dg.grids[i];
            dc.num_x = c.num_x; dc.num_y = c.num_y;
            dc.start_x = c.start_x; dc.start_y = c.start_y; dc.width_x = c.width_x; dc.width_y = c.width_y;
            dc.inv_width_x = 1.f / c.width_x; dc.inv_width_y = 1.f / c.width_y;
            o_grid[i] = arena.add(c.cell_to_string);
        }
        o_sx = arena.add(g.string_x); o_sy = arena.add(g.string_y);
        o_smin = arena.add(g.string_min_z); o_smax = arena.add(g.string_max_z);
        o_sset = arena.add(g.string_set);
        o_lcount = arena.add(g.set_layer_count); o_lstart = arena.add(g.set_start_z); o_lheight = arena.add(g.set_layer_height);
        o_l2d = arena.add(g.layer_to_dom);
        o_tdx = arena.add(g.tmpl_dx); o_tdy = arena.add(g.tmpl_dy); o_tz = arena.add(g.tmpl_z);
        o_tstart = arena.add(g.string_tmpl_start);
        o_mx = arena.add(g.string_mean_x); o_my = arena.add(g.string_mean_y);
        // index -> ID tables with the reference's range checks (…OpenCL.cxx:1579-1590)
        std::vector<int16_t> sid(g.num_strings);
        std::vector<uint32_t> doff(g.num_strings);
        std::vector<uint16_t> dids;
        for (int i = 0; i < g.num_strings; ++i) {
            const int id = g.string_index_to_id[i];
            if (id < -32768 || id > 32767)
                throw std::runtime_error("Your detector I3Geometry uses a string ID \"" + std::to_string(id) + "\". Large IDs like that are currently not supported by clsim.");
            sid[i] = static_cast<int16_t>(id);
            doff[i] = static_cast<uint32_t>(dids.size());
            for (uint32_t d : g.dom_index_to_id[i]) {
                if (d > 65535u)
                    throw std::runtime_error("Your detector I3Geometry uses a OM ID \"" + std::to_string(d) + "\". Large IDs like that are currently not supported by clsim.");
                dids.push_back(static_cast<uint16_t>(d));
            }
        }
        o_sid = arena.add(sid); o_doff = arena.add(doff); o_dids = arena.add(dids);

        // xy pixel map for the fast kernel (see device_scene.h); built for the pixel budget the
        // caller found to fit into shared memory
        {
            const CollisionMap map = build_collision_map(g, near_pixel_budget);
            dg.near_x0 = map.x0; dg.near_y0 = map.y0; dg.near_inv_pixel = map.inv_pixel;
            dg.near_off_x = map.off_x; dg.near_off_y = map.off_y;
            dg.near_nx = map.nx; dg.near_ny = map.ny;
            o_near_info = arena.add(map.info);
        }
    }

    s.stop_detected = t.stop_detected; s.save_all = t.save_all; s.fixed_abs = t.fixed_abs; s.pancake = t.pancake;
    s.history_entries = t.history_entries;
    s.prescale = t.prescale; s.fixed_abs_lens = t.fixed_abs_lens; s.pancake_factor = t.pancake_factor;
    s.inv_pancake_factor = t.pancake ? 1.f / t.pancake_factor : 1.f;
    s.generic_transforms = std::getenv("CLSIMCU_GENERIC_TRANSFORMS") ? 1 : 0;   // (here, once, in the creating thread: not per launch)

    if (e.d_arena) { cudaFree(e.d_arena); e.d_arena = nullptr; }
    if (e.d_scene) { cudaFree(e.d_scene); e.d_scene = nullptr; }
    CUDA_OK(cudaMalloc(&e.d_arena, arena.bytes().size()));
    CUDA_OK(cudaMemcpy(e.d_arena, arena.bytes().data(), arena.bytes().size(), cudaMemcpyHostToDevice));
    const uint8_t *b = e.d_arena;
    dm.a_dust400 = at<float>(b, o_adust); dm.delta_tau = at<float>(b, o_dtau); dm.b400 = at<float>(b, o_b400);
    dm.abs_dust = at<float>(b, o_absd); dm.abs_tau = at<float>(b, o_abst);
    dm.tilt_dist = at<float>(b, o_tdist); dm.tilt_corr = at<float>(b, o_tcorr);
    for (int i = 0; i < s.num_generators; ++i) {
        s.generators[i].xs = at<float>(b, o_gx[i]);
        s.generators[i].density = at<float>(b, o_gd[i]);
        s.generators[i].cumulative = at<float>(b, o_gc[i]);
    }
    s.bias.v = at<float>(b, o_bias);
    if (t.has_geometry) {
        for (int i = 0; i < dg.num_grids; ++i) dg.grids[i].cell_to_string = at<uint16_t>(b, o_grid[i]);
        dg.string_x = at<float>(b, o_sx); dg.string_y = at<float>(b, o_sy);
        dg.string_min_z = at<float>(b, o_smin); dg.string_max_z = at<float>(b, o_smax);
        dg.string_set = at<uint8_t>(b, o_sset);
        dg.set_layer_count = at<uint16_t>(b, o_lcount);
        dg.set_start_z = at<float>(b, o_lstart); dg.set_layer_height = at<float>(b, o_lheight);
        dg.layer_to_dom = at<uint16_t>(b, o_l2d);
        dg.tmpl_dx = at<int16_t>(b, o_tdx); dg.tmpl_dy = at<int16_t>(b, o_tdy); dg.tmpl_z = at<float>(b, o_tz);
        dg.string_tmpl_start = at<uint32_t>(b, o_tstart);
        dg.string_mean_x = at<float>(b, o_mx); dg.string_mean_y = at<float>(b, o_my);
        dg.string_index_to_id = at<int16_t>(b, o_sid);
        dg.dom_id_offset = at<uint32_t>(b, o_doff);
        dg.dom_ids = at<uint16_t>(b, o_dids);
        dg.near_info = at<uint32_t>(b, o_near_info);
    }
    CUDA_OK(cudaMalloc(&e.d_scene, sizeof(DevScene)));
    CUDA_OK(cudaMemcpy(e.d_scene, &s, sizeof(DevScene), cudaMemcpyHostToDevice));
}

} // namespace

int report_error(int code, const std::string &msg) { return fail(code, msg); }
std::string prime_cache_path();
std::string prime_cache_file() { return prime_cache_path(); }

std::string prime_cache_path()
{
    if (const char *p = std::getenv("CLSIMCU_SAFEPRIMES_CACHE")) return p;
    // next to the shared library: <dir of libclsimcuda.so>/data/safeprimes_base32.bin
    Dl_info info;
    if (dladdr(reinterpret_cast<const void *>(&prime_cache_path), &info) && info.dli_fname) {
        std::string path(info.dli_fname);
        const size_t slash = path.rfind('/');
        if (slash != std::string::npos) return path.substr(0, slash) + "/data/safeprimes_base32.bin";
    }
    return std::string();
}

namespace {

void launch(clsimcu_engine &e, const LaunchArgs &args, cudaStream_t stream)
{
    int rc;
    if (e.kernel_mode == CLSIMCU_KERNEL_REFERENCE) rc = launch_reference_kernel(e.scene, args, stream);
    else rc = launch_fast_kernel(e.scene, args, e.fast_blocks, stream);
    if (rc != 0) throw CudaError(launch_error_text(rc));
}

void submit_loop(clsimcu_engine *e)
{
    try {
        CUDA_OK(cudaSetDevice(e->device));
        Bunch bunch;
        while (e->inbox.get(bunch)) {
            int si;
            if (!e->free_slots.get(si)) break;
            Slot &s = e->slots[si];
            s.identifier = bunch.identifier;
            s.num_steps = bunch.num_steps;
            s.generated = bunch.generated;
            s.staging = bunch.staging;
            const size_t source_bytes = (bunch.num_sources * sizeof(clsimcu_step_source) + 15) & ~size_t(15);
            if (bunch.generator) {
                if (!s.d_sources) CUDA_OK(cudaMalloc(&s.d_sources, kMaxSourcesPerBunch * (sizeof(clsimcu_step_source) + sizeof(uint64_t)) + 64));
                CUDA_OK(cudaMemcpyAsync(s.d_sources, e->staging[bunch.staging], source_bytes + (bunch.num_sources + 1) * sizeof(uint64_t),
                                        cudaMemcpyHostToDevice, s.xfer));
            } else {
                CUDA_OK(cudaMemcpyAsync(s.d_steps, e->staging[bunch.staging], s.num_steps * sizeof(clsimcu_step), cudaMemcpyHostToDevice, s.xfer));
            }
            CUDA_OK(cudaMemsetAsync(s.d_counters, 0, 2 * sizeof(uint32_t), s.xfer));
            CUDA_OK(cudaMemsetAsync(s.d_stats, 0, 8 * sizeof(unsigned long long), s.xfer));
            CUDA_OK(cudaEventRecord(s.uploaded, s.xfer));
            {
                std::lock_guard<std::mutex> lk(e->compute_mutex);
                CUDA_OK(cudaStreamWaitEvent(e->compute, s.uploaded, 0));
                if (bunch.generator) {
                    // the bunch is made where it is consumed: 48 bytes per step never cross PCIe
                    StepGenLaunch g{reinterpret_cast<const clsimcu_step_source *>(s.d_sources), reinterpret_cast<const uint64_t *>(s.d_sources + source_bytes),
                                    static_cast<uint32_t>(bunch.num_sources), s.num_steps, s.d_steps};
                    stepgen_enqueue(bunch.generator, g, e->compute);
                }
                CUDA_OK(cudaEventRecord(s.k_start, e->compute));
                LaunchArgs a{};
                a.steps = s.d_steps;
                a.num_steps = static_cast<uint32_t>(s.num_steps);
                a.max_hits = static_cast<uint32_t>(e->max_hits);
                a.photons = s.d_photons;
                a.history = s.d_history;
                a.history_ring = e->d_history_ring;
                a.hit_counter = s.d_counters;
                a.work_counter = s.d_counters + 1;
                a.stats = s.d_stats;
                a.rng_x = e->d_rng_x;
                a.rng_a = e->d_rng_a;
                a.count_stats = 0;
                a.scene_dev = e->d_scene;
                a.rng_creation_offset = static_cast<uint32_t>(e->fast_blocks) * e->fast_threads;
                launch(*e, a, e->compute);
                CUDA_OK(cudaEventRecord(s.k_stop, e->compute));
                if (e->mcpe) {
                    // the hits never leave HBM as photons: thin them to photo-electrons right behind the kernel
                    CUDA_OK(cudaMemsetAsync(s.d_mcpe_counters, 0, kMcpeCounters * sizeof(uint32_t), e->compute));
                    McpeLaunch l{s.d_photons, s.d_counters, static_cast<uint32_t>(e->max_hits), nullptr, s.d_mcpes, static_cast<uint32_t>(e->max_hits),
                                 s.d_mcpe_counters};
                    mcpe_enqueue(e->mcpe, l, e->compute);
                    CUDA_OK(cudaEventRecord(s.converted, e->compute));
                }
            }
            CUDA_OK(cudaStreamWaitEvent(s.xfer, e->mcpe ? s.converted : s.k_stop, 0));
            if (e->mcpe) CUDA_OK(cudaMemcpyAsync(s.h_mcpe_counters, s.d_mcpe_counters, kMcpeCounters * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.xfer));
            CUDA_OK(cudaMemcpyAsync(s.h_counters, s.d_counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.xfer));
            CUDA_OK(cudaEventRecord(s.counted, s.xfer));
            if (!e->in_flight.put(si)) break;
        }
    } catch (const std::exception &ex) {
        e->abort_pipeline(ex.what());
    }
    e->in_flight.close();
}

// Device ring layout -> forward order, unused rows NaN (…OpenCL.cxx:940-989)
void unroll_history(const float *raw, const clsimcu_photon *photons, size_t n, int entries, std::vector<float> &out)
{
    out.assign(n * entries * 4, std::nanf(""));
    for (size_t i = 0; i < n; ++i) {
        const uint32_t scat = photons[i].num_scatters;
        if (scat == 0) continue;
        const uint32_t recorded = std::min<uint32_t>(scat, static_cast<uint32_t>(entries));
        uint32_t at = (scat <= static_cast<uint32_t>(entries)) ? 0u : scat % entries;
        for (uint32_t j = 0; j < recorded; ++j) {
            std::memcpy(&out[(i * entries + j) * 4], &raw[(i * entries + at) * 4], 4 * sizeof(float));
            if (++at >= static_cast<uint32_t>(entries)) at = 0;
        }
    }
}

void drain_loop(clsimcu_engine *e)
{
    try {
        CUDA_OK(cudaSetDevice(e->device));
        int si;
        while (e->in_flight.get(si)) {
            Slot &s = e->slots[si];
            CUDA_OK(cudaEventSynchronize(s.counted));
            e->free_staging.put(s.staging);   // the upload is long done
            s.staging = -1;
            const auto now = std::chrono::steady_clock::now();
            float k_ms = 0.f;
            CUDA_OK(cudaEventElapsedTime(&k_ms, s.k_start, s.k_stop));
            const uint32_t counted = s.h_counters[0];
            const size_t n = std::min<size_t>(counted, e->max_hits);
            if (counted > e->max_hits)
                std::fprintf(stderr, "clsimcuda: Maximum number of photons exceeded, only receiving %zu of %u photons\n", e->max_hits, counted);
            auto res = std::make_shared<HostResult>();
            res->identifier = s.identifier;
            res->generated = s.generated;
            res->counted = counted;
            size_t n_pe = 0;
            if (e->mcpe) {
                const std::string bad = mcpe_error_text(e->mcpe, s.h_mcpe_counters);
                if (!bad.empty()) throw std::runtime_error(bad);
                n_pe = std::min<size_t>(s.h_mcpe_counters[kMcpeSurvivors], e->max_hits);
                if (n_pe > 0) {
                    CUDA_OK(cudaMemcpyAsync(s.h_mcpes, s.d_mcpes, n_pe * sizeof(clsimcu_mcpe), cudaMemcpyDeviceToHost, s.xfer));
                    CUDA_OK(cudaStreamSynchronize(s.xfer));
                    res->mcpes.assign(s.h_mcpes, s.h_mcpes + n_pe);
                }
            }
            if (n > 0 && (!e->mcpe || e->mcpe_keep_photons)) {
                CUDA_OK(cudaMemcpyAsync(s.h_photons, s.d_photons, n * sizeof(clsimcu_photon), cudaMemcpyDeviceToHost, s.xfer));
                if (e->history_entries > 0)
                    CUDA_OK(cudaMemcpyAsync(s.h_history, s.d_history, n * e->history_entries * 4 * sizeof(float), cudaMemcpyDeviceToHost, s.xfer));
                CUDA_OK(cudaStreamSynchronize(s.xfer));
                res->photons.reset(new clsimcu_photon[n]);
                res->num_photons = n;
                parallel_copy(res->photons.get(), s.h_photons, n * sizeof(clsimcu_photon));
                if (e->history_entries > 0) unroll_history(s.h_history, s.h_photons, n, e->history_entries, res->history);
            }
            {
                std::lock_guard<std::mutex> lk(e->stats_mutex);
                e->device_ns += static_cast<double>(k_ms) * 1e6;
                e->host_ns += std::chrono::duration<double, std::nano>(now - e->last_stamp).count();
                e->last_stamp = now;
                e->kernel_calls += 1;
                e->photons_generated += s.generated;
                e->photons_at_doms += counted;
            }
            e->free_slots.put(si);
            e->outbox.put(res);
        }
    } catch (const std::exception &ex) {
        e->abort_pipeline(ex.what());
    }
    e->outbox.close();
}

void free_engine(clsimcu_engine *e)
{
    cudaSetDevice(e->device);
    for (Slot &s : e->slots) {
        if (s.xfer) cudaStreamSynchronize(s.xfer);
        cudaFree(s.d_steps); cudaFree(s.d_sources); cudaFree(s.d_photons); cudaFree(s.d_history); cudaFree(s.d_counters); cudaFree(s.d_stats);
        cudaFreeHost(s.h_photons); cudaFreeHost(s.h_history); cudaFreeHost(s.h_counters); cudaFreeHost(s.h_stats);
        cudaFree(s.d_mcpes); cudaFree(s.d_mcpe_counters); cudaFreeHost(s.h_mcpes); cudaFreeHost(s.h_mcpe_counters);
        if (s.converted) cudaEventDestroy(s.converted);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
        if (s.k_start) cudaEventDestroy(s.k_start);
        if (s.k_stop) cudaEventDestroy(s.k_stop);
        if (s.counted) cudaEventDestroy(s.counted);
        if (s.xfer) cudaStreamDestroy(s.xfer);
    }
    for (size_t i = 0; i < e->staging_count; ++i) cudaFreeHost(e->staging[i]);
    cudaFree(e->d_history_ring); cudaFree(e->d_res_history);
    cudaFree(e->d_arena); cudaFree(e->d_scene); cudaFree(e->d_rng_x); cudaFree(e->d_rng_a);
    cudaFree(e->d_res_steps); cudaFree(e->d_res_photons); cudaFree(e->d_res_counters); cudaFree(e->d_res_stats);
    cudaFree(e->d_tag_x); cudaFree(e->d_tag_a); cudaFree(e->d_l2_flush); cudaFree(e->d_tab_steps);
    cudaFreeHost(e->h_res_counters); cudaFreeHost(e->h_res_stats);
    if (e->compute) cudaStreamDestroy(e->compute);
    delete e;
}

} // namespace
} // namespace clsimcu

// ---- seam for the table-maker variant (tabulate.cu) -----------------------------------------------------
namespace clsimcu {

int engine_device(const clsimcu_engine *e) { return e->device; }
size_t engine_max_items(const clsimcu_engine *e) { return e->max_items; }

std::string engine_launch_tabulate(clsimcu_engine *e, const clsimcu_step *steps, size_t n, const TabulateArgs *d_tab)
{
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        if (!e->d_tab_steps) CUDA_OK(cudaMalloc(&e->d_tab_steps, e->max_items * sizeof(clsimcu_step)));
        // pageable source: the copy has returned from the host buffer when this call returns; stream order keeps
        // the previous launch from being overwritten
        CUDA_OK(cudaMemcpyAsync(e->d_tab_steps, steps, n * sizeof(clsimcu_step), cudaMemcpyHostToDevice, e->compute));
        LaunchArgs a{};
        a.steps = e->d_tab_steps;
        a.num_steps = static_cast<uint32_t>(n);
        a.rng_x = e->d_rng_x;
        a.rng_a = e->d_rng_a;
        a.scene_dev = e->d_scene;
        a.tabulate = d_tab;
        if (e->kernel_mode == CLSIMCU_KERNEL_FAST) {
            // the persistent kernel's bookkeeping: work counter (zeroed per launch), creation streams behind the
            // propagation streams; table mode records no photons, the hit counter stays untouched
            Slot &s = e->slots[0];
            CUDA_OK(cudaMemsetAsync(s.d_counters, 0, 2 * sizeof(uint32_t), e->compute));
            a.hit_counter = s.d_counters;
            a.work_counter = s.d_counters + 1;
            a.stats = s.d_stats;
            a.count_stats = 0;
            a.rng_creation_offset = static_cast<uint32_t>(e->fast_blocks) * e->fast_threads;
            if (const int rc = launch_fast_kernel(e->scene, a, e->fast_blocks, e->compute)) throw CudaError(launch_error_text(rc));
        } else if (const int rc = launch_reference_kernel(e->scene, a, e->compute)) {
            throw CudaError(launch_error_text(rc));
        }
    } catch (const std::exception &ex) {
        return ex.what();
    }
    return std::string();
}

std::string engine_copy_on_stream(clsimcu_engine *e, void *dst, const void *src, size_t bytes, bool to_device, bool wait)
{
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        if (bytes) CUDA_OK(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, e->compute));
        if (wait) CUDA_OK(cudaStreamSynchronize(e->compute));
    } catch (const std::exception &ex) {
        return ex.what();
    }
    return std::string();
}

} // namespace clsimcu

extern "C" {

int clsimcu_device_count(int *count)
{
    if (!count) return fail(CLSIMCU_ERR_INVALID, "count pointer is NULL");
    int n = 0;
    const cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) {
        *count = 0;
        return fail(CLSIMCU_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); libclsimcuda has no CPU fallback");
    }
    *count = n;
    return CLSIMCU_OK;
}

const char *clsimcu_last_error(void) { return t_last_error.c_str(); }
const char *clsimcu_version(void) { return "clsimcuda 0.1 (sm_100a)"; }
size_t clsimcu_sizeof_config(void) { return sizeof(clsimcu_config); }

int clsimcu_create(const clsimcu_config *config, clsimcu_engine **out)
{
    if (!config || !out) return fail(CLSIMCU_ERR_INVALID, "config or engine pointer is NULL");
    *out = nullptr;
    if (config->struct_size != static_cast<int32_t>(sizeof(clsimcu_config)))
        return fail(CLSIMCU_ERR_INVALID, "clsimcu_config.struct_size does not match this library");
    if (config->kernel_mode != CLSIMCU_KERNEL_FAST && config->kernel_mode != CLSIMCU_KERNEL_REFERENCE)
        return fail(CLSIMCU_ERR_INVALID, "unknown kernel_mode");
    clsimcu_engine *e = new clsimcu_engine();
    try {
        build_scene_tables(*config, e->tables);
    } catch (const std::exception &ex) {
        delete e;
        return fail(CLSIMCU_ERR_INVALID, ex.what());
    }
    try {
        int count = 0;
        cudaError_t ce = cudaGetDeviceCount(&count);
        if (ce != cudaSuccess || count == 0)
            throw CudaError(std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); libclsimcuda has no CPU fallback");
        if (config->device < 0 || config->device >= count) throw CudaError("CUDA device ordinal " + std::to_string(config->device) + " does not exist");
        e->device = config->device;
        CUDA_OK(cudaSetDevice(e->device));
        e->kernel_mode = config->kernel_mode;
        e->history_entries = config->photon_history_entries;
        e->save_all = config->save_all_photons != 0;
        e->granularity = config->workgroup_size ? config->workgroup_size : 1;
        e->max_items = config->max_num_workitems ? config->max_num_workitems : 10240; // class default (…OpenCL.cxx:94)
        if (e->max_items % e->granularity != 0)
            throw std::runtime_error("The maximum number of work items (" + std::to_string(e->max_items) + ") must be a multiple of the workgroup size (" + std::to_string(e->granularity) + ").");
        // output capacity (…OpenCL.cxx:266-277)
        size_t per_item = config->output_photons_per_workitem ? config->output_photons_per_workitem : 10;
        if (e->save_all && !config->output_photons_per_workitem) {
            per_item = static_cast<size_t>(10000. * config->save_all_photons_prescale);
            if (per_item < 1) per_item = 1;
        }
        e->max_hits = std::min<size_t>(e->max_items * per_item, 0xffffffffull);
        if (!e->save_all && e->max_hits < 1000) e->max_hits = 1000;

        if (e->kernel_mode == CLSIMCU_KERNEL_FAST) {
            // the pixel map of the collision test takes whatever shared memory the other tables leave
            const char *why = nullptr;
            int budget = 40000;
            if (const char *env = std::getenv("CLSIMCU_PIXEL_BUDGET")) budget = std::max(512, std::atoi(env)); // tuning knob
            for (;;) {
                upload_tables(*e, budget);
                if (fast_kernel_supports(e->scene, &why)) break;
                if (budget <= 512 || !fast_kernel_smem_is_the_problem(e->scene))
                    throw std::runtime_error(std::string("the fast kernel does not support this configuration (") + why + "); use CLSIMCU_KERNEL_REFERENCE");
                budget = budget * 7 / 8;
            }
            fast_kernel_geometry(e->device, &e->fast_blocks, &e->fast_threads);
            if (e->history_entries > 0) {
                const size_t ring_bytes = static_cast<size_t>(e->history_entries) * e->fast_blocks * e->fast_threads * 4 * sizeof(float);
                CUDA_OK(cudaMalloc(&e->d_history_ring, ring_bytes));
                CUDA_OK(cudaMemset(e->d_history_ring, 0, ring_bytes));
            }
            if (e->max_items > (size_t(1) << kFastKernelStepIndexBits))
                throw std::runtime_error("the fast kernel takes bunches of at most 2^27 steps (max_num_workitems is " + std::to_string(e->max_items) + ")");
        } else {
            upload_tables(*e, 4096);
        }

        // RNG streams: one per work item (reference order) or one per resident thread (fast)
        // (fast: two per resident thread, one for photon creation and one for propagation)
        size_t need = (e->kernel_mode == CLSIMCU_KERNEL_REFERENCE) ? e->max_items : 2 * static_cast<size_t>(e->fast_blocks) * e->fast_threads;
        e->rng_n = config->rng_n ? static_cast<size_t>(config->rng_n) : need;
        if (e->rng_n < need)
            throw std::runtime_error("rng_n (" + std::to_string(e->rng_n) + ") is smaller than the number of RNG streams this configuration needs (" + std::to_string(need) + ")");
        std::vector<uint32_t> a(e->rng_n);
        std::vector<uint64_t> x(e->rng_n);
        if (config->rng_a) {
            if (!config->rng_x) throw std::runtime_error("rng_a given without rng_x");
            std::memcpy(a.data(), config->rng_a, e->rng_n * sizeof(uint32_t));
            std::memcpy(x.data(), config->rng_x, e->rng_n * sizeof(uint64_t));
        } else {
            safeprime_multipliers(config->rng_first_multiplier, e->rng_n, a.data(), prime_cache_path());
            seed_rng_states(config->rng_seed, a.data(), x.data(), e->rng_n);
        }
        CUDA_OK(cudaMalloc(&e->d_rng_x, e->rng_n * sizeof(uint64_t)));
        CUDA_OK(cudaMalloc(&e->d_rng_a, e->rng_n * sizeof(uint32_t)));
        CUDA_OK(cudaMemcpy(e->d_rng_x, x.data(), e->rng_n * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(e->d_rng_a, a.data(), e->rng_n * sizeof(uint32_t), cudaMemcpyHostToDevice));

        CUDA_OK(cudaStreamCreateWithFlags(&e->compute, cudaStreamNonBlocking));
        const int nslots = config->enable_double_buffering ? 3 : 1;
        e->slots.resize(nslots);
        for (int i = 0; i < nslots; ++i) {
            Slot &s = e->slots[i];
            s.index = i;
            CUDA_OK(cudaStreamCreateWithFlags(&s.xfer, cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreate(&s.uploaded));
            CUDA_OK(cudaEventCreate(&s.k_start));
            CUDA_OK(cudaEventCreate(&s.k_stop));
            CUDA_OK(cudaEventCreate(&s.counted));
            CUDA_OK(cudaEventCreate(&s.converted));
            CUDA_OK(cudaMalloc(&s.d_steps, e->max_items * sizeof(clsimcu_step)));
            CUDA_OK(cudaMalloc(&s.d_photons, e->max_hits * sizeof(clsimcu_photon)));
            CUDA_OK(cudaMalloc(&s.d_counters, 2 * sizeof(uint32_t)));
            CUDA_OK(cudaMalloc(&s.d_stats, 8 * sizeof(unsigned long long)));
            CUDA_OK(cudaHostAlloc(&s.h_photons, e->max_hits * sizeof(clsimcu_photon), cudaHostAllocDefault));
            CUDA_OK(cudaHostAlloc(&s.h_counters, 2 * sizeof(uint32_t), cudaHostAllocDefault));
            CUDA_OK(cudaHostAlloc(&s.h_stats, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
            if (e->history_entries > 0) {
                const size_t hb = e->max_hits * e->history_entries * 4 * sizeof(float);
                CUDA_OK(cudaMalloc(&s.d_history, hb));
                CUDA_OK(cudaHostAlloc(&s.h_history, hb, cudaHostAllocDefault));
            }
            e->free_slots.put(i);
        }
        // every staging buffer up front (5 bunches can wait in the inbox, one per slot is in flight): pinning 48 MiB
        // takes tens of milliseconds and stalls the device -- not something to do between two bunches
        e->staging_max = 5 + static_cast<size_t>(nslots);
        for (size_t i = 0; i < e->staging_max; ++i) {
            CUDA_OK(cudaHostAlloc(&e->staging[i], e->max_items * sizeof(clsimcu_step), cudaHostAllocDefault));
            e->staging_count = i + 1;
            e->free_staging.put(static_cast<int>(i));
        }
        // the tables, the scene and the RNG rows went up from pageable memory on the default stream, which the engine's
        // streams do not synchronise with: everything has arrived before the first launch can be made
        CUDA_OK(cudaDeviceSynchronize());
        e->last_stamp = std::chrono::steady_clock::now();
        e->submit_thread = std::thread(submit_loop, e);
        e->drain_thread = std::thread(drain_loop, e);
    } catch (const CudaError &ex) {
        const std::string msg = ex.what();
        free_engine(e);
        return fail(CLSIMCU_ERR_CUDA, msg);
    } catch (const std::exception &ex) {
        const std::string msg = ex.what();
        free_engine(e);
        return fail(CLSIMCU_ERR_INVALID, msg);
    }
    *out = e;
    return CLSIMCU_OK;
}

int clsimcu_destroy(clsimcu_engine *e)
{
    if (!e) return fail(CLSIMCU_ERR_INVALID, "engine is NULL");
    e->stopping = true;
    e->inbox.close();
    e->free_slots.close();
    e->free_staging.close();
    if (e->submit_thread.joinable()) e->submit_thread.join();
    e->in_flight.close();
    if (e->drain_thread.joinable()) e->drain_thread.join();
    e->outbox.close();
    free_engine(e);
    return CLSIMCU_OK;
}

// A wait on one of the engine's queues ended without an item: a worker thread died of an error (its text is the
// answer), or the engine is being destroyed.
static int interrupted(clsimcu_engine *e)
{
    std::string err;
    if (e->check_async_error(err)) return fail(CLSIMCU_ERR_CUDA, err);
    return fail(CLSIMCU_ERR_INTERRUPTED, "engine is shutting down");
}

// A pinned staging buffer for an incoming bunch; waits for one to come back when all are in use.
static int acquire_staging(clsimcu_engine *e, int &index)
{
    if (!e->free_staging.get(index)) return interrupted(e);
    return CLSIMCU_OK;
}

int clsimcu_enqueue_sources(clsimcu_engine *e, clsimcu_step_generator *g, const clsimcu_step_source *sources, size_t n, uint32_t identifier)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "I3CLSimStepToPhotonConverterCUDA is not initialized!");
    if (!g) return fail(CLSIMCU_ERR_STATE, "I3CLSimLightSourceToStepConverterPPC is not initialized!");
    std::string err;
    if (e->check_async_error(err)) return fail(CLSIMCU_ERR_CUDA, err);
    if (stepgen_device(g) != e->device) return fail(CLSIMCU_ERR_INVALID, "the step generator lives on another device than the engine");
    if (!sources || n == 0) return fail(CLSIMCU_ERR_INVALID, "Steps are empty!");
    if (n > kMaxSourcesPerBunch) return fail(CLSIMCU_ERR_INVALID, "more than 65536 step sources in one bunch");
    std::vector<uint64_t> first;
    uint64_t photons = 0;
    const std::string bad = stepgen_layout(sources, n, first, &photons);
    if (!bad.empty()) return fail(CLSIMCU_ERR_INVALID, bad);
    const uint64_t total = first[n];
    // the preconditions of EnqueueSteps (…OpenCL.cxx:1527-1540) on the bunch that will exist on the device
    if (total == 0) return fail(CLSIMCU_ERR_INVALID, "Steps are empty!");
    if (total > e->max_items) return fail(CLSIMCU_ERR_INVALID, "Number of steps is greater than maximum number of work items!");
    if (total % e->granularity != 0) return fail(CLSIMCU_ERR_INVALID, "The number of steps is not a multiple of the workgroup size!");
    const size_t source_bytes = (n * sizeof(clsimcu_step_source) + 15) & ~size_t(15);
    if (source_bytes + (n + 1) * sizeof(uint64_t) > e->max_items * sizeof(clsimcu_step))
        return fail(CLSIMCU_ERR_INVALID, "too many step sources for this engine's staging buffers (raise max_num_workitems)");
    e->enqueued_any = true;
    Bunch b;
    b.identifier = identifier;
    b.num_steps = total;
    b.generated = photons;
    b.generator = g;
    b.num_sources = n;
    if (int rc = acquire_staging(e, b.staging)) return rc;
    uint8_t *dst = reinterpret_cast<uint8_t *>(e->staging[b.staging]);
    std::memcpy(dst, sources, n * sizeof(clsimcu_step_source));
    std::memcpy(dst + source_bytes, first.data(), (n + 1) * sizeof(uint64_t));
    if (!e->inbox.put(std::move(b))) return interrupted(e);
    return CLSIMCU_OK;
}

int clsimcu_enqueue(clsimcu_engine *e, const clsimcu_step *steps, size_t n, uint32_t identifier)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "I3CLSimStepToPhotonConverterCUDA is not initialized!");
    std::string err;
    if (e->check_async_error(err)) return fail(CLSIMCU_ERR_CUDA, err);
    // same preconditions, same order as EnqueueSteps (…OpenCL.cxx:1527-1540)
    if (!steps) return fail(CLSIMCU_ERR_INVALID, "Steps pointer is (null)!");
    if (n == 0) return fail(CLSIMCU_ERR_INVALID, "Steps are empty!");
    if (n > e->max_items) return fail(CLSIMCU_ERR_INVALID, "Number of steps is greater than maximum number of work items!");
    if (n % e->granularity != 0) return fail(CLSIMCU_ERR_INVALID, "The number of steps is not a multiple of the workgroup size!");
    e->enqueued_any = true;
    Bunch b;
    b.identifier = identifier;
    b.num_steps = n;
    if (int rc = acquire_staging(e, b.staging)) return rc;
    parallel_copy(e->staging[b.staging], steps, n * sizeof(clsimcu_step));
    for (size_t i = 0; i < n; ++i) b.generated += steps[i].num_photons;
    if (!e->inbox.put(std::move(b))) return interrupted(e);
    return CLSIMCU_OK;
}

int clsimcu_get_result(clsimcu_engine *e, clsimcu_result *r)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "I3CLSimStepToPhotonConverterCUDA is not initialized!");
    if (!r) return fail(CLSIMCU_ERR_INVALID, "result pointer is NULL");
    std::shared_ptr<HostResult> res;
    if (!e->outbox.get(res)) return interrupted(e);
    auto *holder = new std::shared_ptr<HostResult>(res);
    static clsimcu_photon empty_photon;
    r->identifier = res->identifier;
    r->reserved0 = 0;
    r->num_photons = res->num_photons;
    r->photons = res->num_photons == 0 ? &empty_photon : res->photons.get();
    r->history = (e->history_entries > 0 && !res->history.empty()) ? res->history.data() : nullptr;
    r->num_photons_generated = res->generated;
    r->num_hits_counted = res->counted;
    r->opaque = holder;
    r->mcpes = res->mcpes.empty() ? nullptr : res->mcpes.data();
    r->num_mcpes = res->mcpes.size();
    return CLSIMCU_OK;
}

int clsimcu_release_result(clsimcu_engine *, clsimcu_result *r)
{
    if (!r) return fail(CLSIMCU_ERR_INVALID, "result pointer is NULL");
    delete static_cast<std::shared_ptr<HostResult> *>(r->opaque);
    std::memset(r, 0, sizeof *r);
    return CLSIMCU_OK;
}

int clsimcu_attach_mcpe_converter(clsimcu_engine *e, clsimcu_mcpe_converter *c, int keep_photons)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "I3CLSimStepToPhotonConverterCUDA is not initialized!");
    if (!c) return fail(CLSIMCU_ERR_INVALID, "MCPE converter is NULL");
    if (e->enqueued_any || e->mcpe) return fail(CLSIMCU_ERR_STATE, "the MCPE converter must be attached once, before the first EnqueueSteps");
    if (mcpe_device(c) != e->device) return fail(CLSIMCU_ERR_INVALID, "the MCPE converter lives on another device than the engine");
    if (e->save_all) return fail(CLSIMCU_ERR_INVALID, "saveAllPhotons records photons that are on no DOM: they cannot be converted to MCPEs");
    try {
        CUDA_OK(cudaSetDevice(e->device));
        for (Slot &s : e->slots) {
            CUDA_OK(cudaMalloc(&s.d_mcpes, e->max_hits * sizeof(clsimcu_mcpe)));
            CUDA_OK(cudaMalloc(&s.d_mcpe_counters, kMcpeCounters * sizeof(uint32_t)));
            CUDA_OK(cudaHostAlloc(&s.h_mcpes, e->max_hits * sizeof(clsimcu_mcpe), cudaHostAllocDefault));
            CUDA_OK(cudaHostAlloc(&s.h_mcpe_counters, kMcpeCounters * sizeof(uint32_t), cudaHostAllocDefault));
        }
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    e->mcpe_keep_photons = keep_photons != 0;
    e->mcpe = c;   // read by the submit thread only after a bunch has been enqueued (queue hand-off orders it)
    return CLSIMCU_OK;
}

int clsimcu_queue_size(clsimcu_engine *e, size_t *size)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "I3CLSimStepToPhotonConverterCUDA is not initialized!");
    *size = e->inbox.size();
    return CLSIMCU_OK;
}

int clsimcu_more_photons_available(clsimcu_engine *e, int *available)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "I3CLSimStepToPhotonConverterCUDA is not initialized!");
    *available = e->outbox.empty() ? 0 : 1;
    return CLSIMCU_OK;
}

int clsimcu_workgroup_size(clsimcu_engine *e, size_t *size)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    *size = e->granularity;
    return CLSIMCU_OK;
}

int clsimcu_max_num_workitems(clsimcu_engine *e, size_t *size)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    *size = e->max_items;
    return CLSIMCU_OK;
}

int clsimcu_get_statistics(clsimcu_engine *e, double out[8])
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    std::lock_guard<std::mutex> lk(e->stats_mutex);
    const double gen = static_cast<double>(e->photons_generated);
    out[0] = e->device_ns;
    out[1] = e->host_ns;
    out[2] = static_cast<double>(e->kernel_calls);
    out[3] = gen;
    out[4] = static_cast<double>(e->photons_at_doms);
    out[5] = e->device_ns / gen;
    out[6] = e->host_ns / gen;
    out[7] = e->device_ns / e->host_ns;
    return CLSIMCU_OK;
}

// ---- resident path ---------------------------------------------------------------------------

int clsimcu_upload_resident(clsimcu_engine *e, const clsimcu_step *steps, size_t n)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    if (!steps || n == 0) return fail(CLSIMCU_ERR_INVALID, "Steps are empty!");
    if (n > e->max_items) return fail(CLSIMCU_ERR_INVALID, "Number of steps is greater than maximum number of work items!");
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        if (!e->d_res_steps) {
            CUDA_OK(cudaMalloc(&e->d_res_steps, e->max_items * sizeof(clsimcu_step)));
            e->res_cap = e->max_hits;
            CUDA_OK(cudaMalloc(&e->d_res_photons, e->res_cap * sizeof(clsimcu_photon)));
            CUDA_OK(cudaMalloc(&e->d_res_counters, 2 * sizeof(uint32_t)));
            CUDA_OK(cudaMalloc(&e->d_res_stats, 8 * sizeof(unsigned long long)));
            CUDA_OK(cudaHostAlloc(&e->h_res_counters, 2 * sizeof(uint32_t), cudaHostAllocDefault));
            CUDA_OK(cudaHostAlloc(&e->h_res_stats, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
            CUDA_OK(cudaMalloc(&e->d_l2_flush, kL2FlushBytes));
            if (e->history_entries > 0) CUDA_OK(cudaMalloc(&e->d_res_history, e->res_cap * e->history_entries * 4 * sizeof(float)));
            if (e->save_all) {
                CUDA_OK(cudaMalloc(&e->d_tag_x, 2 * e->res_cap * sizeof(uint64_t)));
                CUDA_OK(cudaMalloc(&e->d_tag_a, 2 * e->res_cap * sizeof(uint32_t)));
            }
        }
        CUDA_OK(cudaMemcpy(e->d_res_steps, steps, n * sizeof(clsimcu_step), cudaMemcpyHostToDevice));
        // (a copy from pageable memory may return while its last piece is still on its way from the driver's staging
        // buffer, and the compute stream does not synchronise with the default stream: wait for it here)
        CUDA_OK(cudaStreamSynchronize(nullptr));
        e->res_steps = n;
        uint64_t gen = 0;
        for (size_t i = 0; i < n; ++i) gen += steps[i].num_photons;
        e->res_generated_per_run = gen;
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

int clsimcu_run_resident(clsimcu_engine *e, int repeat, double *kernel_ms, uint64_t *photons_generated, uint64_t *hits_counted,
                         uint64_t *segments)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    if (e->res_steps == 0) return fail(CLSIMCU_ERR_STATE, "no resident bunch uploaded");
    if (repeat < 1) return fail(CLSIMCU_ERR_INVALID, "repeat must be >= 1");
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        std::vector<cudaEvent_t> ev(2 * repeat);
        for (auto &x : ev) CUDA_OK(cudaEventCreate(&x));
        uint64_t hits = 0, segs = 0;
        double ms = 0.;
        CUDA_OK(cudaMemsetAsync(e->d_res_stats, 0, 8 * sizeof(unsigned long long), e->compute));
        for (int r = 0; r < repeat; ++r) {
            CUDA_OK(cudaMemsetAsync(e->d_res_counters, 0, 2 * sizeof(uint32_t), e->compute));
            CUDA_OK(cudaMemsetAsync(e->d_l2_flush, r & 0xff, kL2FlushBytes, e->compute)); // L2 flush, outside the timed events
            LaunchArgs a{};
            a.steps = e->d_res_steps;
            a.num_steps = static_cast<uint32_t>(e->res_steps);
            a.max_hits = static_cast<uint32_t>(e->res_cap);
            a.photons = e->d_res_photons;
            a.history = e->d_res_history;
            a.history_ring = e->d_history_ring;
            a.hit_counter = e->d_res_counters;
            a.work_counter = e->d_res_counters + 1;
            a.stats = e->d_res_stats;
            a.rng_x = e->d_rng_x;
            a.rng_a = e->d_rng_a;
            a.rng_tag_x = e->d_tag_x;
            a.rng_tag_a = e->d_tag_a;
            a.count_stats = 1;
            a.scene_dev = e->d_scene;
            a.rng_creation_offset = static_cast<uint32_t>(e->fast_blocks) * e->fast_threads;
            CUDA_OK(cudaEventRecord(ev[2 * r], e->compute));
            launch(*e, a, e->compute);
            CUDA_OK(cudaEventRecord(ev[2 * r + 1], e->compute));
            CUDA_OK(cudaMemcpyAsync(e->h_res_counters, e->d_res_counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->compute));
            CUDA_OK(cudaStreamSynchronize(e->compute));
            hits += e->h_res_counters[0];
            e->res_last_hits = e->h_res_counters[0];
            float one = 0.f;
            CUDA_OK(cudaEventElapsedTime(&one, ev[2 * r], ev[2 * r + 1]));
            ms += one;
        }
        CUDA_OK(cudaMemcpy(e->h_res_stats, e->d_res_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        segs = e->h_res_stats[1];
        if (std::getenv("CLSIMCU_DEBUG_STATS")) // counters of a -DCLSIMCU_DEBUG_COUNTERS build of the fast kernel
            std::fprintf(stderr, "clsimcu debug stats: %llu %llu %llu %llu %llu %llu\n", e->h_res_stats[2], e->h_res_stats[3], e->h_res_stats[4],
                         e->h_res_stats[5], e->h_res_stats[6], e->h_res_stats[7]);
        for (auto &x : ev) cudaEventDestroy(x);
        if (kernel_ms) *kernel_ms = ms;
        // photons CREATED, as counted by the kernel itself (stats[0]) -- not the host's sum over the step records
        if (photons_generated) *photons_generated = e->h_res_stats[0];
        if (hits_counted) *hits_counted = hits;
        if (segments) *segments = segs;
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

int clsimcu_download_resident(clsimcu_engine *e, clsimcu_photon *out, size_t cap, size_t *n)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        const size_t have = std::min<size_t>(e->res_last_hits, e->res_cap);
        const size_t k = std::min(have, cap);
        if (k > 0 && out) CUDA_OK(cudaMemcpy(out, e->d_res_photons, k * sizeof(clsimcu_photon), cudaMemcpyDeviceToHost));
        if (n) *n = have;
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

int clsimcu_download_resident_rng_tags(clsimcu_engine *e, uint64_t *x, uint32_t *a, size_t cap)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    if (!e->d_tag_x) return fail(CLSIMCU_ERR_STATE, "RNG tags are only recorded in save-all mode on the resident path");
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        const size_t k = std::min(std::min<size_t>(e->res_last_hits, e->res_cap), cap);
        if (k > 0) {
            CUDA_OK(cudaMemcpy(x, e->d_tag_x, 2 * k * sizeof(uint64_t), cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemcpy(a, e->d_tag_a, 2 * k * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        }
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

int clsimcu_download_resident_history(clsimcu_engine *e, float *out, size_t cap, size_t *n)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    if (e->history_entries <= 0 || !e->d_res_history) return fail(CLSIMCU_ERR_STATE, "no photon history was requested (photon_history_entries is 0)");
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        const size_t have = std::min<size_t>(e->res_last_hits, e->res_cap);
        const size_t k = std::min(have, cap);
        if (k > 0 && out) {
            std::vector<clsimcu_photon> photons(k);
            std::vector<float> raw(k * e->history_entries * 4);
            CUDA_OK(cudaMemcpy(photons.data(), e->d_res_photons, k * sizeof(clsimcu_photon), cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemcpy(raw.data(), e->d_res_history, raw.size() * sizeof(float), cudaMemcpyDeviceToHost));
            std::vector<float> ordered;
            unroll_history(raw.data(), photons.data(), k, e->history_entries, ordered);
            std::memcpy(out, ordered.data(), ordered.size() * sizeof(float));
        }
        if (n) *n = have;
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

// ---- test hooks --------------------------------------------------------------------------------

int clsimcu_rng_get(clsimcu_engine *e, uint64_t *x, uint32_t *a, size_t n)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    if (n > e->rng_n) return fail(CLSIMCU_ERR_INVALID, "more RNG streams requested than exist");
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        CUDA_OK(cudaStreamSynchronize(e->compute));
        if (x) CUDA_OK(cudaMemcpy(x, e->d_rng_x, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (a) CUDA_OK(cudaMemcpy(a, e->d_rng_a, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

int clsimcu_rng_set(clsimcu_engine *e, const uint64_t *x, const uint32_t *a, size_t n)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    if (n > e->rng_n) return fail(CLSIMCU_ERR_INVALID, "more RNG streams given than exist");
    try {
        std::lock_guard<std::mutex> lk(e->compute_mutex);
        CUDA_OK(cudaSetDevice(e->device));
        CUDA_OK(cudaStreamSynchronize(e->compute));
        if (x) CUDA_OK(cudaMemcpy(e->d_rng_x, x, n * sizeof(uint64_t), cudaMemcpyHostToDevice));
        if (a) CUDA_OK(cudaMemcpy(e->d_rng_a, a, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaStreamSynchronize(nullptr));   // (pageable source: see clsimcu_upload_resident)
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_CUDA, ex.what());
    }
    return CLSIMCU_OK;
}

static int copy_text(const std::string &s, char *buf, size_t cap, size_t *needed)
{
    if (needed) *needed = s.size() + 1;
    if (!buf) return CLSIMCU_OK;
    if (cap < s.size() + 1) return fail(CLSIMCU_ERR_INVALID, "buffer too small");
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return CLSIMCU_OK;
}

int clsimcu_describe_tables(clsimcu_engine *e, char *buf, size_t cap, size_t *needed)
{
    if (!e) return fail(CLSIMCU_ERR_STATE, "engine is NULL");
    return copy_text(describe_scene_tables(e->tables), buf, cap, needed);
}

int clsimcu_describe_tables_from_config(const clsimcu_config *config, char *buf, size_t cap, size_t *needed)
{
    if (!config) return fail(CLSIMCU_ERR_INVALID, "config is NULL");
    if (config->struct_size != static_cast<int32_t>(sizeof(clsimcu_config)))
        return fail(CLSIMCU_ERR_INVALID, "clsimcu_config.struct_size does not match this library");
    try {
        SceneTables t;
        build_scene_tables(*config, t);
        return copy_text(describe_scene_tables(t), buf, cap, needed);
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_INVALID, ex.what());
    }
}

int clsimcu_describe_collision_map_from_config(const clsimcu_config *config, int32_t pixel_budget, char *buf, size_t cap, size_t *needed)
{
    if (!config) return fail(CLSIMCU_ERR_INVALID, "config is NULL");
    if (config->struct_size != static_cast<int32_t>(sizeof(clsimcu_config)))
        return fail(CLSIMCU_ERR_INVALID, "clsimcu_config.struct_size does not match this library");
    if (pixel_budget < 1) return fail(CLSIMCU_ERR_INVALID, "pixel_budget must be positive");
    try {
        SceneTables t;
        build_scene_tables(*config, t);
        if (!t.has_geometry) return fail(CLSIMCU_ERR_INVALID, "the configuration has no geometry");
        return copy_text(describe_collision_map(t.geometry, build_collision_map(t.geometry, pixel_budget)), buf, cap, needed);
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_INVALID, ex.what());
    }
}

int clsimcu_safeprime_multipliers(uint64_t first, uint64_t n, uint32_t *a)
{
    if (!a && n) return fail(CLSIMCU_ERR_INVALID, "output pointer is NULL");
    try {
        safeprime_multipliers(first, n, a, prime_cache_path());
    } catch (const std::exception &ex) {
        return fail(CLSIMCU_ERR_INVALID, ex.what());
    }
    return CLSIMCU_OK;
}

int clsimcu_seed_rng_states(uint64_t seed, const uint32_t *a, uint64_t *x, uint64_t n)
{
    if ((!a || !x) && n) return fail(CLSIMCU_ERR_INVALID, "multiplier or state pointer is NULL");
    seed_rng_states(seed, a, x, static_cast<size_t>(n));
    return CLSIMCU_OK;
}

} // extern "C"
