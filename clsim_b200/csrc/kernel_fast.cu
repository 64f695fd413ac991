// kernel_fast.cu -- persistent photon-propagation kernel for B200 (sm_100a).  The product path.
//
// Design (B200-first, not the reference's one-work-item-per-step loop):
//
//  * Persistent CTAs (a multiple of the SM count).  A WARP owns a step at a time and pulls the
//    next one from a global work counter, so step bunches of any size balance across the 148 SMs
//    and there is no per-step tail.
//  * One photon per LANE.  The kernel alternates between two phases per warp:
//      FAST  -- a straight-line loop with the photon state in registers: every iteration moves
//               each live photon to its next EVENT (scatter, absorption, ice-layer boundary, or
//               the range limit of the collision map).  A lane whose photon ended takes the next
//               one from the warp's QUEUE in shared memory right inside the loop, so all 32 lanes
//               stay busy.  No calls, no rare paths, no spills.
//      SLOW  -- out of line, lane state parked in shared memory; entered when a leg might touch
//               a DOM (about one leg in 600) or the queue has run dry (every 32 photons): the
//               reference's full collision test, hit output, and the queue refill by all 32
//               lanes at once (photon creation: wavelength table search, Cherenkov cone, the
//               wavelength-only transcendental factors of the ice model, lifetime).
//    Batching the rare work this way keeps its cost per photon small and, more importantly,
//    keeps it from diverging the hot loop on almost every iteration.
//  * A photon carries a 3-word BIRTH TAG (creation-stream state, step index, creating lane)
//    instead of its 10-word start-of-flight record: one photon in a thousand is detected, and for
//    those the record is re-created from the tag (creation is deterministic).
//  * Two MWC streams per lane (multipliers from the safe-prime table): one drives creation, one
//    drives propagation, so a photon's propagation draws are contiguous in its stream and any
//    photon can be replayed by a checker from two recorded states.
//  * Flights are cut at ice-layer boundaries, so the reference's data-dependent walk over layers
//    (propagation_kernel.c.cl:647-662) becomes straight-line code; the budgets (scattering /
//    absorption lengths left) are carried in registers.  Same distances up to fp32 rounding.
//  * The wavelength dependence of the ice (powr/exp of R4) is hoisted to once per photon: per
//    iteration only two FMAs on per-layer coefficients remain, which live in shared memory as one
//    float4 per layer together with the collision tables.
//  * DOM collision: an xy pixel map in shared memory names the string nearest to the photon and
//    how far the photon may fly before any other string comes into range (flights are cut there
//    too).  A 2-D segment/cylinder test against that one string rules out > 99.9 % of the
//    segments in ~25 instructions; the rest run the reference's string / cell walk in the slow
//    phase.
//  * The kernel is bound by instruction issue, so the arithmetic is written for few instructions: sm_100a's packed
//    fp32 instructions (FFMA2 / FMUL2 / FADD2, two operations per issue slot) on the (x, y) pairs of position and
//    direction, the layer coefficients and the wavelength factors; the SL + HG scattering mix from constants folded
//    on the host (both samplers evaluated, one selected: no divergent branch); one reciprocal root for the whole
//    rotation; approximate MUFU intrinsics throughout.
//  * The loop head (ballots, refill) runs every third leg, and only the first of the three legs consults the
//    collision map: it leaves the lane a clearance (distance it may fly before any string can come into play) on
//    which the other two fly without map, test or range limit.
//  * Hits: warp-aggregated atomic reservation, five 16-byte stores per record, string/DOM IDs
//    and the wavelength-bias weight applied on the device.
//
// Tensor cores are not used: there is no contraction anywhere on this path.
//
// Physics restated from resources/kernels/propagation_kernel.c.cl and
// sparse_collision_kernel.c.cl (citations at each block); the arithmetic is re-formulated, so
// agreement with the reference is statistical (per-DOM counts, time and angle distributions)
// and per photon within fp32 tolerance when a photon is replayed from its recorded RNG states.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/clsimcuda.h"
#include "device_scene.h"
#include "tabulate_device.h"

namespace clsimcu {
namespace {

#ifndef CLSIMCU_THREADS
#define CLSIMCU_THREADS 1024
#endif
#ifndef CLSIMCU_BLOCKS_PER_SM
#define CLSIMCU_BLOCKS_PER_SM 1
#endif
#ifndef CLSIMCU_IDLE_LIMIT
#define CLSIMCU_IDLE_LIMIT 1
#endif
#ifndef CLSIMCU_IDLE_LIMIT_SAVE_ALL
#define CLSIMCU_IDLE_LIMIT_SAVE_ALL 8
#endif
#ifndef CLSIMCU_HOT_UNROLL
#define CLSIMCU_HOT_UNROLL 3
#endif
#ifndef CLSIMCU_REFILL_BATCH
#define CLSIMCU_REFILL_BATCH 3
#endif
#ifndef CLSIMCU_WLEN_BIN_RECORDS
#define CLSIMCU_WLEN_BIN_RECORDS 1  // A/B knob: generator 0's trapezoid inversion from per-bin records
#endif
constexpr int kThreads = CLSIMCU_THREADS;
// legs a lane flies between two looks at the warp's state (ballots, refill decisions) in the hot loop
constexpr int kHotUnroll = CLSIMCU_HOT_UNROLL;
// ... for the ice models with tilt or anisotropy, whose legs are twice as long in code (instruction cache)
#ifndef CLSIMCU_HOT_UNROLL_TILTED
#define CLSIMCU_HOT_UNROLL_TILTED 3
#endif
constexpr int kHotUnrollTilted = CLSIMCU_HOT_UNROLL_TILTED;
#ifdef CLSIMCU_LOOK_EVERY_LEG
constexpr bool kLookEveryLeg = true;    // A/B knob: consult the collision map on every leg
#else
constexpr bool kLookEveryLeg = false;
#endif
constexpr int kRefillBatch = CLSIMCU_REFILL_BATCH;            // lanes without a photon that make the warp stop for a refill
constexpr int kWarpsPerBlock = kThreads / 32;
constexpr int kBlocksPerSM = CLSIMCU_BLOCKS_PER_SM;
constexpr int kIdleLimit = CLSIMCU_IDLE_LIMIT;                  // parked lanes (possible DOM contact) that end a fast phase
constexpr int kIdleLimitSaveAll = CLSIMCU_IDLE_LIMIT_SAVE_ALL;  // save-all: every photon ends in the slow phase
constexpr float kSpeedOfLight = 0.299792458f;
constexpr float kPi = 3.14159265359f;
constexpr float kEpsilon = 0.00001f;
constexpr float kLn2 = 0.69314718056f;
constexpr uint32_t kSmemBudget = 227u * 1024u;
constexpr int kGen0MaxEntries = 256;   // the Cherenkov generator of the IceCube setups has 43
constexpr int kGen0Guide = 64;         // cells of the guide table over (0, 1]

// queue slot (one photon waiting for a lane), four 16-byte chunks so that a lane takes a photon with four loads:
//   0: (x, y, dx, dy)   1: (z, dz, lifetime in absorption lengths, ice layer)   2: (f_scat, f_pure, f_dust, -)
//   3: BIRTH TAG (creation-stream state lo, hi, step index | creating lane << 27, -)
// f_*: the three wavelength-only ice factors.  The start-of-flight record a hit needs (start point, direction, time,
// wavelength, lifetime) is not carried along: one photon in a thousand is detected, and for those the
// record is re-created from the tag (creation is deterministic).
constexpr int kQueueChunks = 4;
constexpr int kQueueWords = 4 * kQueueChunks;
constexpr uint32_t kStepIndexBits = kFastKernelStepIndexBits; // third tag word = step index | (creating lane << 27)
// per-lane running state, parked in shared memory between fast phases
enum StateWord {
    kPx = 0, kPy, kPz, kDx, kDy, kDz, kAbsLeft, kScaLeft, kPath, kFScat, kFDust, kFPure, kScatters, kLayer, kStatus, kRngLo, kRngHi,
    kZEff, kInvAniso, kPendTravel, kPendWho, kStateWords
};
// per-lane birth tag of the photon in flight (3 words) and, in the save-all variants, the state of the
// propagation stream when it started (2 words): what a checker needs to replay the photon
constexpr int kTagRecordWords = 4;   // per thread: birth tag (3 words) | flights (reference: segments) of the photons the lane has finished
constexpr int kPopTagWords = 2;
// per-warp control block
enum WarpCtl { kWLeft = 0, kWStepIndex, kWMore, kWQueued, kWCreated, kWarpCtlWords = 8 };
constexpr int kWarpStepWords = 16; // the warp's step record (12 words) + its direction (3)
// compile-time offsets (in words) inside the per-thread and per-warp regions of shared memory
constexpr int kOffBirthTag = kStateWords * kThreads;
constexpr int kOffPopTag = kOffBirthTag + kTagRecordWords * kThreads;
constexpr int kPerThreadWords = kStateWords + kTagRecordWords;
constexpr int kOffWarpStep = kWarpsPerBlock * kQueueWords * 32;
constexpr int kOffWarpCtl = kOffWarpStep + kWarpsPerBlock * kWarpStepWords;

enum Status : uint32_t { kActive = 0, kFrozen = 1, kDying = 2, kDead = 3 };
// rare variants of the kernel (template parameter V, a bit set): photon history ring, table-maker sink, StopDetectedPhotons = false
// (kept out of the default instantiation: ptxas allocates registers over the whole call graph, and the slow path's extra
// live values cost the hot loop spills)
constexpr int kVarHist = 1, kVarTab = 2, kVarNonStop = 4;
// ... and one frequent specialisation: direction transforms of the form [[a, c, 0], [d, b, 0], [0, 0, e]] (the ppc anisotropy:
// a stretch in the horizontal plane and along z), five multiplications instead of nine and five constants instead of nine
constexpr int kVarBlockTransforms = 8;

// ---- approximate MUFU wrappers ---------------------------------------------------------------
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_pow(float x, float y) { return mufu_ex2(y * mufu_lg2(x)); }
__device__ __forceinline__ float fast_ln(float x) { return kLn2 * mufu_lg2(x); }

struct Mwc {
    uint64_t x;
    uint32_t a;
    // mwcrng_kernel.cl:12-28
    // (written as one mad.wide on the two halves: ptxas then emits IMAD.WIDE + a two-instruction carry chain instead
    // of shuffling the state through an aligned register pair)
    __device__ __forceinline__ uint32_t next()
    {
        const uint32_t lo = static_cast<uint32_t>(x), hi = static_cast<uint32_t>(x >> 32);
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(x) : "r"(lo), "r"(a), "l"(static_cast<uint64_t>(hi)));
        return static_cast<uint32_t>(x);
    }
    // the conversion rounds toward zero so 1.0 is never returned
    __device__ __forceinline__ float co() { return __uint2float_rz(next()) * 2.3283064365386963e-10f; }
    // 1 - u/2^32 in one FMA (the scaling is exact, so this equals the two-step form)
    __device__ __forceinline__ float oc() { return __fmaf_rn(__uint2float_rz(next()), -2.3283064365386963e-10f, 1.0f); }
};

__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ uint64_t pack64(float lo, float hi)
{
    return static_cast<uint64_t>(__float_as_uint(lo)) | (static_cast<uint64_t>(__float_as_uint(hi)) << 32);
}

// Shared-memory plan, carved out of the dynamic allocation.
struct SmemLayout {
    uint32_t off_layers, off_strings, off_sets, off_string_set, off_layer_to_dom, off_cells, off_near, off_tilt_dist, off_tilt_corr;
    uint32_t off_gen0;    // generator 0's cumulative | density | guide tables (gen0_n entries each, 64 guide bytes), see draw_wavelength
    uint32_t off_gen0_bins; // ... and its per-bin inversion records (float4 + float2 per bin)
    uint32_t gen0_n;      // 0: not staged
    uint32_t off_state;   // per-thread arrays: state | birth tag | segment counter | propagation-stream tag (save-all only)
    uint32_t off_queue;   // per-warp arrays: photon queues | step records | control blocks
    uint32_t cell_offset[kMaxSubdetectors];
    uint32_t total;
};

// First bytes of the dynamic shared memory: the layout and the launch arguments, so that
// out-of-line device functions find everything from the shared-memory base alone.
struct SmemHeader {
    SmemLayout lay;
    LaunchArgs args;
    TabulateArgs tab;   // table-maker variant: the table's description (args.tabulate points at the original in HBM)
};

struct SmemPlan {
    float4 *layers;          // [num_layers + 1] (b400, 1+0.01*dTau, D*aDust+E, lower boundary z); the upper boundary of a
                             // layer is the lower one of the next entry; -/+1e30 where there is no layer beyond
    float4 *strings;         // [num_strings] (x, y, zmax+R, zmin-R)
    float4 *sets;            // [num_sets] (start_z, 1/height, num_layers, row offset)
    uint8_t *string_set;     // [num_strings]
    uint16_t *layer_to_dom;  // [layer_table_size]
    uint16_t *cells;         // concatenated grids
    uint32_t *near;          // xy pixel map, see device_scene.h
    float2 *tilt_dist;       // [tilt_nd] (distance along the tilt direction, 1 / (this - previous)); ice tilt only
    float4 *tilt_corr;       // [tilt_nd - 1][tilt_nz - 1] cells of the correction table: (lower column, upper column, their steps to the next z node)
};

__host__ __device__ constexpr uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

// The regions of fixed size come first, at compile-time offsets: the hot loop then addresses the layer table, the
// pixel map, the lane's state column and the warp's queue as  register + immediate  and keeps no base pointers alive
// (the kernel runs at the 64-register limit; a base pointer in a register is a spill somewhere else).
constexpr int kMaxStagedLayers = 256;   // entries of the layer table, the sentinel included
constexpr uint32_t kSmLayers = align16(static_cast<uint32_t>(sizeof(SmemHeader)));
constexpr uint32_t kSmQueue = kSmLayers + kMaxStagedLayers * 16u;
constexpr uint32_t kSmState = kSmQueue + align16(kWarpsPerBlock * (kQueueWords * 32 + kWarpStepWords + kWarpCtlWords) * 4u);
constexpr uint32_t kSmNear = kSmState + kPerThreadWords * kThreads * 4u;   // (save-all: the propagation-stream tags sit here, there is no pixel map)

__host__ SmemLayout plan_smem(const DevScene &s)
{
    SmemLayout L{};
    L.off_layers = kSmLayers;
    L.off_queue = kSmQueue;
    L.off_state = kSmState;
    uint32_t at = align16(kSmNear + (s.save_all ? kPopTagWords * kThreads * 4u : 0u));
    L.off_near = at; at = align16(at + s.geo.near_nx * s.geo.near_ny * 4);
    L.off_strings = at; at = align16(at + (s.geo.num_strings + 1) * 16);
    L.off_sets = at; at = align16(at + s.geo.num_sets * 16);
    L.off_string_set = at; at = align16(at + s.geo.num_strings);
    L.off_layer_to_dom = at; at = align16(at + s.geo.layer_table_size * 2);
    L.off_cells = at;
    uint32_t cells = 0;
    for (int i = 0; i < s.geo.num_grids; ++i) {
        L.cell_offset[i] = cells;
        cells += s.geo.grids[i].num_x * s.geo.grids[i].num_y;
    }
    at = align16(at + cells * 2);
    L.off_tilt_dist = at; at = align16(at + (s.medium.tilt_nd + (s.medium.tilt_nd & 1)) * 8 + 8 * kTiltLutMaxCells);   // + the interval grid
    L.off_tilt_corr = at; at = align16(at + (s.medium.tilt_nd > 1 ? (s.medium.tilt_nd - 1) * (s.medium.tilt_nz - 1) * 16 : 0));
    L.off_gen0 = at;
    L.gen0_n = 0;
    if (s.num_generators >= 1 && (s.generators[0].kind == CLSIMCU_WLEN_INTERP_EQUAL || s.generators[0].kind == CLSIMCU_WLEN_INTERP_UNEQUAL) &&
        s.generators[0].n >= 2 && s.generators[0].n <= kGen0MaxEntries) {
        L.gen0_n = static_cast<uint32_t>(s.generators[0].n);
        at = align16(at + 2 * L.gen0_n * 4 + kGen0Guide);
        L.off_gen0_bins = at;
        at = align16(at + (CLSIMCU_WLEN_BIN_RECORDS ? L.gen0_n * 24 : 0));
    }
    L.total = at;
    return L;
}

__device__ __forceinline__ uint8_t *smem_base()
{
    extern __shared__ __align__(16) uint8_t smem[];
    return smem;
}

// word `w` of the calling thread's tag record; `st` is its column of the per-thread state arrays (the record is laid out
// like them, [word][thread]: every access is `st` + a compile-time offset, no second pointer to keep alive in the hot loop)
__device__ __forceinline__ uint32_t &tag_word(float *st, int w) { return reinterpret_cast<uint32_t *>(st + kOffBirthTag)[w * kThreads]; }

__device__ __forceinline__ SmemPlan table_plan(const SmemLayout &lay)
{
    uint8_t *smem = smem_base();
    SmemPlan sp;
    sp.layers = reinterpret_cast<float4 *>(smem + lay.off_layers);
    sp.strings = reinterpret_cast<float4 *>(smem + lay.off_strings);
    sp.sets = reinterpret_cast<float4 *>(smem + lay.off_sets);
    sp.string_set = smem + lay.off_string_set;
    sp.layer_to_dom = reinterpret_cast<uint16_t *>(smem + lay.off_layer_to_dom);
    sp.cells = reinterpret_cast<uint16_t *>(smem + lay.off_cells);
    sp.near = reinterpret_cast<uint32_t *>(smem + lay.off_near);
    sp.tilt_dist = reinterpret_cast<float2 *>(smem + lay.off_tilt_dist);
    sp.tilt_corr = reinterpret_cast<float4 *>(smem + lay.off_tilt_corr);
    return sp;
}

struct V3 {
    float x, y, z;
};

// R8 (propagation_kernel.c.cl:83-129) with approximate MUFU ops
__device__ __forceinline__ void rotate_by(float cosa, float sina, V3 &d, float rnd, bool renormalize = true)
{
    float sinb, cosb;
    __sincosf(2.0f * kPi * rnd, &sinb, &cosb);
    // sin^2(theta) from x and y, not as 1 - z^2: for a direction whose length is off by eps the rotation below then
    // gives a length off by at most eps again (with 1 - z^2 the error is amplified by sin^2(a)/sin^2(theta) near the
    // poles), so the length only random-walks by rounding and need not be restored after every scatter
    const float s2 = fmaf(d.x, d.x, d.y * d.y);
    float nx, ny, nz;
    if (s2 > 0.f) {
        const float inv_s = mufu_rsqrt(s2);
        const float sinth = s2 * inv_s;
        const float k = sina * inv_s;
        nx = d.x * cosa - (d.y * cosb + d.z * d.x * sinb) * k;
        ny = d.y * cosa + (d.x * cosb - d.z * d.y * sinb) * k;
        nz = d.z * cosa + sina * sinb * sinth;
    } else {
        nx = sina * cosb;
        ny = sina * sinb;
        nz = (d.z > 0.f) ? cosa : ((d.z < 0.f) ? -cosa : cosa * d.z);
    }
    // the rotation preserves the length up to rounding (a few 1e-7 per scatter with the approximate sine and
    // cosine), so one Newton step of 1/sqrt about 1 is exact to fp32 here and keeps the special-function unit
    // free; the hot loop leaves it to the start of each fast phase
    if (renormalize) {
        const float inv = fmaf(nx * nx + ny * ny + nz * nz, -0.5f, 1.5f);
        nx *= inv; ny *= inv; nz *= inv;
    }
    d.x = nx; d.y = ny; d.z = nz;
}

__device__ __forceinline__ float phase_index(const DevMedium &m, float wlen)
{
    const float u = wlen * 1e6f;
    return m.n_phase[0] + u * (m.n_phase[1] + u * (m.n_phase[2] + u * (m.n_phase[3] + u * m.n_phase[4])));
}
__device__ __forceinline__ float inv_group_velocity(const DevMedium &m, float wlen)
{
    const float u = wlen * 1e6f;
    const float np = m.n_phase[0] + u * (m.n_phase[1] + u * (m.n_phase[2] + u * (m.n_phase[3] + u * m.n_phase[4])));
    const float corr = m.n_group[0] + u * (m.n_group[1] + u * (m.n_group[2] + u * (m.n_group[3] + u * m.n_group[4])));
    return np * corr * mufu_rcp(m.c_light);
}

// R3a: same bin as the reference's linear scan (cumulative is non-decreasing): smallest k in [0, n-2] with
// cumulative[k+1] >= r.  `staged` != nullptr: the generator's cumulative and density tables sit in shared memory
// behind a guide table (guide[c] = the bin of r = c/64, a lower bound for every r in that cell), and the bin is found
// by a forward scan of a step or two from there; otherwise by bisection over the tables in global memory.
__device__ float draw_wavelength(const DevWlenGenerator &g, Mwc &rng, const float *staged = nullptr, const float *bins = nullptr)
{
    if (g.kind == CLSIMCU_WLEN_CONSTANT) return g.value;
    const float r = rng.oc();
    if (g.kind == CLSIMCU_WLEN_NO_DISPERSION) return mufu_rcp(g.min_val + r * g.range);
    int k;
    float below, b, b_next;
    if (staged) {
        const float *cum = staged, *dens = staged + g.n;
        const uint8_t *guide = reinterpret_cast<const uint8_t *>(staged + 2 * g.n);
        k = guide[min(__float2int_rz(r * static_cast<float>(kGen0Guide)), kGen0Guide - 1)];
        while (k < g.n - 2 && cum[k + 1] < r) ++k;
#if CLSIMCU_WLEN_BIN_RECORDS
        if (bins) {
            // the four cases of the reference's inversion (I3CLSimRandomValueInterpolatedDistribution.cxx:308-334) as
            // ONE expression on per-bin constants made when the tables were staged:
            //   x0 + (sqrt(dy c1 + c0) - c0) c2 + dy c3
            // c0 = 1, c1 = 2 slope / b^2, c2 = b / slope in the general case; b == 0: c0 = 0, c1 = 2 / slope, c2 = 1;
            // slope == 0: c3 = 1 / b, the rest 0; both 0: all 0
            const float4 rec = reinterpret_cast<const float4 *>(bins)[k];
            const float2 lin = reinterpret_cast<const float2 *>(bins + 4 * g.n)[k];
            const float dy = r - rec.x;
            return fmaf(dy, lin.y, fmaf(mufu_sqrt(fmaf(dy, rec.z, lin.x)) - lin.x, rec.w, rec.y));
        }
#endif
        below = (k == 0) ? 0.f : cum[k];
        b = dens[k];
        b_next = dens[k + 1];
    } else {
        int lo = 0, hi = g.n - 2;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(g.cumulative + mid + 1) >= r) hi = mid;
            else lo = mid + 1;
        }
        k = lo;
        below = (k == 0) ? 0.f : __ldg(g.cumulative + k);
        b = __ldg(g.density + k);
        b_next = __ldg(g.density + k + 1);
    }
    // (approximate reciprocals and root: the correction to x0 is at most one table interval, a few percent of the
    // wavelength, so their 2^-22 relative error stays below 1e-7 of the result)
    float x0, slope;
    if (g.kind == CLSIMCU_WLEN_INTERP_UNEQUAL) {
        x0 = __ldg(g.xs + k);
        slope = (b_next - b) * mufu_rcp(__ldg(g.xs + k + 1) - x0);
    } else {
        x0 = static_cast<float>(k) * g.dx + g.x0;
        slope = (b_next - b) * mufu_rcp(g.dx);
    }
    const float dy = r - below;
    if ((b == 0.f) && (slope == 0.f)) return x0;
    if (b == 0.f) return x0 + mufu_sqrt(2.f * dy * mufu_rcp(slope));
    if (slope == 0.f) return x0 + dy * mufu_rcp(b);
    return x0 + (mufu_sqrt(dy * (2.f * slope) * mufu_rcp(b * b) + 1.f) - 1.f) * b * mufu_rcp(slope);
}

__device__ float bias_at(const DevBias &b, float wlen)
{
    if (b.kind == CLSIMCU_BIAS_CONSTANT) return b.value;
    float whole;
    float frac = modff((wlen - b.x0) / b.dx, &whole);
    int bin = static_cast<int>(whole);
    if ((bin < 0) || ((bin == 0) && (frac < 0))) { bin = 0; frac = 0.f; }
    else if (bin >= b.n - 1) { bin = b.n - 2; frac = 1.f; }
    const float lo = __ldg(b.v + bin), hi = __ldg(b.v + bin + 1);
    return lo + (hi - lo) * frac;
}

__device__ void to_spherical(float x, float y, float z, float &theta, float &phi)
{
    const float inv = rsqrtf(x * x + y * y + z * z);
    theta = 0.f;
    if (fabsf(z * inv) <= 1.f) theta = acosf(z * inv);
    else if (z < 0.f) theta = kPi;
    phi = atan2f(y, x);
    if (phi < 0.f) phi += 2.f * kPi;
}

__device__ __forceinline__ void dom_centre(const DevGeometry &g, int string, int dom, float &x, float &y, float &z)
{
    const uint32_t at = __ldg(g.string_tmpl_start + string) + static_cast<uint32_t>(dom);
    x = static_cast<float>(__ldg(g.tmpl_dx + at)) * g.tmpl_scale_x + __ldg(g.string_mean_x + string);
    y = static_cast<float>(__ldg(g.tmpl_dy + at)) * g.tmpl_scale_y + __ldg(g.string_mean_y + string);
    z = __ldg(g.tmpl_z + at);
}

// R4a (I3CLSimScalarFieldIceTiltZShift.cxx:145-216): the z correction, interpolated linearly between the z nodes of two
// neighbouring columns of the tilt table and then between the columns.  Restated for few instructions: the table sits in
// shared memory as one float4 per (column interval, z interval) -- the two columns' values at the lower z node and their
// steps to the upper one -- so the z interpolation of both columns is one LDS.128 and one packed FMA; the column interval
// comes from a uniform grid over the distance along the tilt direction (see the staging code), its reciprocal width from
// the distance table.  Same piecewise-bilinear function as the reference's, different rounding in the last bits.
__device__ __forceinline__ float tilt_shift(const DevMedium &m, const float2 *dist, const float4 *corr, float x, float y, float z)
{
    const float zr = fmaf(z, m.tilt_inv_dz, m.tilt_zr_offset);
    const int k = min(max(__float2int_rd(zr), 0), m.tilt_nz - 2);
    const float above = zr - static_cast<float>(k);
    const float nr = m.tilt_lnx * x + m.tilt_lny * y;
    // j = first interval end in [1, nd-1] with nr < dist[j], or nd-1 (the reference's scan, :178-186).  The distances
    // ascend, so j = 1 + the number of interior nodes not above nr: read from a uniform grid over nr whose cells hold at
    // most one node each -- (nodes below the cell, the node inside it) -- behind the table
    int j = 1;
    if (m.tilt_lut_n > 0) {
        const float2 *lut = dist + m.tilt_nd + (m.tilt_nd & 1);
        const float2 e = lut[min(max(__float2int_rz(fmaf(nr, m.tilt_lut_scale, m.tilt_lut_offset)), 0), m.tilt_lut_n - 1)];
        j = __float_as_int(e.x) + ((nr < e.y) ? 0 : 1);
    } else {
#pragma unroll 1
        for (int i = 1; i < m.tilt_nd - 1; ++i) j += (nr < dist[i].x) ? 0 : 1;
    }
    const float2 here = dist[j];
    const float w_lo = (here.x - nr) * here.y;
    const float4 c = corr[(j - 1) * (m.tilt_nz - 1) + k];
    const float2 v = __ffma2_rn(make_float2(c.z, c.w), make_float2(above, above), make_float2(c.x, c.y));   // (lower, upper column) at z
    return fmaf(w_lo, v.x - v.y, v.y);
}

__device__ __forceinline__ void apply_matrix(const float *M, V3 &d)
{
    const float nx = M[0] * d.x + M[1] * d.y + M[2] * d.z;
    const float ny = M[3] * d.x + M[4] * d.y + M[5] * d.z;
    const float nz = M[6] * d.x + M[7] * d.y + M[8] * d.z;
    const float inv = mufu_rsqrt(nx * nx + ny * ny + nz * nz);
    d.x = nx * inv; d.y = ny * inv; d.z = nz * inv;
}

// ... for a matrix whose z row and column are (0, 0, e): I3CLSimVectorTransformMatrix.cxx:101-133 with the zeros left out
__device__ __forceinline__ void apply_block_matrix(const float *M, float2 &dxy, float &dz)
{
    // (scalar on purpose: the constants sit in uniform registers, packing them into register pairs costs more than it saves)
    const float nx = fmaf(M[1], dxy.y, M[0] * dxy.x), ny = fmaf(M[4], dxy.y, M[3] * dxy.x), nz = M[8] * dz;
    const float inv = mufu_rsqrt(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
    dxy = __fmul2_rn(make_float2(nx, ny), make_float2(inv, inv));
    dz = nz * inv;
}

// ---- R6: DOM collision, restated for SIMT -----------------------------------------------------
// String level (sparse_collision_kernel.c.cl:27-192): cylinder pre-test with the global maximum
// radius (quirk 7), vertical extent, then the z-layer table of the string's set, at most one DOM
// per layer, ray-sphere test with the pancake factor; the closest entry point wins.
struct Collision {
    float travel;
    int string, dom;
    bool hit;
};
// StopDetectedPhotons = false (sparse_collision_kernel.c.cl:166-187): every DOM whose entry point lies on the leg is
// recorded and the photon flies on.  The slow phase walks the leg's hits in order of their entry points: the search
// below returns the nearest hit BEYOND a lower bound (entry distance, then string and DOM as tie-breakers), and is
// called again with the hit it returned as the new bound until nothing is left.  Each DOM has one entry point, so
// none is recorded twice -- the reference's bit masks (:85-104, 250-269) without their aliasing.
struct After {
    float entry;
    int key;   // string << 16 | DOM of the hit that was returned last; -1: none yet
};

template <bool NONSTOP>
__device__ __forceinline__ void test_string(const SmemPlan &sp, const DevGeometry &geo, int s, const V3 &pos, const V3 &dir,
                                            float inv_xy2, float inv_pancake, Collision &c, After after)
{
    const float4 sv = sp.strings[s];
    const float cross = (pos.x - sv.x) * dir.y - (pos.y - sv.y) * dir.x;
    if (cross * cross * inv_xy2 > geo.string_max_radius * geo.string_max_radius) return;
    if ((dir.z > 0.f) && (pos.z > sv.z)) return;
    if ((dir.z < 0.f) && (pos.z < sv.w)) return;
    const float4 set = sp.sets[sp.string_set[s]];
    const int nl = static_cast<int>(set.z);
    const int l0 = __float2int_rz((pos.z - set.x) * set.y);
    const int l1 = __float2int_rz((pos.z + dir.z * c.travel - set.x) * set.y);
    const int la = min(max(min(l0, l1), 0), nl - 1), lb = min(max(max(l0, l1), 0), nl - 1);
    const uint16_t *row = sp.layer_to_dom + static_cast<int>(set.w);
    const float r_om2 = geo.om_radius * geo.om_radius;
    for (int l = la; l <= lb; ++l) {
        const int dom = row[l];
        if (dom == 0xFFFF) continue;
        float qx, qy, qz;
        dom_centre(geo, s, dom, qx, qy, qz);
        const float rx = qx - pos.x, ry = qy - pos.y, rz = qz - pos.z;
        const float along = rx * dir.x + ry * dir.y + rz * dir.z;
        float disc = along * along - (rx * rx + ry * ry + rz * rz) + r_om2;
        if (disc < 0.f) continue;
        disc = mufu_sqrt(disc) * inv_pancake;
        const float entry = along - disc;
        if (entry < 0.f) continue; // started inside (or behind): let it leave (quirk 9)
        if (NONSTOP && (entry < after.entry || (entry == after.entry && ((s << 16) | dom) <= after.key))) continue;   // already recorded
        if (entry < c.travel || (NONSTOP && c.hit && entry == c.travel && ((s << 16) | dom) < ((c.string << 16) | c.dom))) {
            c.travel = entry;
            c.hit = true;
            c.string = s;
            c.dom = dom;
        }
    }
}

// The collision test proper, for the few segments the pixel map cannot rule out.  `who` >= 0:
// only that string can be reached (string level of the reference); `who` < 0: the reference's
// walk over the xy cells covered by the segment (sparse_collision_kernel.c.cl:194-303, 305-460).
template <bool NONSTOP>
__device__ __noinline__ Collision collide(const DevScene *scene, int who, V3 pos, V3 dir, float travel, After after = After{-1.f, -1})
{
    const SmemLayout &lay = reinterpret_cast<const SmemHeader *>(smem_base())->lay;
    const SmemPlan sp = table_plan(lay);
    const DevGeometry &geo = scene->geo;
    Collision c{travel, 0, 0, false};
    const float dir_xy2 = dir.x * dir.x + dir.y * dir.y;
    if (!(dir_xy2 > 0.f)) return c; // sparse_collision_kernel.c.cl:511-512
    const float inv_xy2 = mufu_rcp(dir_xy2);
    const float inv_pancake = scene->inv_pancake_factor;
    if (who >= 0) {
        test_string<NONSTOP>(sp, geo, who, pos, dir, inv_xy2, inv_pancake, c, after);
        return c;
    }
    for (int gI = 0; gI < geo.num_grids; ++gI) {
        const DevCellGrid &cg = geo.grids[gI];
        const float ex = pos.x + dir.x * c.travel, ey = pos.y + dir.y * c.travel;
        const int x0 = __float2int_rz((pos.x - cg.start_x) * cg.inv_width_x), x1 = __float2int_rz((ex - cg.start_x) * cg.inv_width_x);
        const int y0 = __float2int_rz((pos.y - cg.start_y) * cg.inv_width_y), y1 = __float2int_rz((ey - cg.start_y) * cg.inv_width_y);
        const int xa = min(max(min(x0, x1), 0), cg.num_x - 1), xb = min(max(max(x0, x1), 0), cg.num_x - 1);
        const int ya = min(max(min(y0, y1), 0), cg.num_y - 1), yb = min(max(max(y0, y1), 0), cg.num_y - 1);
        const uint16_t *cells = sp.cells + lay.cell_offset[gI];
        for (int cy = ya; cy <= yb; ++cy) {
            for (int cx = xa; cx <= xb; ++cx) {
                const int s = cells[cy * cg.num_x + cx];
                if (s == 0xFFFF) continue;
                // (a string that sits in several cells is simply tested again: the search is for a minimum)
                test_string<NONSTOP>(sp, geo, s, pos, dir, inv_xy2, inv_pancake, c, after);
            }
        }
    }
    return c;
}

// ---- R3: photon creation ------------------------------------------------------------------------
// createPhotonFromTrack (propagation_kernel.c.cl:132-184).  Deterministic in (step, creation-stream
// state): the queue fill calls it to make a photon, the hit path calls it again with the photon's birth
// tag to get the start-of-flight record back.
struct StepView {
    float x, y, z, t, length, beta;
    uint32_t source;
    V3 axis; // direction of travel (propagation_kernel.c.cl:482-489)
};
struct Born {
    V3 pos;
    float t;
    V3 dir;
    float wlen, life;
    float first_sample;   // table-maker variant
};

__device__ __forceinline__ V3 step_axis(float theta, float phi)
{
    float sth, cth, sph, cph;
    __sincosf(theta, &sth, &cth);
    __sincosf(phi, &sph, &cph);
    return V3{sth * cph, sth * sph, cth};
}

__device__ __forceinline__ Born create_core(const DevScene *scene, const StepView &s, Mwc &rng, float tab_step = 0.f)
{
    const DevMedium &m = scene->medium;
    Born b;
    const float shift = s.length * rng.co();
    b.dir = s.axis;
    if (scene->num_generators <= 1 || s.source == 0) {
        const SmemLayout &lay = reinterpret_cast<const SmemHeader *>(smem_base())->lay;
        b.wlen = draw_wavelength(scene->generators[0], rng, lay.gen0_n ? reinterpret_cast<const float *>(smem_base() + lay.off_gen0) : nullptr,
                                 (CLSIMCU_WLEN_BIN_RECORDS && lay.gen0_n) ? reinterpret_cast<const float *>(smem_base() + lay.off_gen0_bins) : nullptr);
        const float cos_c = fminf(1.f, mufu_rcp(s.beta * phase_index(m, b.wlen)));
        const float sin_c = mufu_sqrt(1.f - cos_c * cos_c);
        rotate_by(cos_c, sin_c, b.dir, rng.co());
    } else {
        b.wlen = (s.source < static_cast<uint32_t>(scene->num_generators)) ? draw_wavelength(scene->generators[s.source], rng) : 0.f;
    }
    // (table-maker variant: the offset of the first sampling point is drawn here, propagation_kernel.c.cl:566-569)
    b.first_sample = tab_step > 0.f ? tab_step * rng.oc() : 0.f;
    b.life = scene->fixed_abs ? scene->fixed_abs_lens : -fast_ln(rng.oc());
    b.pos = V3{s.x + s.axis.x * shift, s.y + s.axis.y * shift, s.z + s.axis.z * shift};
    b.t = s.t + shift * mufu_rcp(kSpeedOfLight * s.beta);
    return b;
}

// Queue fill: one photon of the warp's current step into queue slot `slot` (four 16-byte chunks), together
// with the wavelength-only factors of the ice model (R4, …_Optimizers.cxx:123-250).  The scene is read
// from its copy in global memory (uniform addresses).  Returns the advanced creation-stream state.
__device__ __noinline__ uint64_t create_photon(const DevScene *scene, const uint32_t *wstep, float4 *slot, uint64_t rng_x, uint32_t rng_a,
                                               uint32_t tag_step)
{
    const DevMedium &m = scene->medium;
    Mwc rng{rng_x, rng_a};
    StepView s;
    s.x = __uint_as_float(wstep[0]); s.y = __uint_as_float(wstep[1]); s.z = __uint_as_float(wstep[2]); s.t = __uint_as_float(wstep[3]);
    s.length = __uint_as_float(wstep[6]); s.beta = __uint_as_float(wstep[7]);
    s.source = wstep[11] & 0xffu;
    s.axis = V3{__uint_as_float(wstep[12]), __uint_as_float(wstep[13]), __uint_as_float(wstep[14])};
    const LaunchArgs &largs = reinterpret_cast<const SmemHeader *>(smem_base())->args;
    const float tab_step = largs.tabulate ? largs.tabulate->step_length : 0.f;
    const Born b = create_core(scene, s, rng, tab_step);
    // derived here, where all 32 lanes work, rather than when a lane takes the photon
    const int layer = min(max(__float2int_rz((b.pos.z - m.z0) * m.inv_h), 0), m.num_layers - 1);
    const float nm = b.wlen * 1e9f;
    const float f_scat = fast_pow(b.wlen * m.inv_ref_wlen, -m.alpha);              // 1/scatLen = b400 * this
    const float f_dust = fast_pow(nm, -m.kappa);                                  // dust term factor
    const float f_pure = m.A * mufu_ex2(-m.B * mufu_rcp(nm) * 1.44269504089f);    // pure-ice term
    slot[0] = make_float4(b.pos.x, b.pos.y, b.dir.x, b.dir.y);
    slot[1] = make_float4(b.pos.z, b.dir.z, b.life, __int_as_float(layer));
    slot[2] = make_float4(f_scat, f_pure, f_dust, 0.f);
    if (largs.tabulate)
        slot[3] = make_float4(b.t, inv_group_velocity(m, b.wlen), __uint_as_float(wstep[9]), b.first_sample);   // (time, 1/v_g, step weight, first sample)
    else
        slot[3] = make_float4(__uint_as_float(static_cast<uint32_t>(rng_x)), __uint_as_float(static_cast<uint32_t>(rng_x >> 32)), __uint_as_float(tag_step), 0.f);
    return rng.x;
}

// ---- R10: hit output ----------------------------------------------------------------------------
// Called by the lanes whose photon was just detected (or absorbed, in save-all mode): `pos` is the end
// point, `path` the full path length.  The photon's start-of-flight record is re-created from its birth
// tag into the lane's state words `st` (the photon is over, its state is dead).  Reservation is
// aggregated over the lanes that arrive together; the record leaves as five 16-byte stores.
__device__ __noinline__ void emit_record(const DevScene *scene_dev, float *st, V3 pos, V3 dir, float path, uint32_t scatters,
                                         int hit_string, int hit_dom, float abs_left_at_end, bool save_all, uint32_t rng_a)
{
    const LaunchArgs &args = reinterpret_cast<const SmemHeader *>(smem_base())->args;
    const DevScene &scene = *scene_dev;
    const unsigned peers = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(args.hit_counter, static_cast<uint32_t>(__popc(peers)));
    base = __shfl_sync(peers, base, leader);
    const uint32_t slot = base + __popc(peers & ((1u << lane) - 1u));
    if (slot >= args.max_hits) return; // counted but dropped (quirk 10)

    // birth tag -> start-of-flight record
    const uint32_t tag_step = tag_word(st, 2);
    const uint32_t step_index = tag_step & ((1u << kStepIndexBits) - 1u);
    const uint32_t made_by = (blockIdx.x * kThreads + (threadIdx.x & ~31u)) + (tag_step >> kStepIndexBits);
    const uint64_t birth_x = static_cast<uint64_t>(tag_word(st, 0)) | (static_cast<uint64_t>(tag_word(st, 1)) << 32);
    const uint32_t birth_a = __ldg(args.rng_a + args.rng_creation_offset + made_by);
    const clsimcu_step *step = static_cast<const clsimcu_step *>(args.steps) + step_index;
    StepView sv;
    sv.x = __ldg(&step->x); sv.y = __ldg(&step->y); sv.z = __ldg(&step->z); sv.t = __ldg(&step->t);
    sv.length = __ldg(&step->length); sv.beta = __ldg(&step->beta);
    sv.source = __ldg(reinterpret_cast<const uint32_t *>(step) + 11) & 0xffu;
    sv.axis = step_axis(__ldg(&step->theta), __ldg(&step->phi));
    Mwc birth{birth_x, birth_a};
    const Born born = create_core(scene_dev, sv, birth);

    const DevGeometry &geo = scene.geo;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    uint32_t ids = 0;
    if (!save_all) {
        dom_centre(geo, hit_string, hit_dom, qx, qy, qz);
        if (scene.pancake) {
            // undo the pancake: shift the DOM centre along the component of (pos - dom) perpendicular
            // to the direction (propagation_kernel.c.cl:340-355); that component is the same
            // anywhere along the ray, so the end point can be used
            const float px = pos.x - qx, py = pos.y - qy, pz = pos.z - qz;
            const float along = px * dir.x + py * dir.y + pz * dir.z;
            const float k = (scene.pancake_factor - 1.f) * scene.inv_pancake_factor;
            qx += k * (px - along * dir.x); qy += k * (py - along * dir.y); qz += k * (pz - along * dir.z);
        }
        const int16_t sid = __ldg(geo.string_index_to_id + hit_string);
        const uint16_t oid = __ldg(geo.dom_ids + __ldg(geo.dom_id_offset + hit_string) + hit_dom);
        ids = static_cast<uint32_t>(static_cast<uint16_t>(sid)) | (static_cast<uint32_t>(oid) << 16);
    }
    const float ivg = inv_group_velocity(scene.medium, born.wlen);
    float th, ph, sth, sph;
    to_spherical(dir.x, dir.y, dir.z, th, ph);
    to_spherical(born.dir.x, born.dir.y, born.dir.z, sth, sph);
    float4 *dst = reinterpret_cast<float4 *>(static_cast<clsimcu_photon *>(args.photons) + slot);
    dst[0] = make_float4(pos.x - qx, pos.y - qy, pos.z - qz, born.t + path * ivg);
    dst[1] = make_float4(th, ph, born.wlen, path);
    dst[2] = make_float4(__uint_as_float(scatters), __ldg(&step->weight) / bias_at(scene.bias, born.wlen),
                         __uint_as_float(__ldg(&step->identifier)), __uint_as_float(ids));
    dst[3] = make_float4(born.pos.x, born.pos.y, born.pos.z, born.t);
    dst[4] = make_float4(sth, sph, 1.f / ivg, born.life - abs_left_at_end);
    if (args.history && args.history_ring) {
        // the lane's ring as it is (the reference writes the raw ring too, propagation_kernel.c.cl:387-392; the host
        // puts it in order, ...OpenCL.cxx:940-989); fourth component: absorption lengths used so far
        const int n = scene.history_entries;
        const float4 *ring = reinterpret_cast<const float4 *>(args.history_ring) + (blockIdx.x * kThreads + threadIdx.x);
        float4 *out = reinterpret_cast<float4 *>(args.history) + static_cast<size_t>(slot) * n;
        for (int i = 0; i < n; ++i) {
            float4 e = ring[static_cast<size_t>(i) * (static_cast<size_t>(gridDim.x) * kThreads)];
            e.w = born.life - e.w;
            out[i] = e;
        }
    }
    if (save_all && args.rng_tag_x) {
        const uint32_t *ptag = reinterpret_cast<const uint32_t *>(st + kOffPopTag);
        args.rng_tag_x[2 * static_cast<size_t>(slot)] = birth_x;
        args.rng_tag_x[2 * static_cast<size_t>(slot) + 1] = static_cast<uint64_t>(ptag[0 * kThreads]) | (static_cast<uint64_t>(ptag[1 * kThreads]) << 32);
        args.rng_tag_a[2 * static_cast<size_t>(slot)] = birth_a;
        args.rng_tag_a[2 * static_cast<size_t>(slot) + 1] = rng_a;
    }
}

// The rest of a flight that was cut short by a DOM: the reference reports distInAbsLens for the
// UNSHORTENED segment (propagation_kernel.c.cl:718), i.e. the absorption budget left where the
// flight would have ended.  Only hit photons whose flight was not going to end in this layer
// come here, so the loop the main path avoids is acceptable.
__device__ __noinline__ float abs_left_at_end_of_flight(const float4 *layers, float z0, float h, int num_layers, int layer, float z,
                                                        float dz, float inv_dz, float sca_left, float abs_left, float f_scat,
                                                        float f_dust, float f_pure)
{
    const bool up = !(dz < 0.f);
    for (;;) {
        const float4 c = layers[layer];
        const float b = c.x * f_scat, a = c.z * f_dust + c.y * f_pure;
        const float zb = z0 + h * static_cast<float>(layer + (up ? 1 : 0));
        const float d_b = fmaxf((zb - z) * inv_dz, 0.f);
        const bool can_cross = up ? (layer < num_layers - 1) : (layer > 0);
        const bool absorbed = abs_left * b < sca_left * a;
        const float d_sa = (absorbed ? abs_left : sca_left) * mufu_rcp(absorbed ? a : b);
        if (!(can_cross && d_b < d_sa)) return absorbed ? 0.f : abs_left - d_sa * a;
        sca_left = fmaxf(sca_left - d_b * b, 1e-30f);
        abs_left -= d_b * a;
        z = zb;
        layer += up ? 1 : -1;
    }
}

// ---- the photon in flight ---------------------------------------------------------------------
// Quantities that are always used together sit in register PAIRS (float2): sm_100a has packed fp32 instructions
// (FFMA2 / FMUL2 / FADD2: two independent fp32 operations per issue slot, with scalar-broadcast, swap and per-half
// negation of the operands for free), and this kernel is bound by instruction issue, not by the fp32 pipes.
struct Lane {
    float2 pxy;                       // position x, y
    float pz;
    float2 dxy;                       // direction x, y
    float dz;
    float inv_dz;                     // 1/dz (+-inf for dz == 0), refreshed whenever the direction changes
    float2 bud;                       // (abs_left, sca_left): budgets in absorption / scattering lengths; sca_left == 0 marks "draw a new flight"
    float path;
    float2 f_sp;                      // (f_scat, f_pure): wavelength-only factors of the ice model
    float f_dust;
    float z_eff;                      // TILT only: z in the untilted layer frame, carried along the flight
    float inv_aniso;                  // ANISO only: abs_left is held scaled by 1/inv_aniso during a flight
    uint32_t scatters;
    int layer;
    uint32_t status;
    uint64_t rng_x;
    // table-maker variant only (TAB): time at the start of the flight's current leg, 1 / group velocity, the step's
    // weight, distance from the leg's start to the next sampling point of the path
    float t, inv_vg, weight, rem;
};

// 1/dz without a guard: dz == 0 gives +-inf, and a leg then never reaches the boundary ahead (the reference treats
// |dz| < 1e-5 as "stays in its layer", propagation_kernel.c.cl:669; a photon that flat flies > 1e5 layer heights per layer)
__device__ __forceinline__ float raw_inv_dz(float dz) { return mufu_rcp(dz); }

template <bool TILT, bool ANISO, bool TAB = false> __device__ __forceinline__ void load_lane(Lane &L, const float *st)
{
    if (TAB) {
        // (table mode records no photons: the birth-tag words and the parking words hold the four extra values)
        L.t = st[(kOffBirthTag + 0 * kThreads)]; L.inv_vg = st[(kOffBirthTag + 1 * kThreads)]; L.weight = st[(kOffBirthTag + 2 * kThreads)];
        L.rem = st[kPendTravel * kThreads];
    }
    L.pxy = make_float2(st[kPx * kThreads], st[kPy * kThreads]); L.pz = st[kPz * kThreads];
    L.dxy = make_float2(st[kDx * kThreads], st[kDy * kThreads]); L.dz = st[kDz * kThreads];
    L.inv_dz = raw_inv_dz(L.dz);
    L.bud = make_float2(st[kAbsLeft * kThreads], st[kScaLeft * kThreads]); L.path = st[kPath * kThreads];
    L.f_sp = make_float2(st[kFScat * kThreads], st[kFPure * kThreads]); L.f_dust = st[kFDust * kThreads];
    L.scatters = __float_as_uint(st[kScatters * kThreads]);
    L.layer = __float_as_int(st[kLayer * kThreads]);
    L.status = __float_as_uint(st[kStatus * kThreads]);
    L.rng_x = pack64(st[kRngLo * kThreads], st[kRngHi * kThreads]);
    L.z_eff = TILT ? st[kZEff * kThreads] : 0.f;
    L.inv_aniso = ANISO ? st[kInvAniso * kThreads] : 1.f;
}

template <bool TILT, bool ANISO, bool TAB = false> __device__ __forceinline__ void store_lane(const Lane &L, float *st)
{
    if (TAB) {
        st[(kOffBirthTag + 0 * kThreads)] = L.t; st[(kOffBirthTag + 1 * kThreads)] = L.inv_vg; st[(kOffBirthTag + 2 * kThreads)] = L.weight;
        st[kPendTravel * kThreads] = L.rem;
    }
    st[kPx * kThreads] = L.pxy.x; st[kPy * kThreads] = L.pxy.y; st[kPz * kThreads] = L.pz;
    st[kDx * kThreads] = L.dxy.x; st[kDy * kThreads] = L.dxy.y; st[kDz * kThreads] = L.dz;
    st[kAbsLeft * kThreads] = L.bud.x; st[kScaLeft * kThreads] = L.bud.y; st[kPath * kThreads] = L.path;
    st[kFScat * kThreads] = L.f_sp.x; st[kFDust * kThreads] = L.f_dust; st[kFPure * kThreads] = L.f_sp.y; // lanes change photons inside a fast phase
    st[kScatters * kThreads] = __uint_as_float(L.scatters);
    st[kLayer * kThreads] = __int_as_float(L.layer);
    st[kStatus * kThreads] = __uint_as_float(L.status);
    st[kRngLo * kThreads] = __uint_as_float(static_cast<uint32_t>(L.rng_x));
    st[kRngHi * kThreads] = __uint_as_float(static_cast<uint32_t>(L.rng_x >> 32));
    if (TILT) st[kZEff * kThreads] = L.z_eff;
    if (ANISO) st[kInvAniso * kThreads] = L.inv_aniso;
}

// Second look at a leg the 2-D cylinder test could not rule out, still in the hot loop (one leg in ~600 gets here): the
// reference's string test (sparse_collision_kernel.c.cl:107-192) with approximate arithmetic and margins, deciding only
// whether the leg MAY enter a DOM of string `who`.  The z-layer table of the string's set names a DOM in every layer its
// sphere touches, so the layers the leg covers name the candidates; a candidate counts when the leg reaches its sphere
// (oversized, flattened by the pancake factor) from outside, with margins that cover the rounding of the approximate root.
// Nine in ten of the legs that a mere "a DOM is near in z" test parked were false alarms (hits are one leg in 30 000);
// photons that start inside a DOM (flashers) leave it without a visit to the slow phase.
#ifdef CLSIMCU_Z_RANGE_PRETEST
__device__ __noinline__ bool dom_within_reach(const DevScene *scene, int who, float z, float dz, float travel, float t, float dxy2, float out2)
{
    const SmemPlan sp = table_plan(reinterpret_cast<const SmemHeader *>(smem_base())->lay);
    const DevGeometry &geo = scene->geo;
    float s0 = 0.f, s1 = travel, slack = 0.01f;
    if (dxy2 > 1e-12f) {
        const float inv = mufu_rcp(dxy2);
        const float sq = mufu_sqrt(fmaxf(fmaf(t, t, -dxy2 * out2), 0.f));
        s0 = fmaxf((t - sq) * inv, 0.f);
        s1 = fminf((t + sq) * inv, travel);
        slack = fmaf(5e-4f * inv, fabsf(t * dz), 0.01f);
        if (s0 > s1 + slack) return false;
    }
    const float za = fmaf(dz, s0, z), zb = fmaf(dz, s1, z);
    const float zlo = fminf(za, zb) - slack, zhi = fmaxf(za, zb) + slack;
    const float4 set = sp.sets[sp.string_set[who]];
    const int nl = static_cast<int>(set.z);
    const int la = min(max(__float2int_rd((zlo - set.x) * set.y), 0), nl - 1), lb = min(max(__float2int_rd((zhi - set.x) * set.y), 0), nl - 1);
    const uint16_t *row = sp.layer_to_dom + static_cast<int>(set.w);
    const uint32_t first = __ldg(geo.string_tmpl_start + who);
    for (int l = la; l <= lb; ++l) {
        const int dom = row[l];
        if (dom == 0xFFFF) continue;
        const float zd = __ldg(geo.tmpl_z + first + static_cast<uint32_t>(dom));
        if ((zd >= zlo - geo.om_radius) && (zd <= zhi + geo.om_radius)) return true;
    }
    return false;
}
#endif
__device__ __noinline__ bool dom_may_be_hit(const DevScene *scene, int who, float px, float py, float pz, float dx, float dy, float dz, float travel)
{
    const SmemPlan sp = table_plan(reinterpret_cast<const SmemHeader *>(smem_base())->lay);
    const DevGeometry &geo = scene->geo;
    const float4 set = sp.sets[sp.string_set[who]];
    const int nl = static_cast<int>(set.z);
    const float margin = 2e-3f;   // metres; far above the rounding of a root of a number of the size of (leg length)^2
    const float z_end = fmaf(dz, travel + margin, pz);
    const int l0 = __float2int_rz((fminf(pz, z_end) - margin - set.x) * set.y), l1 = __float2int_rz((fmaxf(pz, z_end) + margin - set.x) * set.y);
    const int la = min(max(l0, 0), nl - 1), lb = min(max(l1, 0), nl - 1);
    const uint16_t *row = sp.layer_to_dom + static_cast<int>(set.w);
    const float r_om2 = geo.om_radius * geo.om_radius;
    for (int l = la; l <= lb; ++l) {
        const int dom = row[l];
        if (dom == 0xFFFF) continue;
        float qx, qy, qz;
        dom_centre(geo, who, dom, qx, qy, qz);
        const float rx = qx - px, ry = qy - py, rz = qz - pz;
        const float along = fmaf(rx, dx, fmaf(ry, dy, rz * dz));
        const float r2 = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
        const float disc = fmaf(along, along, r_om2 - r2);
        const float tol = fmaf(1e-5f, r2, 1e-4f);                     // cancellation in along^2 - r^2
        if (disc < -tol) continue;                                     // the line misses the sphere
        const float entry = along - mufu_sqrt(fmaxf(disc, 0.f)) * scene->inv_pancake_factor;
        // the reference lets a photon that starts inside (or beyond) a sphere leave it: entry < 0 is no hit (quirk 9)
        if (entry < -margin - tol || entry > travel + margin + tol) continue;
        return true;
    }
    return false;
}

// The part of an iteration that decides where the photon goes next: the event that ends this leg
// of the flight (scatter / absorption inside the current ice layer, the layer boundary, or the
// range limit of the collision map) and what is left of the two budgets afterwards.
struct Leg {
    float2 q;                 // (b, a): 1/scattering length, 1/absorption length in the current layer
    float zb, zc;             // boundary ahead, current z (both in the layer frame)
    float travel;
    float2 rem;               // budgets (abs, sca) left at the end of the leg
    float2 o;                 // from the photon to the axis of the nearest string (xy)
    float d_b, cap;           // distance to the layer boundary ahead; range limit of the collision map (+inf: none)
    uint32_t cell;            // pixel-map word: byte offset of that string's record | range bits
    bool absorbed, limited;
    bool looked;              // the collision map was consulted for this leg (o, cap and cell are set)
};

// +1 for a photon going up (or flat), -1 for one going down, from the sign of 1/dz
__device__ __forceinline__ int layer_step(float inv_dz) { return (__float_as_int(inv_dz) >> 31) | 1; }
// ... as a step of the layer record's address
__device__ __forceinline__ int layer_address_step(float inv_dz) { return ((__float_as_int(inv_dz) >> 31) & -32) + 16; }
// pixel-map word -> string index / "the strings are too dense here for the map: the reference's cell walk instead"
__device__ __forceinline__ int cell_string(uint32_t cell) { return static_cast<int>((cell & 0xffffu) >> 4); }
__device__ __forceinline__ bool cell_walk(uint32_t cell) { return (cell & 0xffff0000u) == 0x7f800000u; }

// Loads from 32-bit shared-memory addresses.  The hot loop holds the layer table and the pixel map as such addresses: the
// compiler re-derives a generic pointer's shared address (S2UR / ULEA, or a base register it then spills) at every use.
__device__ __forceinline__ float4 lds128(uint32_t a) { float4 v; asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float lds32f(uint32_t a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
// fire-and-forget float add to global memory (RED.E.ADD.F32)
__device__ __forceinline__ void red_add(float *p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}
// the shared-memory address of dynamic shared memory's first byte, opaque to the optimiser (so that it is computed once)
__device__ __forceinline__ uint32_t smem_address()
{
    uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_base()));
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    return a;
}

// LOOK_ALWAYS legs consult the collision map; the others are planned without it (no range limit) and may only be
// flown if they are shorter than the photon's clearance (see advance_photon).
// `L.layer` is the shared-memory ADDRESS of the photon's layer record (one LDS.128 and one LDS with no address arithmetic;
// a layer change is +-16); `near` the address of the pixel map.
template <bool TILT, bool SAVE_ALL, bool LOOK_ALWAYS>
__device__ __forceinline__ Leg plan_leg(const Lane &L, const DevScene &scene, const float4 *strings, uint32_t near)
{
    const DevGeometry &geo = scene.geo;
    Leg g;
    const float4 c = lds128(static_cast<uint32_t>(L.layer));
    const float z_up = lds32f(static_cast<uint32_t>(L.layer) + 28u);
    g.q = __fmul2_rn(make_float2(c.x, c.y), L.f_sp);         // (b400 * f_scat, (1 + 0.01 dTau) * f_pure)
    g.q.y = fmaf(c.z, L.f_dust, g.q.y);
    g.zb = (L.inv_dz < 0.f) ? c.w : z_up;                    // +-1e30 when there is no layer beyond
    g.zc = TILT ? L.z_eff : L.pz;
    g.d_b = fmaxf((g.zb - g.zc) * L.inv_dz, 0.f);            // (0 * inf = NaN -> 0: a flat photon on a boundary steps over it)
    // (the budgets are written by predicated code in several places; ptxas does not keep them in an aligned register
    // pair, and packed instructions on them cost more in moves than they save -- measured)
#ifdef CLSIMCU_PACKED_BUDGETS
    const float2 cmp = __fmul2_rn(L.bud, g.q);               // (abs_left * b, sca_left * a)
    g.absorbed = cmp.x < cmp.y;                              // d_absorb < d_scatter inside this layer
#else
    g.absorbed = L.bud.x * g.q.x < L.bud.y * g.q.y;          // d_absorb < d_scatter inside this layer
#endif
    const float d_sa = (g.absorbed ? L.bud.x : L.bud.y) * mufu_rcp(g.absorbed ? g.q.y : g.q.x);
    // pixel map: nearest string and the range within which no other string can be touched
    g.cell = 0u;
    g.cap = __int_as_float(0x7f800000);
    g.o = make_float2(0.f, 0.f);
    g.looked = !SAVE_ALL && LOOK_ALWAYS;
    if (g.looked) {
        // (float -> unsigned conversion saturates below at 0)
        const uint32_t px = min(__float2uint_rz(fmaf(L.pxy.x, geo.near_inv_pixel, geo.near_off_x)), static_cast<uint32_t>(geo.near_nx - 1));
        const uint32_t py = min(__float2uint_rz(fmaf(L.pxy.y, geo.near_inv_pixel, geo.near_off_y)), static_cast<uint32_t>(geo.near_ny - 1));
        g.cell = lds32(near + 4u * (py * geo.near_nx + px));
        g.cap = __uint_as_float(g.cell & 0xffff0000u);       // +inf: no limit (and every leg takes the reference's cell walk, see cell_walk)
        const float2 sxy = *reinterpret_cast<const float2 *>(reinterpret_cast<const uint8_t *>(strings) + (g.cell & 0xffffu));
        g.o = __fadd2_rn(sxy, make_float2(-L.pxy.x, -L.pxy.y));
    }
    const float d_geo = g.looked ? fminf(g.d_b, g.cap) : g.d_b;
    g.limited = d_geo < d_sa;                                // the flight goes on after this leg
    g.travel = fminf(d_geo, d_sa);
#ifdef CLSIMCU_PACKED_BUDGETS
    g.rem = __ffma2_rn(make_float2(-g.travel, -g.travel), make_float2(g.q.y, g.q.x), L.bud);
    g.rem.y = fmaxf(g.rem.y, 1e-30f);
#else
    g.rem.x = fmaf(-g.travel, g.q.y, L.bud.x);
    g.rem.y = fmaxf(fmaf(-g.travel, g.q.x, L.bud.y), 1e-30f);
#endif
    return g;
}

// R8 (propagation_kernel.c.cl:83-129) on a direction held as (xy pair, z); the azimuth comes as the raw 32-bit
// draw, so that its scaling to [0, 2 pi) is one multiplication (bit-identical to 2 pi * (draw * 2^-32): the power of
// two is exact).  With k = sin(a)/sin(theta), u = cos(b) k, w = cos(a) - z sin(b) k, m = sin(a) sin(theta):
//     (x, y)' = (x, y) w + (-y, x) u        z' = z cos(a) + m sin(b)
// sin^2(theta) is taken from x and y, not as 1 - z^2: for a direction whose length is off by eps the rotation then
// gives a length off by at most eps again (with 1 - z^2 the error is amplified by sin^2(a)/sin^2(theta) near the
// poles), so the length only random-walks by rounding and is restored once per fast phase, not per scatter.
// The scattering angle comes as its cosine and sin^2 (in [0, 1]); k and m take ONE reciprocal root:
//     r = 1/sqrt(sin^2(a) sin^2(theta)),   k = sin^2(a) r,   m = sin^2(a) sin^2(theta) r
// (the special-function unit, 8 cycles per warp instruction and scheduler, is the next bound after issue).
__device__ __forceinline__ void rotate_packed(float cosa, float sina2, float2 &dxy, float &dz, uint32_t draw)
{
    float sinb, cosb;
    __sincosf(__uint2float_rz(draw) * (2.0f * kPi * 2.3283064365386963e-10f), &sinb, &cosb);
    const float s2 = fmaf(dxy.x, dxy.x, dxy.y * dxy.y);
    if (s2 > 1e-20f) {
        // (sin^2(a) is 0 or at least an ulp of 1, so the product is 0 or a normal number; 0 gives k = m = 0)
        const float x = sina2 * s2;
        const float r = mufu_rsqrt(fmaxf(x, 1e-36f));
        const float2 km = __fmul2_rn(make_float2(sina2, x), make_float2(r, r));   // (k, m)
        const float u = cosb * km.x;
        const float w = fmaf(-dz * sinb, km.x, cosa);
        const float nz = fmaf(sinb, km.y, dz * cosa);
        const float2 along = __fmul2_rn(dxy, make_float2(w, w));
        dxy = __ffma2_rn(make_float2(-dxy.y, dxy.x), make_float2(u, u), along);
        dz = nz;
    } else {
        // within 1e-10 rad of a pole
        const float sina = mufu_sqrt(sina2);
        dxy = make_float2(sina * cosb, sina * sinb);
        dz = (dz > 0.f) ? cosa : ((dz < 0.f) ? -cosa : cosa * dz);
    }
}

// R5, first part: a new flight (after creation or a scatter) draws its length in scattering lengths and, where the ice
// is tilted or anisotropic, takes the layer and the absorption scaling of its direction (propagation_kernel.c.cl:599-631).
template <bool TILT, bool ANISO>
__device__ __forceinline__ void start_flight(Lane &L, const DevMedium &m, uint32_t layers, const float2 *tilt_dist, const float4 *tilt_corr, uint32_t rng_a)
{
    if (L.bud.y <= 0.f) {
        Mwc rng{L.rng_x, rng_a};
        if (TILT) {
            L.z_eff = L.pz - tilt_shift(m, tilt_dist, tilt_corr, L.pxy.x, L.pxy.y, L.pz);
            L.layer = static_cast<int>(layers) + 16 * min(max(__float2int_rz((L.z_eff - m.z0) * m.inv_h), 0), m.num_layers - 1);
        }
        if (ANISO) {
            // R4b: 1/f = (B2-nB)*An/2 (I3CLSimScalarFieldAnisotropyAbsLenScaling.cxx:92-134)
            const float n0 = m.azx * L.dxy.x + m.azy * L.dxy.y, n1 = m.neg_azy * L.dxy.x + m.azx * L.dxy.y;
            const float s0 = n0 * n0, s1 = n1 * n1, s2 = L.dz * L.dz;
            const float nB = s0 * m.rl[0] + s1 * m.rl[1] + s2 * m.rl[2];
            const float An = s0 * m.l[0] + s1 * m.l[1] + s2 * m.l[2];
            L.inv_aniso = (m.B2 - nB) * An * 0.5f;
            L.bud.x *= mufu_rcp(L.inv_aniso);
        }
        L.bud.y = -fast_ln(rng.oc());
        L.rng_x = rng.x;
    }
}

// One iteration of the hot loop: move the photon to its next event.  A leg that might touch a DOM
// parks the lane (status kFrozen, the leg's length and string in the state words kPend*) with
// nothing but the scattering-length draw applied; the slow phase runs the full collision test
// and either ends the photon there or flies this leg itself (finish_leg) and sends the lane back.
template <bool TILT, bool ANISO, bool SAVE_ALL, bool MIXED, int V = 0>
__device__ __forceinline__ void finish_leg(Lane &L, const Leg &g, const DevScene &scene, uint32_t rng_a, float4 *ring);

template <bool TILT, bool ANISO, bool SAVE_ALL, bool MIXED, bool LOOK_ALWAYS, int V = 0>
__device__ __forceinline__ void advance_photon(Lane &L, float &clearance, const DevScene &scene, const DevScene *scene_dev, uint32_t layers,
                                               const float4 *strings, uint32_t near, const float2 *tilt_dist,
                                               const float4 *tilt_corr, uint32_t rng_a, float *st, float4 *ring)
{
    const DevMedium &m = scene.medium;

    start_flight<TILT, ANISO>(L, m, layers, tilt_dist, tilt_corr, rng_a);
    const Leg g = plan_leg<TILT, SAVE_ALL, LOOK_ALWAYS>(L, scene, strings, near);
    // `clearance`: how far the photon may still fly before any string can come within the collision radius, as
    // known from the lane's last look at the collision map minus what it has flown since.  The first leg after each
    // look at the warp state consults the map and sets it; the legs after that one do not look: a leg shorter than
    // the clearance needs neither the map nor a test nor a range limit, and a lane whose leg is not waits for the
    // next leg that looks (a few lanes in a hundred; cheaper than a divergent look-up in most warp-legs).
    if (!SAVE_ALL && !LOOK_ALWAYS && !(g.travel < clearance)) return;

    // ------------------------------------------------------------------ R6: DOM collision, cheap part
    if (g.looked) {
        // is the one string in range within reach of this leg at all?  (a few legs in a thousand; where the map
        // has no range every leg is)
        const float R = scene.geo.string_max_radius;
        const float o2 = fmaf(g.o.x, g.o.x, g.o.y * g.o.y);
        const float reach = g.travel + R;
        if (!(o2 > reach * reach)) {                             // (o2 is NaN where the map has no range)
            // 2-D segment / cylinder test
            const bool walk = cell_walk(g.cell);
            const float t = fmaf(g.o.x, L.dxy.x, g.o.y * L.dxy.y);
            const float dxy2 = fmaf(L.dxy.x, L.dxy.x, L.dxy.y * L.dxy.y);
            const float out2 = o2 - R * R;                       // > 0: the photon starts outside the cylinder
            const bool miss = (o2 > reach * reach) || ((t <= 0.f) && (out2 > 0.f)) || (out2 * dxy2 > t * t);
            if (!miss || walk) {
#ifdef CLSIMCU_DEBUG_COUNTERS
                {
                    const LaunchArgs &dbg = reinterpret_cast<const SmemHeader *>(smem_base())->args;
                    atomicAdd(dbg.stats + 2, 1ull);                                  // legs parked
                    if (walk) atomicAdd(dbg.stats + 3, 1ull);                        // ... in a dense-string pixel
                    if (out2 <= 0.f) atomicAdd(dbg.stats + 4, 1ull);                 // ... starting inside the cylinder
                }
#endif
#ifndef CLSIMCU_NO_Z_PRETEST
#ifdef CLSIMCU_Z_RANGE_PRETEST
                if (walk || dom_within_reach(scene_dev, cell_string(g.cell), L.pz, L.dz, g.travel, t, dxy2, out2))
#else
                if (walk || dom_may_be_hit(scene_dev, cell_string(g.cell), L.pxy.x, L.pxy.y, L.pz, L.dxy.x, L.dxy.y, L.dz, g.travel))
#endif
#endif
                {
                    L.status = kFrozen;
                    st[kPendTravel * kThreads] = g.travel;
                    st[kPendWho * kThreads] = __int_as_float(walk ? -1 : cell_string(g.cell));
                    return;
                }
            }
        }
        // from here: to the cylinder around the named string, or to where any other string comes into range (the
        // comparison fails for the NaN of a pixel without a range: no clearance there)
        clearance = (o2 > 0.f) ? fminf(mufu_sqrt(o2) - R, g.cap) - 0.01f : 0.f;
    }
    clearance -= g.travel;
    finish_leg<TILT, ANISO, SAVE_ALL, MIXED, V>(L, g, scene, rng_a, ring);
}

// The table-maker variant's sink (savePath, propagation_kernel.c.cl:226-304): every step_length metres along the leg
// the photon adds  weight x angular acceptance x exp(-absorption lengths used so far)  to the bin of the table its
// position and delay time fall into -- straight into the table in HBM (RED.ADD.F32), no entry buffers.  The
// absorption used at a sampling point is taken from the leg's own layer (the reference interpolates linearly over a
// whole multi-layer segment, :296-297; the same within a layer).  A point beyond the table's radius or delay-time
// range ends the photon.  Returns false when the photon is to be stopped.
template <bool ANISO>
__device__ __forceinline__ bool sample_leg(Lane &L, const Leg &g, bool flying, const DevScene &scene)
{
    // WARP-COOPERATIVE: called by all 32 lanes, converged.  Lane i has n_i sampling points on its leg (0 for a lane
    // without a flying photon); the warp's points are numbered through and handed out 32 at a time, so that a lane with
    // a long leg does not hold the others (a lane-private loop ran at 4 of 32 lanes: ncu r02_v34_tab).  A point needs
    // the twelve numbers of its leg: fetched from the owning lane by shuffle.
    const TabulateArgs &tb = reinterpret_cast<const SmemHeader *>(smem_base())->tab;
    const int lane = threadIdx.x & 31;
    const float step = tb.step_length;
    int n = 0;
    if (flying && L.rem < g.travel) {
        n = __float2int_rz((g.travel - L.rem) * mufu_rcp(step));
        n += (fmaf(static_cast<float>(n), step, L.rem) < g.travel) ? 1 : 0;          // d_k = rem + k step < travel for k < n
        n -= (n > 0 && !(fmaf(static_cast<float>(n - 1), step, L.rem) < g.travel)) ? 1 : 0;
    }
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - n;
    // per leg: weight x angular acceptance; absorption lengths used up to the start of the leg and per metre in this layer
    // (the budget of an anisotropic medium is held scaled during a flight, see start_flight)
    const float scale = ANISO ? L.inv_aniso : 1.f;
    const float impact_weight = flying ? L.weight * table_angular_acceptance(tb, L.dz) : 0.f;
    const float depth0 = scene.fixed_abs_lens - L.bud.x * scale;
    const float per_metre = flying ? g.q.y * scale : 0.f;
    unsigned stopped = 0u;
    for (int base = 0; base < total; base += 32) {
        const int mine = base + lane;
        // owner of point `mine`: the first lane whose inclusive count exceeds it
        int src = 0;
#pragma unroll
        for (int h = 16; h > 0; h >>= 1) {
            const int below = __shfl_sync(0xffffffffu, incl, src + h - 1);
            if (below <= mine) src += h;
        }
        src = min(src, 31);
        const bool valid = mine < total;
        const int k = mine - __shfl_sync(0xffffffffu, excl, src);
        const float d = fmaf(static_cast<float>(k), step, __shfl_sync(0xffffffffu, L.rem, src));
        const float x = fmaf(__shfl_sync(0xffffffffu, L.dxy.x, src), d, __shfl_sync(0xffffffffu, L.pxy.x, src));
        const float y = fmaf(__shfl_sync(0xffffffffu, L.dxy.y, src), d, __shfl_sync(0xffffffffu, L.pxy.y, src));
        const float z = fmaf(__shfl_sync(0xffffffffu, L.dz, src), d, __shfl_sync(0xffffffffu, L.pz, src));
        const float t = fmaf(__shfl_sync(0xffffffffu, L.inv_vg, src), d, __shfl_sync(0xffffffffu, L.t, src));
        const float w0 = __shfl_sync(0xffffffffu, impact_weight, src);
        const float depth = fmaf(__shfl_sync(0xffffffffu, per_metre, src), d, __shfl_sync(0xffffffffu, depth0, src));
        float c[4];
        table_coordinates_4<true>(tb, table_frame<true>(tb, x, y, z, t), c);
        // a point beyond the table's radius or delay-time range ends the photon (:770-776).  Both coordinates only grow
        // along a leg once they are out (the distance to the reference point is convex along a line, the photon is at
        // least as slow as the table's fastest light), so the points after it on the leg are out as well: none is added.
        const bool out = valid && table_out_of_bounds(tb, c);
        if (valid && !out) {
            const uint32_t index = table_bin_index_4<true>(tb, c);
            const float w = w0 * __expf(-depth);
            // (the table is in global memory: a reduction without a return value, not the generic-address atomic with its
            // address-space query and compare-and-swap fall-backs)
            red_add(tb.table + index, w);
            if (tb.squared) red_add(tb.squared + index, w * w);
        }
        stopped |= __reduce_or_sync(0xffffffffu, out ? (1u << src) : 0u);
    }
    if (flying) {
        L.rem = fmaf(static_cast<float>(n), step, L.rem) - g.travel;
        L.t = fmaf(L.inv_vg, g.travel, L.t);
    }
    return ((stopped >> lane) & 1u) == 0u;
}

// The rest of the iteration, once the leg is known to be free of DOMs: fly it, then scatter (or go on / end).
// HIST: the last scene.history_entries scatter points of the photon are kept in the lane's ring in HBM (`ring` points at
// the lane's column of [entry][thread] float4; the rings of all lanes together are a few tens of MB and live in L2):
// (x, y, z, absorption lengths LEFT); the hit record turns the fourth into "absorption lengths used"
// (propagation_kernel.c.cl:833-837).
template <bool TILT, bool ANISO, bool SAVE_ALL, bool MIXED, int V>
__device__ __forceinline__ void finish_leg(Lane &L, const Leg &g, const DevScene &scene, uint32_t rng_a, float4 *ring)
{
    constexpr bool HIST = (V & kVarHist) != 0, TAB = (V & kVarTab) != 0, BLOCK = (V & kVarBlockTransforms) != 0;
    const DevMedium &m = scene.medium;
    Mwc rng{L.rng_x, rng_a};
    // ------------------------------------------------------------------ advance
    L.pxy = __ffma2_rn(L.dxy, make_float2(g.travel, g.travel), L.pxy);
    L.pz = fmaf(L.dz, g.travel, L.pz);
    L.path += g.travel;

    if (g.limited) {
        // the flight goes on (in the neighbouring layer, or past the range limit of the collision
        // map) with what is left of both budgets
        const bool cross = g.d_b <= g.cap;
        if (cross) L.layer += layer_address_step(L.inv_dz);
        L.bud = g.rem;
        if (TILT) L.z_eff = cross ? g.zb : fmaf(L.dz, g.travel, L.z_eff);
        return;
    }
    L.bud.x = g.absorbed ? 0.f : g.rem.x;
    if (ANISO) L.bud.x *= L.inv_aniso;
    if (L.bud.x < kEpsilon) {
        L.status = (SAVE_ALL && !TAB) ? kDying : kDead;   // (#if defined(SAVE_ALL_PHOTONS) && !defined(TABULATE), :800)
        return;
    }
    // ------------------------------------------------------------------ R9 + R8: scatter
    if (HIST) ring[static_cast<size_t>(L.scatters % static_cast<uint32_t>(scene.history_entries)) * (static_cast<size_t>(gridDim.x) * kThreads)] = make_float4(L.pxy.x, L.pxy.y, L.pz, L.bud.x);
    if (ANISO) {
        if constexpr (BLOCK) {
            apply_block_matrix(m.pre, L.dxy, L.dz);
        } else {
            V3 d{L.dxy.x, L.dxy.y, L.dz};
            apply_matrix(m.pre, d);
            L.dxy = make_float2(d.x, d.y); L.dz = d.z;
        }
    }
    const float ru = __uint2float_rz(rng.next());   // the draw, not yet scaled by 2^-32 (the MIXED constants carry the scale)
    float cs;
    const int scat_kind = MIXED ? CLSIMCU_SCAT_MIXED_SL_HG : m.scat_kind;   // the IceCube models' mix is compiled in
    if (MIXED) {
        // both samplers are evaluated and one is selected (no divergent branch); the constants of
        // I3CLSimRandomValueMixed / ...SimplifiedLiu / ...HenyeyGreenstein are folded on the host (DevMedium)
        const float cos_sl = mufu_ex2(fmaf(m.sl_beta, mufu_lg2(ru), m.sl_off)) - 1.f;
        const float r = mufu_rcp(fmaf(m.hg_h1, ru, m.hg_h0));
        const float cos_hg = fmaf(-m.hg_w, r * r, m.hg_c);
        cs = (ru < m.mix_split) ? cos_sl : cos_hg;
    } else {
        const float rr = ru * 2.3283064365386963e-10f;
        if (scat_kind == CLSIMCU_SCAT_MIXED_SL_HG) {
            const float cos_sl = 2.f * fast_pow(rr * m.inv_f_sl, m.sl_beta) - 1.f;
            const float s = 2.f * ((1.f - rr) * m.inv_one_minus_f_sl) - 1.f;
            const float ii = (1.f - m.g2) * mufu_rcp(1.f + m.g * s);
            const float cos_hg = (1.f + m.g2 - ii * ii) * m.inv_2g;
            cs = (rr < m.f_sl) ? cos_sl : cos_hg;
        } else if (scat_kind == CLSIMCU_SCAT_HG) {
            const float s = 2.f * rr - 1.f;
            const float ii = (1.f - m.g2) * mufu_rcp(1.f + m.g * s);
            cs = (1.f + m.g2 - ii * ii) * m.inv_2g;
        } else {
            cs = 2.f * fast_pow(rr, m.sl_beta) - 1.f;
        }
    }
    // (the samplers stay within [-1, 1] up to the rounding of the approximate exp2 / reciprocal: sin^2 saturates at 0)
    rotate_packed(cs, __saturatef(fmaf(-cs, cs, 1.f)), L.dxy, L.dz, rng.next());
    if (ANISO) {
        if constexpr (BLOCK) {
            apply_block_matrix(m.post, L.dxy, L.dz);
        } else {
            V3 d{L.dxy.x, L.dxy.y, L.dz};
            apply_matrix(m.post, d);
            L.dxy = make_float2(d.x, d.y); L.dz = d.z;
        }
    }
    L.inv_dz = raw_inv_dz(L.dz);
    L.bud.y = 0.f;
    ++L.scatters;
    L.rng_x = rng.x;
}

// Slow phase, for a parked lane: the reference's collision test over the pending leg.  No hit:
// the leg is flown here (the same code as in the hot loop) and the lane goes back.  Hit: the photon
// ends at the DOM and is written out.
template <bool TILT, bool ANISO, bool MIXED, int V>
__device__ __noinline__ uint32_t resolve_parked(const DevScene *scene, float *st, uint32_t rng_a, float4 *ring)
{
    constexpr bool NONSTOP = (V & kVarNonStop) != 0;
    const SmemLayout &lay = reinterpret_cast<const SmemHeader *>(smem_base())->lay;
    const V3 pos{st[kPx * kThreads], st[kPy * kThreads], st[kPz * kThreads]};
    const V3 dir{st[kDx * kThreads], st[kDy * kThreads], st[kDz * kThreads]};
    const int who = __float_as_int(st[kPendWho * kThreads]);
    const float leg = st[kPendTravel * kThreads];
    // (the test first, the leg's plan after it: nothing but the hit has to survive the call)
    Collision col{leg, 0, 0, false};
    if constexpr (!NONSTOP) col = collide<false>(scene, who, pos, dir, leg);
    // the leg again (the plan is a function of the parked state alone)
    const SmemPlan sp = table_plan(lay);
    Lane L;
    load_lane<TILT, ANISO>(L, st);
    const uint32_t sbase = smem_address();
    const Leg g = plan_leg<TILT, false, true>(L, *scene, sp.strings, sbase + lay.off_near);
    const int layer_index = (L.layer - static_cast<int>(sbase + kSmLayers)) >> 4;
    if (NONSTOP || col.hit) {
        // the budgets: distInAbsLens is taken for the unshortened flight (propagation_kernel.c.cl:718)
        float at_end = g.absorbed ? 0.f : g.rem.x;
        const bool cross = g.limited && (g.d_b <= g.cap);
        if (g.limited)
            at_end = abs_left_at_end_of_flight(sp.layers, scene->medium.z0, scene->medium.h, scene->medium.num_layers,
                                               cross ? layer_index + layer_step(L.inv_dz) : layer_index, cross ? g.zb : fmaf(L.dz, g.travel, g.zc),
                                               L.dz, L.inv_dz, g.rem.y, g.rem.x, L.f_sp.x, L.f_dust, L.f_sp.y);
        if (ANISO) at_end *= L.inv_aniso;
        if constexpr (NONSTOP) {
            // record every DOM on the leg, nearest first, then fly the leg
            After after{-1.f, -1};
            for (;;) {
                const Collision c = collide<true>(scene, who, pos, dir, leg, after);
                if (!c.hit) break;
                const V3 at{fmaf(dir.x, c.travel, pos.x), fmaf(dir.y, c.travel, pos.y), fmaf(dir.z, c.travel, pos.z)};
                emit_record(scene, st, at, dir, L.path + c.travel, L.scatters, c.string, c.dom, at_end, false, rng_a);
                after = After{c.travel, (c.string << 16) | c.dom};
            }
        } else {
            const V3 end{fmaf(dir.x, col.travel, pos.x), fmaf(dir.y, col.travel, pos.y), fmaf(dir.z, col.travel, pos.z)};
            emit_record(scene, st, end, dir, L.path + col.travel, L.scatters, col.string, col.dom, at_end, false, rng_a);
            return kDead;
        }
    }
    L.status = kActive;
    finish_leg<TILT, ANISO, false, MIXED, V>(L, g, *scene, rng_a, ring);
    store_lane<TILT, ANISO>(L, st);
    return L.status;
}

// Queue fill, by all 32 lanes of the warp at once: the next photons of the warp's step (and of the
// steps after it, fetched from the global work counter) go to queue slots 0, 1, ...; the slot index is
// the creating lane.  Called with an empty queue.  Control state lives in the warp's control block.
__device__ __noinline__ void fill_queue(const DevScene *scene, float *warp_region)
{
    const LaunchArgs &args = reinterpret_cast<const SmemHeader *>(smem_base())->args;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const unsigned lane_bit = 1u << lane;
    float4 *queue = reinterpret_cast<float4 *>(warp_region) + warp * (kQueueChunks * 32);
    uint32_t *wstep = reinterpret_cast<uint32_t *>(warp_region + kOffWarpStep) + warp * kWarpStepWords;
    uint32_t *wctl = reinterpret_cast<uint32_t *>(warp_region + kOffWarpCtl) + warp * kWarpCtlWords;
    const uint32_t gthread = blockIdx.x * kThreads + tid;
    uint32_t w_left = wctl[kWLeft], w_step_index = wctl[kWStepIndex], queued = 0, w_created = wctl[kWCreated];
    bool w_more = wctl[kWMore] != 0;
    __syncwarp();
    Mwc crng{args.rng_x[args.rng_creation_offset + gthread], __ldg(args.rng_a + args.rng_creation_offset + gthread)};
    bool need = true;
    for (;;) {
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
        if (need_mask == 0) break;
        if (w_left == 0) {
            if (!w_more) break;
            // next step for this warp
            uint32_t idx = 0;
            if (lane == 0) idx = atomicAdd(args.work_counter, 1u);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            // a unit of work is one of `step_chunks` equal parts of a step (1 for bunches that fill the device many times
            // over; small bunches are cut finer so that no warp ends up with twice the photons of its neighbours)
            const uint32_t chunks = args.step_chunks, part = idx % chunks;
            idx /= chunks;
            if (idx >= args.num_steps) { w_more = false; break; }
            __syncwarp();
            if (lane < 12) wstep[lane] = __ldg(reinterpret_cast<const uint32_t *>(args.steps) + static_cast<size_t>(idx) * 12 + lane);
            __syncwarp();
            w_step_index = idx;
            w_left = wstep[8];
            if (chunks > 1u) {
                const unsigned long long n = w_left;
                w_left = static_cast<uint32_t>(n * (part + 1u) / chunks) - static_cast<uint32_t>(n * part / chunks);
            }
            // A step whose position, time, direction, length or beta is not a finite number (or is beyond 2^63: the layer table
            // ends at +-1e30) has its photons counted as created and absorbed on the spot.  The reference flies such photons
            // through its outermost layer until they are absorbed -- no hit either way -- but a photon at infinity has no
            // boundary ahead in this kernel's layer walk: it would step off the layer table and never end.  The reference's
            // own step generator makes such a step once in 2^32 draws (gammaDistributedNumber's log(ry / (1 - ry)) at
            // ry == 1, I3CLSimLightSourceToStepConverterUtils.h:100-108): about one step in fifty runs of 1e10 photons.
            const bool out_of_this_world = lane < 8 && (wstep[lane] & 0x7f800000u) >= 0x5f000000u;
            if (__any_sync(0xffffffffu, out_of_this_world)) {
                w_created += w_left;
                w_left = 0;
            }
            if (lane == 0) {
                const V3 axis = step_axis(__uint_as_float(wstep[4]), __uint_as_float(wstep[5]));
                wstep[12] = __float_as_uint(axis.x);
                wstep[13] = __float_as_uint(axis.y);
                wstep[14] = __float_as_uint(axis.z);
            }
            __syncwarp();
            if (w_left == 0) continue; // dummy step (quirk 11)
        }
        // the lanes still in need are the upper ones, so the filled slots stay a prefix
        const int rank = __popc(need_mask & (lane_bit - 1u));
        const bool take = need && (static_cast<uint32_t>(rank) < w_left);
        if (take) {
            crng.x = create_photon(scene, wstep, queue + kQueueChunks * lane, crng.x, crng.a, w_step_index | (static_cast<uint32_t>(lane) << kStepIndexBits));
            need = false;
        }
        const uint32_t took = __popc(__ballot_sync(0xffffffffu, take));
        queued += took;
        w_created += took;
        w_left -= took;
    }
    args.rng_x[args.rng_creation_offset + gthread] = crng.x;
    __syncwarp();
    if (lane == 0) {
        wctl[kWLeft] = w_left; wctl[kWStepIndex] = w_step_index; wctl[kWMore] = w_more ? 1u : 0u; wctl[kWQueued] = queued;
        wctl[kWCreated] = w_created;
    }
    __syncwarp();
}

// A lane takes the photon in queue slot `slot` (four 16-byte chunks): running state into `L`, birth tag (and,
// save-all, the propagation-stream state) into the lane's tag words.
template <bool SAVE_ALL, bool TAB = false>
__device__ __forceinline__ void take_photon(Lane &L, const float4 *slot, float *st, uint32_t layers)
{
    const float4 c0 = slot[0], c1 = slot[1], c2 = slot[2], c3 = slot[3];
    L.pxy = make_float2(c0.x, c0.y); L.pz = c1.x;
    L.dxy = make_float2(c0.z, c0.w); L.dz = c1.y;
    L.inv_dz = raw_inv_dz(c1.y);
    L.bud = make_float2(c1.z, 0.f);
    L.path = 0.f;
    L.f_sp = make_float2(c2.x, c2.y); L.f_dust = c2.z;
    L.scatters = 0u;
    L.layer = static_cast<int>(layers) + 16 * __float_as_int(c1.w);
    L.status = kActive;
    if (TAB) {
        // (the fourth chunk carries the table-mode values in place of the birth tag, see create_photon)
        L.t = c3.x; L.inv_vg = c3.y; L.weight = c3.z; L.rem = c3.w;
        return;
    }
    tag_word(st, 0) = __float_as_uint(c3.x); tag_word(st, 1) = __float_as_uint(c3.y); tag_word(st, 2) = __float_as_uint(c3.z);
    if (SAVE_ALL) {
        float *ptag = st + kOffPopTag;
        ptag[0 * kThreads] = __uint_as_float(static_cast<uint32_t>(L.rng_x));
        ptag[1 * kThreads] = __uint_as_float(static_cast<uint32_t>(L.rng_x >> 32));
    }
}

// ---- slow phase ------------------------------------------------------------------------------
// Runs converged, once per warp each time the hot loop cannot go on by itself: parked lanes get the full
// collision test (hits are written out), save-all photons that ended are recorded, the queue is
// refilled, lanes without a photon take one.  Returns the number of idle lanes that ends the next fast
// phase, or -1 when the warp is done.
template <bool TILT, bool ANISO, bool SAVE_ALL, bool MIXED, int V>
__device__ __noinline__ int slow_phase(const DevScene *scene, float *st, float *warp_region)
{
    constexpr bool HIST = (V & kVarHist) != 0, TAB = (V & kVarTab) != 0;
    const LaunchArgs &args = reinterpret_cast<const SmemHeader *>(smem_base())->args;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const unsigned lane_bit = 1u << lane;
    uint32_t *nseg = &tag_word(st, 3);
    const float4 *queue = reinterpret_cast<const float4 *>(warp_region) + warp * (kQueueChunks * 32);
    uint32_t *wctl = reinterpret_cast<uint32_t *>(warp_region + kOffWarpCtl) + warp * kWarpCtlWords;
    const uint32_t gthread = blockIdx.x * kThreads + tid;
    const uint32_t rng_a = __ldg(args.rng_a + gthread);
    float4 *ring = HIST ? reinterpret_cast<float4 *>(args.history_ring) + gthread : nullptr;
    uint32_t status = __float_as_uint(st[kStatus * kThreads]);

    // ---- photons whose next leg may touch a DOM: the full collision test
    if (!SAVE_ALL && status == kFrozen) status = resolve_parked<TILT, ANISO, MIXED, V>(scene, st, rng_a, ring);
    // ---- save-all: every photon that ended is recorded with probability `prescale`
    //      (propagation_kernel.c.cl:800-826)
    if (SAVE_ALL && status == kDying) {
        Mwc rng{pack64(st[kRngLo * kThreads], st[kRngHi * kThreads]), rng_a};
        const bool keep = rng.co() < scene->prescale;
        st[kRngLo * kThreads] = __uint_as_float(static_cast<uint32_t>(rng.x));
        st[kRngHi * kThreads] = __uint_as_float(static_cast<uint32_t>(rng.x >> 32));
        if (keep) {
            const V3 pos{st[kPx * kThreads], st[kPy * kThreads], st[kPz * kThreads]};
            const V3 dir{st[kDx * kThreads], st[kDy * kThreads], st[kDz * kThreads]};
            emit_record(scene, st, pos, dir, st[kPath * kThreads], __float_as_uint(st[kScatters * kThreads]), 0, 0, 0.f, true, rng_a);
        }
        status = kDead;
    }
    // ---- safety net: a photon whose state is no longer a number would never end (every comparison that ends a photon
    //      is false for a NaN) and would hold its warp, and with it the launch, for ever: dropped here.  (No input is known
    //      to produce one; a path of 1e9 m is 3 s of light.)
    if (status == kActive && !(st[kPath * kThreads] < 1e9f)) status = kDead;
    __syncwarp();

    // ---- lanes without a photon take the next ones of the warp's queue; an empty queue is refilled
    //      first, and once more at the end so that the hot loop starts with photons in stock
    uint32_t queued = wctl[kWQueued];
    for (;;) {
        const unsigned dead = __ballot_sync(0xffffffffu, status == kDead);
        if (queued == 0) {
            if (!(wctl[kWMore] != 0 || wctl[kWLeft] > 0)) break;
            fill_queue(scene, warp_region);
            queued = wctl[kWQueued];
            if (queued == 0) break; // nothing left to create
        }
        if (dead == 0u) break;
        const uint32_t rank = __popc(dead & (lane_bit - 1u));
        if (status == kDead && rank < queued) {
            // statistics: one flight (reference: segment) per scatter, plus the last one
            *nseg += __float_as_uint(st[kScatters * kThreads]) + 1u;
            Lane L;
            L.rng_x = pack64(st[kRngLo * kThreads], st[kRngHi * kThreads]);
            take_photon<SAVE_ALL, TAB>(L, queue + kQueueChunks * (queued - 1u - rank), st, smem_address() + kSmLayers);
            if (TAB) {
                st[(kOffBirthTag + 0 * kThreads)] = L.t; st[(kOffBirthTag + 1 * kThreads)] = L.inv_vg; st[(kOffBirthTag + 2 * kThreads)] = L.weight;
                st[kPendTravel * kThreads] = L.rem;
            }
            st[kPx * kThreads] = L.pxy.x; st[kPy * kThreads] = L.pxy.y; st[kPz * kThreads] = L.pz;
            st[kDx * kThreads] = L.dxy.x; st[kDy * kThreads] = L.dxy.y; st[kDz * kThreads] = L.dz;
            st[kAbsLeft * kThreads] = L.bud.x; st[kScaLeft * kThreads] = 0.f; st[kPath * kThreads] = 0.f;
            st[kFScat * kThreads] = L.f_sp.x; st[kFDust * kThreads] = L.f_dust; st[kFPure * kThreads] = L.f_sp.y;
            st[kScatters * kThreads] = __uint_as_float(0u);
            st[kLayer * kThreads] = __int_as_float(L.layer);
            status = kActive;
        }
        queued -= min(static_cast<uint32_t>(__popc(dead)), queued);
        __syncwarp();
    }
    st[kStatus * kThreads] = __uint_as_float(status);
    __syncwarp();   // every lane has read the control block before lane 0 rewrites it
    if (lane == 0) wctl[kWQueued] = queued;

    const int n_idle = __popc(__ballot_sync(0xffffffffu, status == kDead));
    __syncwarp();
    if (n_idle == 32) return -1;                    // nothing in flight, nothing queued, nothing to fetch
    // idle lanes remain only when the work has run out: drain
    return (n_idle > 0) ? n_idle + 1 : (SAVE_ALL ? kIdleLimitSaveAll : kIdleLimit);
}

template <bool TILT, bool ANISO, bool SAVE_ALL, bool MIXED, int V>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
propagate_persistent(const __grid_constant__ DevScene scene, const __grid_constant__ LaunchArgs args, const __grid_constant__ SmemLayout lay)
{
    constexpr bool HIST = (V & kVarHist) != 0, TAB = (V & kVarTab) != 0;
    uint8_t *smem = smem_base();
    const DevMedium &m = scene.medium;
    const DevGeometry &geo = scene.geo;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    if (tid == 0) {
        SmemHeader *hdr = reinterpret_cast<SmemHeader *>(smem);
        hdr->lay = lay;
        hdr->args = args;
    }
    if (TAB) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(args.tabulate);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&reinterpret_cast<SmemHeader *>(smem)->tab);
        for (int i = tid; i < static_cast<int>(sizeof(TabulateArgs) / 4); i += kThreads) dst[i] = __ldg(src + i);
    }
    const SmemPlan sp = table_plan(lay);

    // ---- stage the hot tables into shared memory (coalesced reads, once per CTA)
    for (int i = tid; i <= m.num_layers; i += kThreads) {
        const float z_low = (i == 0) ? -1e30f : ((i == m.num_layers) ? 1e30f : m.z0 + m.h * static_cast<float>(i));
        sp.layers[i] = (i < m.num_layers) ? make_float4(__ldg(m.b400 + i), __ldg(m.abs_tau + i), __ldg(m.abs_dust + i), z_low)
                                          : make_float4(0.f, 0.f, 0.f, z_low);
    }
    if (TILT) {
        for (int i = tid; i < m.tilt_nd; i += kThreads) {
            const float here = __ldg(m.tilt_dist + i);
            sp.tilt_dist[i] = make_float2(here, (i > 0) ? 1.f / (here - __ldg(m.tilt_dist + i - 1)) : 0.f);
        }
        // the interval grid (see tilt_shift): cell 0 is everything below dist[1], cell c >= 1 starts at dist[1] + (c - 1) w, the
        // last one has no upper end; w is below the smallest gap between nodes (host), so a cell -- widened by a margin that
        // covers the rounding of the cell index -- holds one node at most
        for (int c = tid; c < m.tilt_lut_n; c += kThreads) {
            float2 *lut = sp.tilt_dist + m.tilt_nd + (m.tilt_nd & 1);
            const double w = 1.0 / static_cast<double>(m.tilt_lut_scale), d1 = static_cast<double>(__ldg(m.tilt_dist + 1)), margin = 1e-3 * w;
            const double lo = d1 + (c - 1) * w - margin, hi = d1 + c * w + margin;
            int below = 1;
            float inside = __int_as_float(0x7f800000);
            for (int i = 1; i <= m.tilt_nd - 2; ++i) {
                const float d = __ldg(m.tilt_dist + i);
                if (c > 0 && static_cast<double>(d) < lo) ++below;
                else if (c == m.tilt_lut_n - 1 || static_cast<double>(d) <= hi) inside = fminf(inside, d);
            }
            lut[c] = make_float2(__int_as_float(below), inside);
        }
        for (int i = tid; i < (m.tilt_nd - 1) * (m.tilt_nz - 1); i += kThreads) {
            const int j = i / (m.tilt_nz - 1), k = i - j * (m.tilt_nz - 1);   // column interval [j, j + 1], z interval [k, k + 1]
            const float lo0 = __ldg(m.tilt_corr + j * m.tilt_nz + k), lo1 = __ldg(m.tilt_corr + j * m.tilt_nz + k + 1);
            const float hi0 = __ldg(m.tilt_corr + (j + 1) * m.tilt_nz + k), hi1 = __ldg(m.tilt_corr + (j + 1) * m.tilt_nz + k + 1);
            sp.tilt_corr[i] = make_float4(lo0, hi0, lo1 - lo0, hi1 - hi0);
        }
    }
    if (lay.gen0_n) {
        const DevWlenGenerator &g0 = scene.generators[0];
        float *cum = reinterpret_cast<float *>(smem + lay.off_gen0), *dens = cum + g0.n;
        uint8_t *guide = reinterpret_cast<uint8_t *>(dens + g0.n);
        for (int i = tid; i < g0.n; i += kThreads) {
            cum[i] = __ldg(g0.cumulative + i);
            dens[i] = __ldg(g0.density + i);
        }
#if CLSIMCU_WLEN_BIN_RECORDS
        for (int k = tid; k < g0.n - 1; k += kThreads) {
            float4 *rec = reinterpret_cast<float4 *>(smem + lay.off_gen0_bins);
            float2 *lin = reinterpret_cast<float2 *>(rec + g0.n);
            const float b = __ldg(g0.density + k), b_next = __ldg(g0.density + k + 1);
            float x0, slope;
            if (g0.kind == CLSIMCU_WLEN_INTERP_UNEQUAL) {
                x0 = __ldg(g0.xs + k);
                slope = (b_next - b) / (__ldg(g0.xs + k + 1) - x0);
            } else {
                x0 = static_cast<float>(k) * g0.dx + g0.x0;
                slope = (b_next - b) / g0.dx;
            }
            const float below = (k == 0) ? 0.f : __ldg(g0.cumulative + k);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            if (b == 0.f && slope != 0.f) { c1 = 2.f / slope; c2 = 1.f; }
            else if (b != 0.f && slope == 0.f) { c3 = 1.f / b; }
            else if (b != 0.f) { c0 = 1.f; c1 = (2.f * slope) / (b * b); c2 = b / slope; }
            rec[k] = make_float4(below, x0, c1, c2);
            lin[k] = make_float2(c0, c3);
        }
#endif
        if (tid < kGen0Guide) {
            // the bin of r = tid/64, by the reference's linear scan
            const float r = static_cast<float>(tid) * (1.f / static_cast<float>(kGen0Guide));
            int k = 0;
            while (k < g0.n - 2 && __ldg(g0.cumulative + k + 1) < r) ++k;
            guide[tid] = static_cast<uint8_t>(k);
        }
    }
    if (!SAVE_ALL) {
        for (int i = tid; i < geo.num_strings; i += kThreads) {
            sp.strings[i] = make_float4(__ldg(geo.string_x + i), __ldg(geo.string_y + i), __ldg(geo.string_max_z + i) + geo.om_radius,
                                        __ldg(geo.string_min_z + i) - geo.om_radius);
            sp.string_set[i] = __ldg(geo.string_set + i);
        }
        // the record the pixel map names where it cannot name a string (see cell_walk): a NaN axis fails every "farther than" test
        if (tid == 0) sp.strings[geo.num_strings] = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), 0.f, 0.f);
        for (int i = tid; i < geo.num_sets; i += kThreads)
            sp.sets[i] = make_float4(__ldg(geo.set_start_z + i), 1.f / __ldg(geo.set_layer_height + i),
                                     static_cast<float>(__ldg(geo.set_layer_count + i)), static_cast<float>(i * geo.max_layers));
        for (int i = tid; i < geo.layer_table_size; i += kThreads) sp.layer_to_dom[i] = __ldg(geo.layer_to_dom + i);
        for (int gI = 0; gI < geo.num_grids; ++gI) {
            const int n = geo.grids[gI].num_x * geo.grids[gI].num_y;
            for (int i = tid; i < n; i += kThreads) sp.cells[lay.cell_offset[gI] + i] = __ldg(geo.grids[gI].cell_to_string + i);
        }
        for (int i = tid; i < geo.near_nx * geo.near_ny; i += kThreads) sp.near[i] = __ldg(geo.near_info + i);
    }

    // ---- lane and warp state
    float *st = reinterpret_cast<float *>(smem + kSmState) + tid;
    float *warp_region = reinterpret_cast<float *>(smem + kSmQueue);
    const uint32_t sbase = smem_address();
    const uint32_t layers = sbase + kSmLayers, near = sbase + kSmNear;   // (the pixel map is not read in save-all mode)
    const float4 *queue = reinterpret_cast<const float4 *>(warp_region) + warp * (kQueueChunks * 32);
    uint32_t *wctl = reinterpret_cast<uint32_t *>(warp_region + kOffWarpCtl) + warp * kWarpCtlWords;
    const uint32_t gthread = blockIdx.x * kThreads + tid;
    const uint32_t rng_a = args.rng_a[gthread];
    float4 *ring = HIST ? reinterpret_cast<float4 *>(args.history_ring) + gthread : nullptr;
    {
        const uint64_t x = args.rng_x[gthread];
        st[kRngLo * kThreads] = __uint_as_float(static_cast<uint32_t>(x));
        st[kRngHi * kThreads] = __uint_as_float(static_cast<uint32_t>(x >> 32));
        st[kStatus * kThreads] = __uint_as_float(static_cast<uint32_t>(kDead));
        st[kScatters * kThreads] = __uint_as_float(0xffffffffu); // no photon yet: counts as 0 flights when replaced
        tag_word(st, 3) = 0u;
        if (lane == 0) {
            wctl[kWLeft] = 0u; wctl[kWStepIndex] = 0xffffffffu; wctl[kWMore] = 1u; wctl[kWQueued] = 0u; wctl[kWCreated] = 0u;
        }
    }
    __syncthreads();

    for (;;) {
        const int limit = slow_phase<TILT, ANISO, SAVE_ALL, MIXED, V>(args.scene_dev, st, warp_region);
        if (limit < 0) break;
        // ---- fast phase: photon state in registers, no calls.  A lane whose photon ended takes the next
        //      one from the warp's queue right here; the phase ends when the queue runs dry, or when
        //      `limit` lanes wait for the slow phase.
        Lane L;
        load_lane<TILT, ANISO, TAB>(L, st);
        if (!ANISO) {
            // the unit length of the direction, restored once per fast phase (see rotate_by)
            const float inv = fmaf(L.dxy.x * L.dxy.x + L.dxy.y * L.dxy.y + L.dz * L.dz, -0.5f, 1.5f);
            L.dxy.x *= inv; L.dxy.y *= inv; L.dz *= inv;
            L.inv_dz = raw_inv_dz(L.dz);
        }
        uint32_t queued = wctl[kWQueued];
        const bool more = (wctl[kWMore] != 0u) || (wctl[kWLeft] > 0u);
        for (;;) {
            const unsigned idle = __ballot_sync(0xffffffffu, L.status != kActive);
            if (idle != 0u) {
                const unsigned dead = __ballot_sync(0xffffffffu, L.status == kDead);
                uint32_t n_dead = __popc(dead);
                const int n_waiting = __popc(idle & ~dead);    // parked (or, save-all, ended) lanes: the slow phase's business
                // Lanes without a photon are served in batches: taking a photon costs the whole warp some
                // fifty instructions however many lanes take one, an idle lane costs 1/32 of an iteration.
                if (n_dead >= static_cast<uint32_t>(kRefillBatch) || n_waiting > 0 || idle == 0xffffffffu) {
                    if (n_dead > 0u && queued > 0u) {
                        const uint32_t rank = __popc(dead & lanemask_lt());
                        if (L.status == kDead && rank < queued) {
                            tag_word(st, 3) += L.scatters + 1u;   // statistics, kept in shared memory: a register here is a spill
                            // (the layer table's address is derived again here rather than kept alive across the loop: a spill otherwise)
                            take_photon<SAVE_ALL, TAB>(L, queue + kQueueChunks * (queued - 1u - rank), st, smem_address() + kSmLayers);
                        }
                        const uint32_t taken = min(n_dead, queued);
                        queued -= taken;
                        n_dead -= taken;
                    }
                    // lanes left without a photon: the slow phase refills the queue (unless the work has run out)
                    if (n_dead > 0u && more) break;
                    if (n_waiting + static_cast<int>(n_dead) >= limit) break;
                }
            }
            if constexpr (TAB) {
                // table mode: no DOMs, no collision map; plan, sample the leg (all lanes together, see sample_leg), fly it
#pragma unroll 1
                for (int leg = 0; leg < kHotUnroll; ++leg) {
                    const bool flying = L.status == kActive;
                    Leg g;
                    g.travel = 0.f; g.q = make_float2(0.f, 0.f);
                    if (flying) {
                        start_flight<TILT, ANISO>(L, m, layers, sp.tilt_dist, sp.tilt_corr, rng_a);
                        g = plan_leg<TILT, true, false>(L, scene, sp.strings, near);
                    }
                    __syncwarp();
                    const bool goes_on = sample_leg<ANISO>(L, g, flying, scene);
                    if (flying) {
                        if (goes_on) finish_leg<TILT, ANISO, SAVE_ALL, MIXED, V>(L, g, scene, rng_a, ring);
                        else L.status = kDead;   // out of the table's range: the photon is stopped; table mode records no photons
                    }
                    __syncwarp();
                }
                continue;
            }
            float clearance = 0.f;   // set by the first leg, used by the others (see plan_leg)
            if (L.status == kActive)
                advance_photon<TILT, ANISO, SAVE_ALL, MIXED, true, V>(L, clearance, scene, args.scene_dev, layers, sp.strings, near, sp.tilt_dist,
                                                                         sp.tilt_corr, rng_a, st, ring);
#pragma unroll
            for (int leg = 1; leg < ((TILT || ANISO) ? kHotUnrollTilted : kHotUnroll); ++leg)
                if (L.status == kActive)
                    advance_photon<TILT, ANISO, SAVE_ALL, MIXED, kLookEveryLeg, V>(L, clearance, scene, args.scene_dev, layers, sp.strings, near,
                                                                                      sp.tilt_dist, sp.tilt_corr, rng_a, st, ring);
        }
        store_lane<TILT, ANISO, TAB>(L, st);
        __syncwarp();   // every lane has read the control block before lane 0 rewrites it
        if (lane == 0) wctl[kWQueued] = queued;
        __syncwarp();
    }

    args.rng_x[gthread] = pack64(st[kRngLo * kThreads], st[kRngHi * kThreads]);
    if (args.count_stats) {
        // warp-level reduction, one atomic per warp; the lane's last photon has not been counted yet
        unsigned long long segs = static_cast<unsigned long long>(tag_word(st, 3)) +
                                  (__float_as_uint(st[kScatters * kThreads]) + 1u);
        for (int o = 16; o > 0; o >>= 1) segs += __shfl_down_sync(0xffffffffu, segs, o);
        if (lane == 0) {
            atomicAdd(args.stats + 0, static_cast<unsigned long long>(wctl[kWCreated]));
            atomicAdd(args.stats + 1, segs);
        }
    }
}

template <bool TILT, bool ANISO, bool SAVE_ALL, bool MIXED, int V = 0>
int launch_mix(const DevScene &scene, const LaunchArgs &args, int blocks, cudaStream_t stream)
{
    const SmemLayout lay = plan_smem(scene);
    auto kernel = propagate_persistent<TILT, ANISO, SAVE_ALL, MIXED, V>;
    // the attribute belongs to the (function, device) pair and engines on several devices launch from several threads
    // of one process: set it on every launch (a host-side table write, no device work)
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)) != cudaSuccess) return -3;
    kernel<<<blocks, kThreads, lay.total, stream>>>(scene, args, lay);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;   // (peek: the caller reads the error text)
}

template <bool TILT, bool ANISO, bool SAVE_ALL>
int launch_variant(const DevScene &scene, const LaunchArgs &args, int blocks, cudaStream_t stream)
{
    if constexpr (SAVE_ALL) {
        // table-maker variant: save-all scene (no DOMs), fixed number of absorption lengths, four table axes
        if (args.tabulate) return launch_mix<TILT, ANISO, true, false, kVarTab>(scene, args, blocks, stream);
        // the save-all variants are checkers' tools: no need to specialise them further
        if (scene.history_entries > 0) return launch_mix<TILT, ANISO, true, false, kVarHist>(scene, args, blocks, stream);
        return launch_mix<TILT, ANISO, true, false, 0>(scene, args, blocks, stream);
    } else {
        // photon history and StopDetectedPhotons = false: checkers' and event-display options, one generic instantiation
        if (scene.history_entries > 0 || !scene.stop_detected) {
            if (scene.history_entries > 0 && !scene.stop_detected) return launch_mix<TILT, ANISO, false, false, kVarHist | kVarNonStop>(scene, args, blocks, stream);
            if (scene.history_entries > 0) return launch_mix<TILT, ANISO, false, false, kVarHist>(scene, args, blocks, stream);
            return launch_mix<TILT, ANISO, false, false, kVarNonStop>(scene, args, blocks, stream);
        }
        if (scene.medium.scat_kind == CLSIMCU_SCAT_MIXED_SL_HG && scene.medium.mix_folded) {
            if constexpr (ANISO) {
                const float *a = scene.medium.pre, *b = scene.medium.post;
                // (scene.generic_transforms: test hook, takes the nine-product form for matrices that would qualify)
                if (a[2] == 0.f && a[5] == 0.f && a[6] == 0.f && a[7] == 0.f && b[2] == 0.f && b[5] == 0.f && b[6] == 0.f && b[7] == 0.f &&
                    !scene.generic_transforms)
                    return launch_mix<TILT, ANISO, false, true, kVarBlockTransforms>(scene, args, blocks, stream);
            }
            return launch_mix<TILT, ANISO, false, true, 0>(scene, args, blocks, stream);
        }
        return launch_mix<TILT, ANISO, false, false, 0>(scene, args, blocks, stream);
    }
}

} // namespace

bool fast_kernel_supports(const DevScene &scene, const char **why)
{
    static const char *k_renorm = "non-renormalising direction transforms are only implemented by the reference-order kernel";
    static const char *k_smem = "geometry/medium tables do not fit into shared memory";
    static const char *k_strings = "more than 4094 strings";
    static const char *k_layers = "more than 255 ice layers";
    if (scene.medium.anisotropy && (!scene.medium.pre_renorm || !scene.medium.post_renorm)) { *why = k_renorm; return false; }
    if (scene.geo.num_strings > 4094) { *why = k_strings; return false; }
    if (scene.medium.num_layers + 1 > kMaxStagedLayers) { *why = k_layers; return false; }
    if (plan_smem(scene).total + 1024u > kSmemBudget / kBlocksPerSM) { *why = k_smem; return false; }
    return true;
}

bool fast_kernel_smem_is_the_problem(const DevScene &scene)
{
    DevScene probe = scene;
    probe.geo.near_nx = probe.geo.near_ny = 0;
    const char *why = nullptr;
    return fast_kernel_supports(probe, &why) && !fast_kernel_supports(scene, &why);
}

void fast_kernel_geometry(int device, int *grid_blocks, int *threads_per_block)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    *grid_blocks = sms * kBlocksPerSM;
    *threads_per_block = kThreads;
}

int launch_fast_kernel(const DevScene &scene, const LaunchArgs &args, int grid_blocks, void *stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (args.num_steps == 0) return 0;
    if (args.num_steps > (1u << kStepIndexBits)) return -1;
    LaunchArgs chunked = args;
    {
        // fewer than 8 steps per resident warp: cut the steps into parts (see fill_queue)
        const uint32_t warps = static_cast<uint32_t>(grid_blocks) * kWarpsPerBlock;
        uint32_t chunks = 1;
        while (chunks < 8u && static_cast<unsigned long long>(args.num_steps) * chunks < 8ull * warps) chunks *= 2u;
        chunked.step_chunks = chunks;
    }
    const bool tilt = scene.medium.tilt_nd > 0, aniso = scene.medium.anisotropy != 0;
    if (scene.save_all) {
        if (tilt && aniso) return launch_variant<true, true, true>(scene, chunked, grid_blocks, stream);
        if (tilt) return launch_variant<true, false, true>(scene, chunked, grid_blocks, stream);
        if (aniso) return launch_variant<false, true, true>(scene, chunked, grid_blocks, stream);
        return launch_variant<false, false, true>(scene, chunked, grid_blocks, stream);
    }
    if (tilt && aniso) return launch_variant<true, true, false>(scene, chunked, grid_blocks, stream);
    if (tilt) return launch_variant<true, false, false>(scene, chunked, grid_blocks, stream);
    if (aniso) return launch_variant<false, true, false>(scene, chunked, grid_blocks, stream);
    return launch_variant<false, false, false>(scene, chunked, grid_blocks, stream);
}

} // namespace clsimcu
