// kernel_fast.cu -- persistent photon-propagation kernel for B200 (sm_100a).  The product path.
//
// Design (B200-first, not the reference's one-work-item-per-step loop):
//
//  * Persistent CTAs (a multiple of the SM count), 8 warps each.  A WARP owns a step at a time
//    and pulls the next one from a global work counter, so step bunches of any size balance
//    across the 148 SMs and there is no per-step tail.
//  * One photon per LANE, one MWC stream per lane (multipliers from the safe-prime table).
//  * Per-lane MAILBOX in shared memory: photon creation (wavelength table search, Cherenkov cone,
//    the wavelength-only transcendental factors of the ice model, lifetime) is done for many
//    lanes at once, ahead of time, whenever >= kRefillThreshold lanes have used up their spare
//    photon.  A lane whose photon dies pops its spare with a dozen shared-memory loads and
//    keeps going; the expensive creation code never runs for one or two lanes only.  The
//    start-of-flight record needed for a hit stays in shared memory, not in registers.
//  * The wavelength dependence of the ice (powr/exp of R4) is hoisted to once per photon: per
//    segment and per layer crossed only two FMAs on per-layer coefficients remain, which live
//    in shared memory as one float4 per layer together with the cell grid and string tables.
//  * One reciprocal per segment in the common same-layer case; SL and HG scattering angles are
//    both evaluated and selected (no divergent branch); fast approximate MUFU intrinsics.
//  * Hits: warp-aggregated atomic reservation, five 16-byte stores per record, string/DOM IDs
//    and the wavelength-bias weight applied on the device.
//
// Tensor cores are not used: there is no contraction anywhere on this path.
//
// Physics restated from resources/kernels/propagation_kernel.c.cl and
// sparse_collision_kernel.c.cl (citations at each block); the arithmetic is re-formulated, so
// agreement with the reference is statistical (per-DOM counts, time and angle distributions)
// and per photon within fp32 tolerance when a photon is replayed from its recorded RNG state.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/clsimcuda.h"
#include "device_scene.h"

namespace clsimcu {
namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;
constexpr int kBlocksPerSM = 4;
constexpr int kRefillThreshold = 14; // lanes with an empty mailbox that trigger a creation pass
constexpr float kSpeedOfLight = 0.299792458f;
constexpr float kPi = 3.14159265359f;
constexpr float kEpsilon = 0.00001f;
constexpr float kLn2 = 0.69314718056f;
constexpr int kStartWords = 10; // x y z t dx dy dz wlen abs_initial step_index
constexpr int kSpareWords = 3;  // scattering factor, dust factor, pure-ice absorption

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
    // mwcrng_kernel.cl:12-28; the conversion rounds toward zero so 1.0 is never returned
    __device__ __forceinline__ float co()
    {
        x = static_cast<uint64_t>(static_cast<uint32_t>(x)) * a + (x >> 32);
        return __uint2float_rz(static_cast<uint32_t>(x)) * 2.3283064365386963e-10f;
    }
    // 1 - u/2^32 in one FMA (the scaling is exact, so this equals the two-step form)
    __device__ __forceinline__ float oc()
    {
        x = static_cast<uint64_t>(static_cast<uint32_t>(x)) * a + (x >> 32);
        return __fmaf_rn(__uint2float_rz(static_cast<uint32_t>(x)), -2.3283064365386963e-10f, 1.0f);
    }
};

// Shared-memory plan, carved out of the dynamic allocation.
struct SmemPlan {
    float4 *layers;          // [num_layers] (b400, D*aDust+E, 1+0.01*dTau, 0)
    float4 *strings;         // [num_strings] (x, y, zmax+R, zmin-R)
    float4 *sets;            // [num_sets] (start_z, 1/height, num_layers, row offset)
    uint8_t *string_set;     // [num_strings]
    uint16_t *layer_to_dom;  // [layer_table_size]
    uint16_t *cells;         // concatenated grids
    uint8_t *near_d1;        // distance field, one byte per pixel
    float *start;            // [2][kStartWords][kThreads]
    float *spare;            // [kSpareWords][kThreads]
    uint32_t *warp_step;     // [kWarpsPerBlock][12]
};

struct SmemLayout {
    uint32_t off_layers, off_strings, off_sets, off_string_set, off_layer_to_dom, off_cells, off_near, off_start, off_spare, off_warp_step;
    uint32_t cell_offset[kMaxSubdetectors];
    uint32_t total;
};

__host__ __device__ inline uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

__host__ SmemLayout plan_smem(const DevScene &s)
{
    SmemLayout L{};
    uint32_t at = 0;
    L.off_layers = at; at = align16(at + s.medium.num_layers * 16);
    L.off_strings = at; at = align16(at + s.geo.num_strings * 16);
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
    L.off_near = at; at = align16(at + s.geo.near_nx * s.geo.near_ny + 4);
    L.off_start = at; at = align16(at + 2 * kStartWords * kThreads * 4);
    L.off_spare = at; at = align16(at + kSpareWords * kThreads * 4);
    L.off_warp_step = at; at = align16(at + kWarpsPerBlock * 12 * 4);
    L.total = at;
    return L;
}

struct V3 {
    float x, y, z;
};

// R8 (propagation_kernel.c.cl:83-129) with approximate MUFU ops
__device__ __forceinline__ void rotate_by(float cosa, float sina, V3 &d, float rnd)
{
    float sinb, cosb;
    __sincosf(2.0f * kPi * rnd, &sinb, &cosb);
    const float s2 = fmaxf(0.f, 1.f - d.z * d.z);
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
    const float inv = mufu_rsqrt(nx * nx + ny * ny + nz * nz);
    d.x = nx * inv; d.y = ny * inv; d.z = nz * inv;
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

// R3a: same bin as the reference's linear scan, found by bisection (cumulative is non-decreasing)
__device__ float draw_wavelength(const DevWlenGenerator &g, Mwc &rng)
{
    if (g.kind == CLSIMCU_WLEN_CONSTANT) return g.value;
    const float r = rng.oc();
    if (g.kind == CLSIMCU_WLEN_NO_DISPERSION) return mufu_rcp(g.min_val + r * g.range);
    // smallest k in [0, n-2] with cumulative[k+1] >= r
    int lo = 0, hi = g.n - 2;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(g.cumulative + mid + 1) >= r) hi = mid;
        else lo = mid + 1;
    }
    const int k = lo;
    const float below = (k == 0) ? 0.f : __ldg(g.cumulative + k);
    const float b = __ldg(g.density + k);
    float x0, slope;
    if (g.kind == CLSIMCU_WLEN_INTERP_UNEQUAL) {
        x0 = __ldg(g.xs + k);
        slope = (__ldg(g.density + k + 1) - b) / (__ldg(g.xs + k + 1) - x0);
    } else {
        x0 = static_cast<float>(k) * g.dx + g.x0;
        slope = (__ldg(g.density + k + 1) - b) / g.dx;
    }
    const float dy = r - below;
    if ((b == 0.f) && (slope == 0.f)) return x0;
    if (b == 0.f) return x0 + sqrtf(2.f * dy / slope);
    if (slope == 0.f) return x0 + dy / b;
    return x0 + (sqrtf(dy * (2.f * slope) / (b * b) + 1.f) - 1.f) * b / slope;
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

// R4a (I3CLSimScalarFieldIceTiltZShift.cxx:145-216)
__device__ float tilt_shift(const DevMedium &m, float x, float y, float z)
{
    const float zr = (z - m.tilt_z0) * mufu_rcp(m.tilt_dz);
    const int k = min(max(__float2int_rd(zr), 0), m.tilt_nz - 2);
    const float above = zr - static_cast<float>(k);
    const float below = 1.f - above;
    const float nr = m.tilt_lnx * x + m.tilt_lny * y;
    int j = 1;
    while (j < m.tilt_nd - 1 && !(nr < __ldg(m.tilt_dist + j))) ++j;
    const float here = __ldg(m.tilt_dist + j), prev = __ldg(m.tilt_dist + j - 1);
    const float w_lo = (here - nr) * mufu_rcp(here - prev);
    const float w_hi = 1.f - w_lo;
    const float *lo_row = m.tilt_corr + (j - 1) * m.tilt_nz + k;
    const float *hi_row = m.tilt_corr + j * m.tilt_nz + k;
    const float v_lo = __ldg(lo_row + 1) * above + __ldg(lo_row) * below;
    const float v_hi = __ldg(hi_row + 1) * above + __ldg(hi_row) * below;
    return v_hi * w_hi + v_lo * w_lo;
}

__device__ __forceinline__ void apply_matrix(const float *M, V3 &d)
{
    const float nx = M[0] * d.x + M[1] * d.y + M[2] * d.z;
    const float ny = M[3] * d.x + M[4] * d.y + M[5] * d.z;
    const float nz = M[6] * d.x + M[7] * d.y + M[8] * d.z;
    const float inv = mufu_rsqrt(nx * nx + ny * ny + nz * nz);
    d.x = nx * inv; d.y = ny * inv; d.z = nz * inv;
}

__device__ __forceinline__ float life_of_current(const float *start0, int buffer)
{
    return start0[buffer * (kStartWords * kThreads) + 8 * kThreads];
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

__device__ __forceinline__ void test_string(const SmemPlan &sp, const DevGeometry &geo, int s, const V3 &pos, const V3 &dir,
                                            float inv_xy2, float inv_pancake, Collision &c)
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
        if (entry < c.travel) {
            c.travel = entry;
            c.hit = true;
            c.string = s;
            c.dom = dom;
        }
    }
}

// Cell level, the reference's walk over the xy cells covered by the segment
// (sparse_collision_kernel.c.cl:194-303, 305-460).  Only taken when the distance field cannot
// narrow the candidates down to one string, so it is kept out of line.
__device__ __noinline__ void cell_walk(const SmemPlan sp, const DevGeometry &geo, const SmemLayout &lay, V3 pos, V3 dir, float inv_xy2,
                                       float inv_pancake, Collision &c)
{
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
                if (s != 0xFFFF) test_string(sp, geo, s, pos, dir, inv_xy2, inv_pancake, c);
            }
        }
    }
}

// ---- R10: hit output ----------------------------------------------------------------------------
// Called by the lanes whose photon was just detected (or absorbed, in save-all mode): `pos` is
// the end point, `path` the full path length.  Reservation is aggregated over the lanes that
// arrive together; the record leaves as five 16-byte stores.
__device__ __noinline__ void emit_record(const DevScene &scene, const LaunchArgs &args, const float *st, V3 pos, V3 dir, float path,
                                         uint32_t scatters, int hit_string, int hit_dom, float dist_abs, bool save_all,
                                         uint64_t tag_create, uint64_t tag_pop, uint64_t tag_resume, uint32_t interrupt_at, uint32_t rng_a)
{
    const unsigned peers = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(args.hit_counter, static_cast<uint32_t>(__popc(peers)));
    base = __shfl_sync(peers, base, leader);
    const uint32_t slot = base + __popc(peers & ((1u << lane) - 1u));
    if (slot >= args.max_hits) return; // counted but dropped (quirk 10)
    const DevGeometry &geo = scene.geo;
    const float stt = st[3 * kThreads], wlen = st[7 * kThreads];
    const uint32_t step_index = __float_as_uint(st[9 * kThreads]);
    const clsimcu_step *step = static_cast<const clsimcu_step *>(args.steps) + step_index;
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
    const float ivg = inv_group_velocity(scene.medium, wlen);
    float th, ph, sth, sph;
    to_spherical(dir.x, dir.y, dir.z, th, ph);
    to_spherical(st[4 * kThreads], st[5 * kThreads], st[6 * kThreads], sth, sph);
    float4 *dst = reinterpret_cast<float4 *>(static_cast<clsimcu_photon *>(args.photons) + slot);
    dst[0] = make_float4(pos.x - qx, pos.y - qy, pos.z - qz, stt + path * ivg);
    dst[1] = make_float4(th, ph, wlen, path);
    dst[2] = make_float4(__uint_as_float(scatters), __ldg(&step->weight) / bias_at(scene.bias, wlen),
                         __uint_as_float(__ldg(&step->identifier)), __uint_as_float(ids));
    dst[3] = make_float4(st[0 * kThreads], st[1 * kThreads], st[2 * kThreads], stt);
    dst[4] = make_float4(sth, sph, 1.f / ivg, dist_abs);
    if (save_all && args.rng_tag_x) {
        args.rng_tag_x[3 * static_cast<size_t>(slot)] = tag_create;
        args.rng_tag_x[3 * static_cast<size_t>(slot) + 1] = tag_pop;
        args.rng_tag_x[3 * static_cast<size_t>(slot) + 2] = tag_resume;
        args.rng_tag_a[2 * static_cast<size_t>(slot)] = rng_a;
        args.rng_tag_a[2 * static_cast<size_t>(slot) + 1] = interrupt_at;
    }
}

template <bool TILT, bool ANISO, bool SAVE_ALL>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
propagate_persistent(const __grid_constant__ DevScene scene, const __grid_constant__ LaunchArgs args, const __grid_constant__ SmemLayout lay)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const DevMedium &m = scene.medium;
    const DevGeometry &geo = scene.geo;
    SmemPlan sp;
    sp.layers = reinterpret_cast<float4 *>(smem + lay.off_layers);
    sp.strings = reinterpret_cast<float4 *>(smem + lay.off_strings);
    sp.sets = reinterpret_cast<float4 *>(smem + lay.off_sets);
    sp.string_set = smem + lay.off_string_set;
    sp.layer_to_dom = reinterpret_cast<uint16_t *>(smem + lay.off_layer_to_dom);
    sp.cells = reinterpret_cast<uint16_t *>(smem + lay.off_cells);
    sp.near_d1 = smem + lay.off_near;
    sp.start = reinterpret_cast<float *>(smem + lay.off_start);
    sp.spare = reinterpret_cast<float *>(smem + lay.off_spare);
    sp.warp_step = reinterpret_cast<uint32_t *>(smem + lay.off_warp_step);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const unsigned lane_bit = 1u << lane;

    // ---- stage the hot tables into shared memory (coalesced reads, once per CTA)
    for (int i = tid; i < m.num_layers; i += kThreads)
        sp.layers[i] = make_float4(__ldg(m.b400 + i), __ldg(m.abs_dust + i), __ldg(m.abs_tau + i), 0.f);
    if (!SAVE_ALL) {
        for (int i = tid; i < geo.num_strings; i += kThreads) {
            sp.strings[i] = make_float4(__ldg(geo.string_x + i), __ldg(geo.string_y + i), __ldg(geo.string_max_z + i) + geo.om_radius,
                                        __ldg(geo.string_min_z + i) - geo.om_radius);
            sp.string_set[i] = __ldg(geo.string_set + i);
        }
        for (int i = tid; i < geo.num_sets; i += kThreads)
            sp.sets[i] = make_float4(__ldg(geo.set_start_z + i), 1.f / __ldg(geo.set_layer_height + i),
                                     static_cast<float>(__ldg(geo.set_layer_count + i)), static_cast<float>(i * geo.max_layers));
        for (int i = tid; i < geo.layer_table_size; i += kThreads) sp.layer_to_dom[i] = __ldg(geo.layer_to_dom + i);
        for (int gI = 0; gI < geo.num_grids; ++gI) {
            const int n = geo.grids[gI].num_x * geo.grids[gI].num_y;
            for (int i = tid; i < n; i += kThreads) sp.cells[lay.cell_offset[gI] + i] = __ldg(geo.grids[gI].cell_to_string + i);
        }
        const int near_words = (geo.near_nx * geo.near_ny + 3) / 4;
        for (int i = tid; i < near_words; i += kThreads)
            reinterpret_cast<uint32_t *>(sp.near_d1)[i] = __ldg(reinterpret_cast<const uint32_t *>(geo.near_d1) + i);
    }
    __syncthreads();

    float *start0 = sp.start + tid;                          // + word*kThreads (+ kStartWords*kThreads for buffer 1)
    float *spare = sp.spare + tid;
    uint32_t *wstep = sp.warp_step + warp * 12;

    const uint32_t gthread = blockIdx.x * kThreads + tid;
    Mwc rng{args.rng_x[gthread], args.rng_a[gthread]};

    // warp-uniform state
    uint32_t w_step_index = 0xffffffffu;
    uint32_t w_left = 0;         // photons of the warp's step not yet handed to a lane
    bool w_more = true;          // the global queue may still have steps
    V3 w_axis{0.f, 0.f, 1.f};
    unsigned spare_mask = 0;     // lanes whose mailbox holds a photon

    // lane state
    bool alive = false;
    int cur = 0;                 // which start buffer holds the photon in flight
    V3 pos{0.f, 0.f, 0.f}, dir{0.f, 0.f, 1.f};
    float abs_left = 0.f, path = 0.f;
    float f_scat = 0.f, f_dust = 0.f, f_pure = 0.f;
    uint32_t scatters = 0;
    int layer = 0;
    unsigned long long n_created = 0, n_segments = 0;
    // RNG bookkeeping for single-photon replay by a checker (save-all variants only; dead code
    // otherwise).  A lane's stream serves, in this order: creation of a photon (x_create), later
    // its propagation from x_pop, interrupted at most once -- after `interrupt_at` scatters -- by
    // the creation of the lane's next spare, after which it resumes from x_resume.
    uint64_t spare_tag_create = 0, cur_tag_create = 0, cur_tag_pop = 0, cur_tag_resume = 0;
    uint32_t cur_interrupt_at = 0xffffffffu;

    const float inv_h = m.inv_h;
    const float inv_fsl = (m.f_sl > 0.f) ? 1.f / m.f_sl : 0.f;
    const float inv_omf = (m.one_minus_f_sl > 0.f) ? 1.f / m.one_minus_f_sl : 0.f;
    const float inv_2g = 1.f / (2.f * m.g);

    for (;;) {
        // One vote per iteration; everything else on the control path runs only when a lane is dead.
        unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
        if (alive_mask != 0xffffffffu) {
            // ---------------------------------------------------------- refill (converged)
            const bool work = w_more || w_left > 0;
            if (work && (__popc(~spare_mask) >= kRefillThreshold || (alive_mask | spare_mask) == 0u)) {
                bool need = !(spare_mask & lane_bit);
                for (;;) {
                    const unsigned need_mask = __ballot_sync(0xffffffffu, need);
                    if (need_mask == 0) break;
                    if (w_left == 0) {
                        if (!w_more) break;
                        // next step for this warp
                        uint32_t idx = 0;
                        if (lane == 0) idx = atomicAdd(args.work_counter, 1u);
                        idx = __shfl_sync(0xffffffffu, idx, 0);
                        if (idx >= args.num_steps) { w_more = false; break; }
                        __syncwarp();
                        if (lane < 12) wstep[lane] = __ldg(reinterpret_cast<const uint32_t *>(args.steps) + static_cast<size_t>(idx) * 12 + lane);
                        __syncwarp();
                        w_step_index = idx;
                        w_left = wstep[8];
                        float st, ct, sph, cph;
                        __sincosf(__uint_as_float(wstep[4]), &st, &ct);
                        __sincosf(__uint_as_float(wstep[5]), &sph, &cph);
                        w_axis.x = st * cph; w_axis.y = st * sph; w_axis.z = ct;
                        if (w_left == 0) continue; // dummy step (quirk 11)
                    }
                    const int rank = __popc(need_mask & (lane_bit - 1u));
                    const bool take = need && (static_cast<uint32_t>(rank) < w_left);
                    if (take) {
                        // R3: createPhotonFromTrack (propagation_kernel.c.cl:132-184)
                        const uint64_t x_before = rng.x;
                        const float s_x = __uint_as_float(wstep[0]), s_y = __uint_as_float(wstep[1]), s_z = __uint_as_float(wstep[2]);
                        const float s_t = __uint_as_float(wstep[3]), s_len = __uint_as_float(wstep[6]), s_beta = __uint_as_float(wstep[7]);
                        const uint32_t source = (wstep[11] & 0xffu);
                        const float shift = s_len * rng.co();
                        V3 d = w_axis;
                        float wlen;
                        if (scene.num_generators <= 1 || source == 0) {
                            wlen = draw_wavelength(scene.generators[0], rng);
                            const float cos_c = fminf(1.f, mufu_rcp(s_beta * phase_index(m, wlen)));
                            const float sin_c = mufu_sqrt(1.f - cos_c * cos_c);
                            rotate_by(cos_c, sin_c, d, rng.co());
                        } else {
                            wlen = (source < static_cast<uint32_t>(scene.num_generators)) ? draw_wavelength(scene.generators[source], rng) : 0.f;
                        }
                        const float life = scene.fixed_abs ? scene.fixed_abs_lens : -fast_ln(rng.oc());
                        // the spare goes to the buffer that is not in flight
                        float *st = start0 + (cur ^ (alive ? 1 : 0)) * (kStartWords * kThreads);
                        st[0 * kThreads] = s_x + w_axis.x * shift;
                        st[1 * kThreads] = s_y + w_axis.y * shift;
                        st[2 * kThreads] = s_z + w_axis.z * shift;
                        st[3 * kThreads] = s_t + shift * mufu_rcp(kSpeedOfLight * s_beta);
                        st[4 * kThreads] = d.x;
                        st[5 * kThreads] = d.y;
                        st[6 * kThreads] = d.z;
                        st[7 * kThreads] = wlen;
                        st[8 * kThreads] = life;
                        st[9 * kThreads] = __uint_as_float(w_step_index);
                        // wavelength-only factors of R4 (…_Optimizers.cxx:123-250), once per photon
                        const float nm = wlen * 1e9f;
                        spare[0 * kThreads] = fast_pow(wlen * m.inv_ref_wlen, -m.alpha);                // 1/scatLen = b400 * this
                        spare[1 * kThreads] = fast_pow(nm, -m.kappa);                                  // dust term factor
                        spare[2 * kThreads] = m.A * mufu_ex2(-m.B * mufu_rcp(nm) * 1.44269504089f);    // pure-ice term
                        need = false;
                        if (SAVE_ALL) {
                            spare_tag_create = x_before;
                            if (alive) {
                                cur_interrupt_at = scatters;
                                cur_tag_resume = rng.x;
                            }
                        }
                        ++n_created;
                    }
                    spare_mask |= __ballot_sync(0xffffffffu, take);
                    const uint32_t wanted = __popc(need_mask);
                    w_left -= min(wanted, w_left);
                }
            }

            // ---------------------------------------------------------- pop
            const unsigned popping = ~alive_mask & spare_mask;
            if (popping & lane_bit) {
                const float *st = start0 + cur * (kStartWords * kThreads);
                pos.x = st[0 * kThreads]; pos.y = st[1 * kThreads]; pos.z = st[2 * kThreads];
                dir.x = st[4 * kThreads]; dir.y = st[5 * kThreads]; dir.z = st[6 * kThreads];
                abs_left = st[8 * kThreads];
                f_scat = spare[0 * kThreads]; f_dust = spare[1 * kThreads]; f_pure = spare[2 * kThreads];
                path = 0.f;
                scatters = 0;
                layer = min(max(__float2int_rz((pos.z - m.z0) * inv_h), 0), m.num_layers - 1);
                alive = true;
                if (SAVE_ALL) {
                    cur_tag_create = spare_tag_create;
                    cur_tag_pop = rng.x;
                    cur_interrupt_at = 0xffffffffu;
                }
            }
            spare_mask &= ~popping;
            alive_mask |= popping;
            if (alive_mask == 0u) {
                if (!w_more && w_left == 0) break; // nothing in flight, nothing spare, nothing to fetch
                continue;
            }
        }

        if (alive) {
            // -------------------------------------------------------------- R5: segment length
            float z_eff = pos.z;
            if (TILT) {
                z_eff = pos.z - tilt_shift(m, pos.x, pos.y, pos.z);
                layer = min(max(__float2int_rz((z_eff - m.z0) * inv_h), 0), m.num_layers - 1);
            }
            float inv_aniso = 1.f;
            if (ANISO) {
                // R4b: 1/f = (B2-nB)*An/2 (I3CLSimScalarFieldAnisotropyAbsLenScaling.cxx:92-134)
                const float n0 = m.azx * dir.x + m.azy * dir.y, n1 = m.neg_azy * dir.x + m.azx * dir.y;
                const float s0 = n0 * n0, s1 = n1 * n1, s2 = dir.z * dir.z;
                const float nB = s0 * m.rl[0] + s1 * m.rl[1] + s2 * m.rl[2];
                const float An = s0 * m.l[0] + s1 * m.l[1] + s2 * m.l[2];
                inv_aniso = (m.B2 - nB) * An * 0.5f;
                abs_left *= mufu_rcp(inv_aniso);
            }
            const float dz = dir.z;
            const float sca_left = -fast_ln(rng.oc());
            float4 c = sp.layers[layer];
            float b = c.x * f_scat;                       // 1/scattering length
            float a = c.y * f_dust + c.z * f_pure;        // 1/absorption length
            float boundary = m.z0 + m.h * static_cast<float>(layer + ((dz < 0.f) ? 0 : 1));
            float ais = (dz * sca_left - (boundary - z_eff) * b) * inv_h;
            float aia = (dz * abs_left - (boundary - z_eff) * a) * inv_h;
            int j = layer;
            if (dz < 0.f) {
                while ((j > 0) && (ais < 0.f) && (aia < 0.f)) {
                    --j;
                    boundary -= m.h;
                    c = sp.layers[j];
                    b = c.x * f_scat;
                    a = c.y * f_dust + c.z * f_pure;
                    ais += b;
                    aia += a;
                }
            } else {
                while ((j < m.num_layers - 1) && (ais > 0.f) && (aia > 0.f)) {
                    ++j;
                    boundary += m.h;
                    c = sp.layers[j];
                    b = c.x * f_scat;
                    a = c.y * f_dust + c.z * f_pure;
                    ais -= b;
                    aia -= a;
                }
            }
            float travel;
            if ((j == layer) || (fabsf(dz) < kEpsilon)) {
                // same layer: d_scatter = sca_left/b, d_absorb = abs_left/a; one reciprocal
                const bool absorbed = abs_left * b < sca_left * a;
                travel = (absorbed ? abs_left : sca_left) * mufu_rcp(absorbed ? a : b);
                abs_left = absorbed ? 0.f : abs_left - travel * a;
            } else {
                const float inv_dz = mufu_rcp(dz);
                const float base = boundary - z_eff;
                const float d_scatter = (ais * m.h * mufu_rcp(b) + base) * inv_dz;
                const float d_absorb = (aia * m.h * mufu_rcp(a) + base) * inv_dz;
                if (d_absorb < d_scatter) {
                    travel = d_absorb;
                    abs_left = 0.f;
                } else {
                    travel = d_scatter;
                    abs_left = (d_absorb - d_scatter) * a;
                }
            }
            if (!TILT) layer = j;
            if (ANISO) abs_left *= inv_aniso;
            ++n_segments;

            // -------------------------------------------------------------- R6: DOM collision
            Collision col{travel, 0, 0, false};
            if (!SAVE_ALL) {
                // distance-field early out: can this segment reach any string at all?
                const int px = min(max(__float2int_rz((pos.x - geo.near_x0) * geo.near_inv_pixel), 0), geo.near_nx - 1);
                const int py = min(max(__float2int_rz((pos.y - geo.near_y0) * geo.near_inv_pixel), 0), geo.near_ny - 1);
                const int pixel = py * geo.near_nx + px;
                const float reach = travel + geo.string_max_radius;
                if (reach >= static_cast<float>(sp.near_d1[pixel])) {
                    const float dir_xy2 = dir.x * dir.x + dir.y * dir.y;
                    if (dir_xy2 > 0.f) {
                        const float inv_xy2 = mufu_rcp(dir_xy2);
                        const uint32_t info = __ldg(geo.near_info + pixel);
                        if (reach < static_cast<float>((info >> 16) & 0xffu)) {
                            test_string(sp, geo, static_cast<int>(info & 0xffffu), pos, dir, inv_xy2, scene.inv_pancake_factor, col);
                        } else {
                            cell_walk(sp, geo, lay, pos, dir, inv_xy2, scene.inv_pancake_factor, col);
                        }
                    }
                }
            }
            bool emit = false;
            float emit_dist_abs = 0.f;
            if (col.hit) {
                // distInAbsLens is taken for the unshortened segment (propagation_kernel.c.cl:718)
                emit = true;
                emit_dist_abs = life_of_current(start0, cur) - abs_left;
                abs_left = 0.f;
            }
            travel = col.travel;

            // -------------------------------------------------------------- advance
            pos.x += dir.x * travel;
            pos.y += dir.y * travel;
            pos.z += dir.z * travel;
            path += travel;

            if (abs_left < kEpsilon) {
                alive = false;
                if (SAVE_ALL) {
                    // propagation_kernel.c.cl:800-826
                    if (rng.co() < scene.prescale) {
                        emit = true;
                        emit_dist_abs = life_of_current(start0, cur);
                    }
                }
                if (emit)
                    emit_record(scene, args, start0 + cur * (kStartWords * kThreads), pos, dir, path, scatters, col.string, col.dom,
                                emit_dist_abs, SAVE_ALL, cur_tag_create, cur_tag_pop, cur_tag_resume, cur_interrupt_at, rng.a);
                cur ^= 1; // the spare (if any) sits in the other buffer
            } else {
                // ---------------------------------------------------------- R9 + R8: scatter
                if (ANISO) apply_matrix(m.pre, dir);
                const float rr = rng.co();
                float cs;
                if (m.scat_kind == CLSIMCU_SCAT_MIXED_SL_HG) {
                    // both samplers are evaluated and one is selected: no divergent branch
                    const float cos_sl = 2.f * fast_pow(rr * inv_fsl, m.sl_beta) - 1.f;
                    const float s = 2.f * ((1.f - rr) * inv_omf) - 1.f;
                    const float ii = (1.f - m.g2) * mufu_rcp(1.f + m.g * s);
                    const float cos_hg = (1.f + m.g2 - ii * ii) * inv_2g;
                    cs = (rr < m.f_sl) ? cos_sl : cos_hg;
                } else if (m.scat_kind == CLSIMCU_SCAT_HG) {
                    const float s = 2.f * rr - 1.f;
                    const float ii = (1.f - m.g2) * mufu_rcp(1.f + m.g * s);
                    cs = (1.f + m.g2 - ii * ii) * inv_2g;
                } else {
                    cs = 2.f * fast_pow(rr, m.sl_beta) - 1.f;
                }
                cs = fminf(fmaxf(cs, -1.f), 1.f);
                const float sn = mufu_sqrt(1.f - cs * cs);
                rotate_by(cs, sn, dir, rng.co());
                if (ANISO) apply_matrix(m.post, dir);
                ++scatters;
            }
        }
    }

    args.rng_x[gthread] = rng.x;
    if (args.count_stats) {
        // warp-level reduction, one atomic per warp
        for (int o = 16; o > 0; o >>= 1) {
            n_created += __shfl_down_sync(0xffffffffu, n_created, o);
            n_segments += __shfl_down_sync(0xffffffffu, n_segments, o);
        }
        if (lane == 0) {
            atomicAdd(args.stats + 0, n_created);
            atomicAdd(args.stats + 1, n_segments);
        }
    }
}

template <bool TILT, bool ANISO, bool SAVE_ALL>
int launch_variant(const DevScene &scene, const LaunchArgs &args, int blocks, cudaStream_t stream)
{
    const SmemLayout lay = plan_smem(scene);
    auto kernel = propagate_persistent<TILT, ANISO, SAVE_ALL>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -3;
        configured = true;
    }
    kernel<<<blocks, kThreads, lay.total, stream>>>(scene, args, lay);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

} // namespace

bool fast_kernel_supports(const DevScene &scene, const char **why)
{
    static const char *k_history = "photon history is only recorded by the reference-order kernel";
    static const char *k_nonstop = "StopDetectedPhotons=false is only implemented by the reference-order kernel";
    static const char *k_renorm = "non-renormalising direction transforms are only implemented by the reference-order kernel";
    static const char *k_smem = "geometry/medium tables do not fit into shared memory";
    if (scene.history_entries > 0) { *why = k_history; return false; }
    if (!scene.save_all && !scene.stop_detected) { *why = k_nonstop; return false; }
    if (scene.medium.anisotropy && (!scene.medium.pre_renorm || !scene.medium.post_renorm)) { *why = k_renorm; return false; }
    if (plan_smem(scene).total > 200u * 1024u / kBlocksPerSM) { *why = k_smem; return false; }
    return true;
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
    const bool tilt = scene.medium.tilt_nd > 0, aniso = scene.medium.anisotropy != 0;
    if (scene.save_all) {
        if (tilt && aniso) return launch_variant<true, true, true>(scene, args, grid_blocks, stream);
        if (tilt) return launch_variant<true, false, true>(scene, args, grid_blocks, stream);
        if (aniso) return launch_variant<false, true, true>(scene, args, grid_blocks, stream);
        return launch_variant<false, false, true>(scene, args, grid_blocks, stream);
    }
    if (tilt && aniso) return launch_variant<true, true, false>(scene, args, grid_blocks, stream);
    if (tilt) return launch_variant<true, false, false>(scene, args, grid_blocks, stream);
    if (aniso) return launch_variant<false, true, false>(scene, args, grid_blocks, stream);
    return launch_variant<false, false, false>(scene, args, grid_blocks, stream);
}

} // namespace clsimcu
