// kernel_reference.cu -- reference-order propagation kernel for sm_100a.
//
// One thread per step, photons of a step created one after another from the step's own MWC
// stream: the reference's work-item mapping (resources/kernels/propagation_kernel.c.cl:406-913),
// so results can be compared photon by photon with a CPU run that uses the same (x, a) pairs.
// It evaluates the reference's precise-math formulation (no fast intrinsics, no fused
// multiply-add: this file is compiled with --fmad=false) and supports every option of the
// path, including the ones the fast kernel leaves out (photon history, non-stopping
// detection, save-all).  It is the slow, exact twin; kernel_fast.cu is the product path.
//
// Not a port of the OpenCL text: tables come from HBM through the read-only path instead of
// JIT-baked constants, the ID rewrite and wavelength-bias weight happen here, and hits are
// written as five 16-byte stores.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/clsimcuda.h"
#include "device_scene.h"
#include "tabulate_device.h"

namespace clsimcu {
namespace {

constexpr float kSpeedOfLight = 0.299792458f; // propagation_kernel.h.cl:144-151
constexpr float kPi = 3.14159265359f;
constexpr float kEpsilon = 0.00001f;          // propagation_kernel.c.cl:505
constexpr int kMaxHistory = 32;

struct Stream {
    uint64_t x;
    uint32_t a;
    // mwcrng_kernel.cl:12-28
    __device__ float co()
    {
        x = (x & 0xffffffffull) * a + (x >> 32);
        return __uint2float_rz(static_cast<uint32_t>(x)) * 2.3283064365386963e-10f; // exact /2^32
    }
    __device__ float oc() { return 1.0f - co(); }
};

struct V3 {
    float x, y, z;
};

struct Flight {
    V3 pos;
    float t;
    V3 dir;
    float wlen;
    V3 start_pos;
    float start_t;
    V3 start_dir;
    uint32_t scatters;
    float path;
    float inv_vg;
    float abs_initial;
};

__device__ __forceinline__ float sq(float v) { return v * v; }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ float cl_sign(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : ((v != v) ? 0.f : v)); }

// ---- medium (R4) ---------------------------------------------------------------------------
__device__ float phase_index(const DevMedium &m, float wlen)
{
    const float u = wlen / 1e-6f;
    return m.n_phase[0] + u * (m.n_phase[1] + u * (m.n_phase[2] + u * (m.n_phase[3] + u * m.n_phase[4])));
}
__device__ float group_velocity(const DevMedium &m, float wlen)
{
    const float u = wlen / 1e-6f;
    const float np = m.n_phase[0] + u * (m.n_phase[1] + u * (m.n_phase[2] + u * (m.n_phase[3] + u * m.n_phase[4])));
    const float corr = m.n_group[0] + u * (m.n_group[1] + u * (m.n_group[2] + u * (m.n_group[3] + u * m.n_group[4])));
    return m.c_light / (np * corr);
}
__device__ float scat_len(const DevMedium &m, int layer, float wlen)
{
    return 1.f / (__ldg(m.b400 + layer) * powf(wlen * m.inv_ref_wlen, -m.alpha));
}
__device__ float abs_len(const DevMedium &m, int layer, float wlen)
{
    const float u = wlen / 1e-9f;
    return 1.f / ((m.D * __ldg(m.a_dust400 + layer) + m.E) * powf(u, -m.kappa) + m.A * expf(-m.B / u) * (1.f + 0.01f * __ldg(m.delta_tau + layer)));
}
__device__ int layer_of(const DevMedium &m, float z) { return static_cast<int>((z - m.z0) / m.h); }
__device__ float layer_floor(const DevMedium &m, int layer) { return (static_cast<float>(layer) * m.h) + m.z0; }

// R4a
__device__ float tilt_shift(const DevMedium &m, const V3 &p)
{
    const float zr = (p.z - m.tilt_z0) / m.tilt_dz;
    const int k = clampi(static_cast<int>(floorf(zr)), 0, m.tilt_nz - 2);
    const float above = zr - static_cast<float>(k);
    const float below = 1.f - above;
    const float nr = m.tilt_lnx * p.x + m.tilt_lny * p.y;
    for (int j = 1; j < m.tilt_nd; j++) {
        const float here = __ldg(m.tilt_dist + j);
        if ((nr < here) || (j == m.tilt_nd - 1)) {
            const float prev = __ldg(m.tilt_dist + j - 1);
            const float width = here - prev;
            const float w_lo = (here - nr) / width;
            const float w_hi = 1.f - w_lo;
            const float v_lo = (__ldg(m.tilt_corr + (j - 1) * m.tilt_nz + k + 1) * above + __ldg(m.tilt_corr + (j - 1) * m.tilt_nz + k) * below);
            const float v_hi = (__ldg(m.tilt_corr + j * m.tilt_nz + k + 1) * above + __ldg(m.tilt_corr + j * m.tilt_nz + k) * below);
            return (v_hi * w_hi + v_lo * w_lo);
        }
    }
    return 0.f;
}

// R4b
__device__ float abs_len_scaling(const DevMedium &m, const V3 &d)
{
    if (!m.anisotropy) return 1.f;
    const float n0 = (m.azx * d.x) + (m.azy * d.y);
    const float n1 = (m.neg_azy * d.x) + (m.azx * d.y);
    const float s0 = n0 * n0, s1 = n1 * n1, s2 = d.z * d.z;
    const float nB = ((s0 * m.rl[0] + s1 * m.rl[1]) + s2 * m.rl[2]) + 0.f;
    const float An = ((s0 * m.l[0] + s1 * m.l[1]) + s2 * m.l[2]) + 0.f;
    return 2.f / ((m.B2 - nB) * An);
}
__device__ void apply_matrix(const float *M, int renorm, V3 &d)
{
    const float nx = (M[0] * d.x) + (M[1] * d.y) + (M[2] * d.z);
    const float ny = (M[3] * d.x) + (M[4] * d.y) + (M[5] * d.z);
    const float nz = (M[6] * d.x) + (M[7] * d.y) + (M[8] * d.z);
    d.x = nx; d.y = ny; d.z = nz;
    if (renorm) {
        const float inv = 1.f / sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        d.x = d.x * inv; d.y = d.y * inv; d.z = d.z * inv;
    }
}

// R9
__device__ float hg_cos(const DevMedium &m, float u)
{
    const float s = 2.f * (u) - 1.f;
    const float ii = ((1.f - m.g2) / (1.f + m.g * s));
    return fminf(fmaxf((1.f + m.g2 - ii * ii) / (2.f * m.g), -1.f), 1.f);
}
__device__ float sl_cos(const DevMedium &m, float u) { return fminf(fmaxf(2.f * powf((u), m.sl_beta) - 1.f, -1.f), 1.f); }
__device__ float scatter_cos(const DevMedium &m, Stream &rng)
{
    if (m.scat_kind == CLSIMCU_SCAT_HG) return hg_cos(m, rng.co());
    if (m.scat_kind == CLSIMCU_SCAT_SL) return sl_cos(m, rng.co());
    const float rr = rng.co();
    if (rr < m.f_sl) return sl_cos(m, rr / m.f_sl);
    return hg_cos(m, (1.f - rr) / m.one_minus_f_sl);
}

// R8
__device__ void rotate_by(float cosa, float sina, V3 &d, float rnd)
{
    const float b = 2.0f * kPi * rnd;
    const float cosb = cosf(b), sinb = sinf(b);
    const float sinth = sqrtf(fmaxf(0.f, 1.f - d.z * d.z));
    if (sinth > 0.f) {
        const V3 o = d;
        d.x = o.x * cosa - ((o.y * cosb + o.z * o.x * sinb) * sina) / sinth;
        d.y = o.y * cosa + ((o.x * cosb - o.z * o.y * sinb) * sina) / sinth;
        d.z = o.z * cosa + sina * sinb * sinth;
    } else {
        d.x = sina * cosb;
        d.y = sina * sinb;
        d.z = cosa * cl_sign(d.z);
    }
    const float inv = 1.f / sqrtf(sq(d.x) + sq(d.y) + sq(d.z));
    d.x *= inv; d.y *= inv; d.z *= inv;
}

// R3a
__device__ float draw_wavelength(const DevWlenGenerator &g, Stream &rng)
{
    if (g.kind == CLSIMCU_WLEN_CONSTANT) return g.value;
    if (g.kind == CLSIMCU_WLEN_NO_DISPERSION) {
        const float r = rng.oc();
        return 1.f / (g.min_val + r * g.range);
    }
    const float r = rng.oc();
    int k = 0;
    float below = 0.f;
    for (;;) {
        const float next = __ldg(g.cumulative + k + 1);
        if (next >= r) break;
        below = next;
        if (k + 2 >= g.n) break; // table ends at 1.0 >= r; guard against reading past it
        ++k;
    }
    const float b = __ldg(g.density + k);
    float x0, slope;
    if (g.kind == CLSIMCU_WLEN_INTERP_UNEQUAL) {
        x0 = __ldg(g.xs + k);
        slope = (__ldg(g.density + k + 1) - b) / (__ldg(g.xs + k + 1) - x0);
    } else {
        x0 = static_cast<float>(k) * (g.dx) + (g.x0);
        slope = (__ldg(g.density + k + 1) - b) / (g.dx);
    }
    const float dy = r - below;
    if ((b == 0.f) && (slope == 0.f)) return x0;
    if (b == 0.f) return x0 + sqrtf(2.f * dy / slope);
    if (slope == 0.f) return x0 + dy / b;
    return x0 + (sqrtf(dy * (2.f * slope) / (b * b) + 1.f) - 1.f) * b / slope;
}

// R3b
__device__ float bias_at(const DevBias &b, float wlen)
{
    if (b.kind == CLSIMCU_BIAS_CONSTANT) return b.value;
    float whole;
    float frac = modff((wlen - b.x0) / b.dx, &whole);
    int bin = static_cast<int>(whole);
    if ((bin < 0) || ((bin == 0) && (frac < 0))) {
        bin = 0;
        frac = 0.f;
    } else if (bin >= b.n - 1) {
        bin = b.n - 2;
        frac = 1.f;
    }
    const float lo = __ldg(b.v + bin), hi = __ldg(b.v + bin + 1);
    return lo + (hi - lo) * frac;
}

__device__ void to_spherical(const V3 &c, float &theta, float &phi)
{
    const float inv = 1.f / sqrtf(c.x * c.x + c.y * c.y + c.z * c.z);
    theta = 0.f;
    if (fabsf(c.z * inv) <= 1.f) theta = acosf(c.z * inv);
    else if (c.z < 0.f) theta = kPi;
    if (theta < 0.f) theta += 2.f * kPi;
    phi = atan2f(c.y, c.x);
    if (phi < 0.f) phi += 2.f * kPi;
}

__device__ void dom_centre(const DevGeometry &g, int string, int dom, float &x, float &y, float &z)
{
    const uint32_t at = __ldg(g.string_tmpl_start + string) + static_cast<uint32_t>(dom);
    x = static_cast<float>(__ldg(g.tmpl_dx + at)) * g.tmpl_scale_x + __ldg(g.string_mean_x + string);
    y = static_cast<float>(__ldg(g.tmpl_dy + at)) * g.tmpl_scale_y + __ldg(g.string_mean_y + string);
    z = __ldg(g.tmpl_z + at);
}

struct Output {
    const DevScene &scene;
    const LaunchArgs &args;
    const clsimcu_step &step;
    const float *ring; // current photon's history ring
};

// R10
__device__ void record_hit(const Output &o, const Flight &f, float length, float dist_abs, int string, int dom)
{
    const DevScene &sc = o.scene;
    const uint32_t slot = atomicAdd(o.args.hit_counter, 1u);
    if (slot >= o.args.max_hits) return;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    int16_t sid = 0;
    uint16_t oid = 0;
    if (!sc.save_all) {
        dom_centre(sc.geo, string, dom, cx, cy, cz);
        if (sc.pancake) {
            const float px = f.pos.x - cx, py = f.pos.y - cy, pz = f.pos.z - cz;
            const float along = px * f.dir.x + py * f.dir.y + pz * f.dir.z;
            const float nx = px - along * f.dir.x, ny = py - along * f.dir.y, nz = pz - along * f.dir.z;
            const float k = ((sc.pancake_factor - 1.f) / sc.pancake_factor);
            cx += k * nx; cy += k * ny; cz += k * nz;
        }
        sid = __ldg(sc.geo.string_index_to_id + string);
        oid = __ldg(sc.geo.dom_ids + __ldg(sc.geo.dom_id_offset + string) + dom);
    }
    float th, ph, sth, sph;
    to_spherical(f.dir, th, ph);
    to_spherical(f.start_dir, sth, sph);
    float4 *dst = reinterpret_cast<float4 *>(static_cast<clsimcu_photon *>(o.args.photons) + slot);
    dst[0] = make_float4(f.pos.x + length * f.dir.x - cx, f.pos.y + length * f.dir.y - cy, f.pos.z + length * f.dir.z - cz,
                         f.t + length * f.inv_vg);
    dst[1] = make_float4(th, ph, f.wlen, f.path + length);
    const uint32_t ids = (static_cast<uint32_t>(static_cast<uint16_t>(sid))) | (static_cast<uint32_t>(oid) << 16);
    dst[2] = make_float4(__uint_as_float(f.scatters), o.step.weight / bias_at(sc.bias, f.wlen), __uint_as_float(o.step.identifier),
                         __uint_as_float(ids));
    dst[3] = make_float4(f.start_pos.x, f.start_pos.y, f.start_pos.z, f.start_t);
    dst[4] = make_float4(sth, sph, 1.f / f.inv_vg, dist_abs);
    if (sc.history_entries > 0 && o.args.history) {
        float4 *h = reinterpret_cast<float4 *>(o.args.history) + static_cast<size_t>(slot) * sc.history_entries;
        for (int i = 0; i < sc.history_entries; ++i) h[i] = make_float4(o.ring[4 * i], o.ring[4 * i + 1], o.ring[4 * i + 2], o.ring[4 * i + 3]);
    }
}

struct Nearest {
    bool found;
    int string, dom;
};

// R6, string level (sparse_collision_kernel.c.cl:27-192)
__device__ void test_string(const Output &o, const Flight &f, int string, float dir_xy2, float &length, float dist_abs, Nearest &hit)
{
    const DevGeometry &g = o.scene.geo;
    const int set = __ldg(g.string_set + string);
    {
        const float miss2 = sq(((f.pos.x - __ldg(g.string_x + string)) * f.dir.y - (f.pos.y - __ldg(g.string_y + string)) * f.dir.x)) / dir_xy2;
        if (miss2 > sq(g.string_max_radius)) return; // global radius on purpose (quirk 7)
    }
    if ((f.dir.z > 0.f) && (f.pos.z > __ldg(g.string_max_z + string) + g.om_radius)) return;
    if ((f.dir.z < 0.f) && (f.pos.z < __ldg(g.string_min_z + string) - g.om_radius)) return;

    const float z_start = __ldg(g.set_start_z + set), z_height = __ldg(g.set_layer_height + set);
    const int layers = __ldg(g.set_layer_count + set);
    int lo = static_cast<int>((f.pos.z - z_start) / z_height);
    int hi = static_cast<int>((f.pos.z + f.dir.z * length - z_start) / z_height);
    if (hi < lo) { const int tmp = lo; lo = hi; hi = tmp; }
    lo = clampi(lo, 0, layers - 1);
    hi = clampi(hi, 0, layers - 1);

    // Non-stop mode: the reference keeps a bit mask of the DOMs it has tested (sparse_collision_kernel.c.cl:85-104) because
    // the z-layer table names a DOM in EVERY layer its sphere touches (two for a DOM that straddles a boundary).  Those
    // layers are adjacent and a layer holds one DOM at most, so "the same DOM as in the layer before" is the whole
    // mask -- without the mask's limit on the number of DOMs per string.
    const uint16_t *row = g.layer_to_dom + static_cast<uint32_t>(set) * g.max_layers;
    int previous = 0xFFFF;
    for (int layer = lo; layer <= hi; ++layer) {
        const int dom = __ldg(row + layer);
        if (dom == 0xFFFF) continue;
        if (!o.scene.stop_detected && dom == previous) continue;
        previous = dom;
        float cx, cy, cz;
        dom_centre(g, string, dom, cx, cy, cz);
        const float rx = cx - f.pos.x, ry = cy - f.pos.y, rz = cz - f.pos.z;
        const float r2 = ((rx * rx + ry * ry) + rz * rz) + 0.f;
        const float along = ((rx * f.dir.x + ry * f.dir.y) + rz * f.dir.z) + 0.f;
        float disc = sq(along) - r2 + g.om_radius * g.om_radius;
        if (disc < 0.f) continue;
        disc = o.scene.pancake ? sqrtf(disc) / o.scene.pancake_factor : sqrtf(disc);
        if (along + disc < 0.f) continue;
        const float entry = along - disc;
        if (entry < 0.f) continue; // started inside: let it out (quirk 9)
        if (entry < length) {
            if (o.scene.stop_detected) {
                length = entry;
                hit.found = true;
                hit.string = string;
                hit.dom = dom;
            } else {
                record_hit(o, f, entry, dist_abs, string, dom);
            }
        }
    }
}

// R6, cell level and driver (sparse_collision_kernel.c.cl:194-303, 462-587)
__device__ bool find_collision(const Output &o, const Flight &f, float &length, float dist_abs)
{
    const DevGeometry &g = o.scene.geo;
    const float dir_xy2 = sq(f.dir.x) + sq(f.dir.y);
    if (dir_xy2 <= 0.f) return false;
    Nearest hit{false, 0, 0};
    for (int s = 0; s < g.num_grids; ++s) {
        const DevCellGrid &c = g.grids[s];
        int x0 = static_cast<int>((f.pos.x - c.start_x) / c.width_x);
        int y0 = static_cast<int>((f.pos.y - c.start_y) / c.width_y);
        int x1 = static_cast<int>((f.pos.x + f.dir.x * length - c.start_x) / c.width_x);
        int y1 = static_cast<int>((f.pos.y + f.dir.y * length - c.start_y) / c.width_y);
        if (x1 < x0) { const int tmp = x0; x0 = x1; x1 = tmp; }
        if (y1 < y0) { const int tmp = y0; y0 = y1; y1 = tmp; }
        x0 = clampi(x0, 0, c.num_x - 1);
        y0 = clampi(y0, 0, c.num_y - 1);
        x1 = clampi(x1, 0, c.num_x - 1);
        y1 = clampi(y1, 0, c.num_y - 1);
        const bool dedup = !o.scene.stop_detected;
        for (int cy = y0; cy <= y1; ++cy) {
            for (int cx = x0; cx <= x1; ++cx) {
                const int string = __ldg(c.cell_to_string + cy * c.num_x + cx);
                if (string == 0xFFFF) continue;
                if (dedup) {
                    // a string may sit in several cells and must be tested once per segment (the reference's string bit
                    // mask, :250-269, as intended -- SURVEY quirk 8): skip it if an earlier cell of this walk named it.
                    // No mask, so no limit on the number of strings.
                    bool seen = false;
                    for (int py = y0; py <= cy && !seen; ++py)
                        for (int px = x0; px <= ((py < cy) ? x1 : cx - 1) && !seen; ++px) seen = __ldg(c.cell_to_string + py * c.num_x + px) == string;
                    if (seen) continue;
                }
                test_string(o, f, string, dir_xy2, length, dist_abs, hit);
            }
        }
    }
    if (o.scene.stop_detected && hit.found) {
        record_hit(o, f, length, dist_abs, hit.string, hit.dom);
        return true;
    }
    return false;
}

// ---- table-maker variant (-DTABULATE) -------------------------------------------------------------------
// The coordinate and binning code is shared with the fast kernel: tabulate_device.h.  Here: the fifth (impact angle)
// coordinate, which draws from the work item's stream (spherical_coordinates.c.cl / cylindrical_coordinates.c.cl).
// OpenCL's dot() of two float4 includes the fourth components: zero for the reference vectors, but wavelength x delay
// time in the impact-angle product (a quirk of the reference, kept).
__device__ void table_coordinates(const TabulateArgs &tb, const V3 &p, float t, V3 dir, float wlen, Stream &rng, float c[5])
{
    const TableFrame fr = table_frame(tb, p.x, p.y, p.z, t);
    table_coordinates_4(tb, fr, c);
    if (tb.ndim <= 4) return;
    const float px = fr.px, py = fr.py, pz = fr.pz, pw = fr.pw, l = fr.l, rx = fr.rx, ry = fr.ry, rz = fr.rz, rw = fr.rw;
    if (tb.geometry == 0) {
        const float sina = sqrtf(rng.co());
        rotate_by(sqrtf(1 - sina * sina), sina, dir, rng.co());
        c[4] = (c[0] > 0) ? ((((dir.x * px + dir.y * py) + dir.z * pz) + wlen * pw) / c[0]) : 1;
    } else {
        const float sina = sqrtf(rng.co());
        rotate_by(sqrtf(1 - sina * sina), sina, dir, rng.co());
        // vector from the nominal Cherenkov emission point; (l - rho/tan_thetaC) is a float4 in the reference, component-wise
        const float k = 1.f / tb.tan_theta_c;
        const float qx = p.x - (tb.ref_pos[0] + (l - rx * k) * tb.ref_dir[0]), qy = p.y - (tb.ref_pos[1] + (l - ry * k) * tb.ref_dir[1]),
                    qz = p.z - (tb.ref_pos[2] + (l - rz * k) * tb.ref_dir[2]), qw = t - (tb.ref_pos[3] + (l - rw * k) * tb.ref_dir[3]);
        const float cdist = sqrtf(qx * qx + qy * qy + qz * qz);
        c[4] = (cdist > 0) ? ((((dir.x * qx + dir.y * qy) + dir.z * qz) + wlen * qw) / cdist) : 1;
    }
}

// savePath (propagation_kernel.c.cl:226-304).  The entries go straight into the table in HBM; there is no entry
// buffer to run out of, so the reference's "return false, restart the photon" branch does not exist.
__device__ void save_path(const TabulateArgs &tb, const clsimcu_step &step, const Flight &f, float travel, float &prev_remainder, float depth,
                          float step_depth, bool &stop, Stream &rng)
{
    const float impact_weight = (tb.ndim > 4) ? step.weight : step.weight * table_angular_acceptance(tb, f.dir.z);
    float d = prev_remainder;
    for (; d < travel; d += tb.step_length) {
        V3 p;
        p.x = f.pos.x + d * f.dir.x;
        p.y = f.pos.y + d * f.dir.y;
        p.z = f.pos.z + d * f.dir.z;
        const float t = f.t + d * f.inv_vg;
        float c[5];
        table_coordinates(tb, p, t, f.dir, f.wlen, rng, c);
        const bool out = (tb.geometry == 0) ? ((c[3] > tb.max3) || (c[0] > tb.max0)) : (c[3] > tb.max3);
        if (out) {
            stop = true;
            break;
        }
        const uint32_t index = table_bin_index(tb, c);
        const float w = impact_weight * expf(-(depth + (d / travel) * step_depth));
        atomicAdd(tb.table + index, w);
        if (tb.squared) atomicAdd(tb.squared + index, w * w);
    }
    prev_remainder = d - travel;
}

__global__ void __launch_bounds__(64) propagate_reference_order(const __grid_constant__ DevScene scene, const __grid_constant__ LaunchArgs args)
{
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= args.num_steps) return;
    const DevMedium &m = scene.medium;

    Stream rng{args.rng_x[item], args.rng_a[item]};
    const clsimcu_step step = static_cast<const clsimcu_step *>(args.steps)[item];

    V3 axis;
    {
        const float rho = sinf(step.theta);
        axis.x = rho * cosf(step.phi);
        axis.y = rho * sinf(step.phi);
        axis.z = cosf(step.theta);
    }

    float ring[4 * kMaxHistory];
    for (int i = 0; i < 4 * scene.history_entries; ++i) ring[i] = 0.f;
    const Output out{scene, args, step, ring};

    uint32_t left = step.num_photons;
    unsigned long long n_created = 0, n_segments = 0;
    {
        // the one place where this kernel leaves the reference's order: a step that is not a finite number (or beyond 2^63)
        // is counted and ended here, as in the fast kernel (kernel_fast.cu, fill_queue).  The reference's text either flies
        // such photons through its outermost layer or -- inf - inf in the absorption budget -- never ends them.
        const uint32_t *w = reinterpret_cast<const uint32_t *>(&step);
        bool out_of_this_world = false;
        for (int i = 0; i < 8; ++i) out_of_this_world |= (w[i] & 0x7f800000u) >= 0x5f000000u;
        if (out_of_this_world) {
            n_created = left;
            left = 0;
        }
    }
    float abs_left = 0.f;
    Flight f;
    f.scatters = 0;
    f.path = 0.f;
    f.inv_vg = 0.f;
    f.abs_initial = 0.f;
    const bool flat_ice = (m.tilt_nd == 0); // getTiltZShift_IS_CONSTANT
    int layer = 0;
    const TabulateArgs *tab = args.tabulate;   // table-maker variant
    float depth = 0.f, prev_remainder = 0.f;

    while (left > 0) {
        if (abs_left < kEpsilon) {
            // R3: createPhotonFromTrack
            const float shift = step.length * rng.co();
            const float inv_speed = 1.f / (kSpeedOfLight * step.beta);
            f.pos.x = step.x + axis.x * shift;
            f.pos.y = step.y + axis.y * shift;
            f.pos.z = step.z + axis.z * shift;
            f.t = step.t + inv_speed * shift;
            f.dir = axis;
            if (scene.num_generators <= 1 || step.source_type == 0) {
                f.wlen = draw_wavelength(scene.generators[0], rng);
                const float cos_c = fminf(1.f, 1.f / (step.beta * phase_index(m, f.wlen)));
                const float sin_c = sqrtf(1.f - cos_c * cos_c);
                rotate_by(cos_c, sin_c, f.dir, rng.co());
            } else {
                const int k = step.source_type;
                f.wlen = (k < scene.num_generators) ? draw_wavelength(scene.generators[k], rng) : 0.f;
            }
            f.start_pos = f.pos;
            f.start_t = f.t;
            f.start_dir = f.dir;
            f.scatters = 0;
            f.path = 0.f;
            if (tab) {
                // randomise the first sub-step (propagation_kernel.c.cl:566-569); depth restarts (:590-592)
                prev_remainder = tab->step_length * rng.oc();
                depth = 0.f;
            }
            if (flat_ice) layer = clampi(layer_of(m, f.pos.z), 0, m.num_layers - 1);
            f.inv_vg = 1.f / group_velocity(m, f.wlen);
            f.abs_initial = scene.fixed_abs ? scene.fixed_abs_lens : -logf(rng.oc());
            abs_left = f.abs_initial;
            ++n_created;
        }

        // R5: distance to the next scatter / absorption through the layers
        float travel;
        {
            float z_eff;
            if (flat_ice) {
                z_eff = f.pos.z - 0.f;
            } else {
                z_eff = f.pos.z - tilt_shift(m, f.pos);
                layer = clampi(layer_of(m, z_eff), 0, m.num_layers - 1);
            }
            const float dz = f.dir.z;
            const float aniso = abs_len_scaling(m, f.dir);
            abs_left *= aniso;
            float boundary = (dz < 0.f) ? (layer_floor(m, layer)) : (layer_floor(m, layer) + m.h);
            const float sca_left = -logf(rng.oc());
            float ls = scat_len(m, layer, f.wlen);
            float la = abs_len(m, layer, f.wlen);
            float ais = (dz * sca_left - ((boundary - z_eff) / ls)) * (1.f / m.h);
            float aia = (dz * abs_left - ((boundary - z_eff) / la)) * (1.f / m.h);
            int j = layer;
            if (dz < 0) {
                while ((j > 0) && (ais < 0.f) && (aia < 0.f)) {
                    --j;
                    boundary -= m.h;
                    ls = scat_len(m, j, f.wlen);
                    la = abs_len(m, j, f.wlen);
                    ais += 1.f / ls;
                    aia += 1.f / la;
                }
            } else {
                while ((j < m.num_layers - 1) && (ais > 0.f) && (aia > 0.f)) {
                    ++j;
                    boundary += m.h;
                    ls = scat_len(m, j, f.wlen);
                    la = abs_len(m, j, f.wlen);
                    ais -= 1.f / ls;
                    aia -= 1.f / la;
                }
            }
            float to_absorption;
            if ((layer == j) || (fabsf(dz) < kEpsilon)) {
                travel = sca_left * ls;
                to_absorption = abs_left * la;
            } else {
                const float inv_dz = 1.f / dz;
                travel = (ais * m.h * ls + boundary - z_eff) * inv_dz;
                to_absorption = (aia * m.h * la + boundary - z_eff) * inv_dz;
            }
            if (flat_ice) layer = j;
            if (to_absorption < travel) {
                travel = to_absorption;
                abs_left = 0.f;
            } else {
                abs_left = (to_absorption - travel) / la;
            }
            abs_left = abs_left / aniso;
        }
        ++n_segments;

        if (!scene.save_all) {
            const bool caught = find_collision(out, f, travel, f.abs_initial - abs_left);
            if (scene.stop_detected && caught) abs_left = 0.f;
        }
        if (tab) {
            // propagation_kernel.c.cl:755-785
            bool stop = false;
            save_path(*tab, step, f, travel, prev_remainder, depth, f.abs_initial - abs_left - depth, stop, rng);
            if (stop) abs_left = 0.f;   // the photon ran off the end of the table
            depth = f.abs_initial - abs_left;
        }

        f.pos.x += f.dir.x * travel;
        f.pos.y += f.dir.y * travel;
        f.pos.z += f.dir.z * travel;
        f.t += f.inv_vg * travel;
        f.path += travel;

        if (abs_left < kEpsilon) {
            --left;
            if (scene.save_all && !tab) {   // #if defined(SAVE_ALL_PHOTONS) && !defined(TABULATE), :800
                if (rng.co() < scene.prescale) record_hit(out, f, 0.f, f.abs_initial, 0, 0);
            }
        } else {
            if (scene.history_entries > 0) {
                float *row = ring + 4 * (f.scatters % scene.history_entries);
                row[0] = f.pos.x; row[1] = f.pos.y; row[2] = f.pos.z;
                row[3] = f.abs_initial - abs_left;
            }
            if (m.anisotropy) apply_matrix(m.pre, m.pre_renorm, f.dir);
            const float cs = scatter_cos(m, rng);
            const float sn = sqrtf(1.f - sq(cs));
            rotate_by(cs, sn, f.dir, rng.co());
            if (m.anisotropy) apply_matrix(m.post, m.post_renorm, f.dir);
            ++f.scatters;
        }
    }

    args.rng_x[item] = rng.x;
    if (args.count_stats) {
        atomicAdd(args.stats + 0, n_created);
        atomicAdd(args.stats + 1, n_segments);
    }
}

} // namespace

int launch_reference_kernel(const DevScene &scene, const LaunchArgs &args, void *stream)
{
    if (args.num_steps == 0) return 0;
    if (scene.history_entries > kMaxHistory) return -2;
    const int threads = 64;
    const unsigned blocks = (args.num_steps + threads - 1) / threads;
    propagate_reference_order<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(scene, args);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;   // (peek: the caller reads the error text)
}

} // namespace clsimcu
