// tabulate_device.h -- device code of the table-maker variant shared by the two kernels: where a point of a photon's
// path falls in the table around the reference particle, and what it weighs.
//
// Restated from resources/kernels/spherical_coordinates.c.cl, cylindrical_coordinates.c.cl (getCoordinates), the code
// the reference generates from its Axes (private/clsim/tabulator/Axes.cxx:71-93 getBinIndex, Axis.cxx:44-60
// GetIndexCode) and I3CLSimFunctionPolynomial.cxx:139-153 (getAngularAcceptance); precise math in both kernels, so
// that a point lands in the same bin whichever kernel propagated the photon.
#pragma once

#include <cstdint>

#include "device_scene.h"

namespace clsimcu {

// FAST = true (the persistent kernel): special-function-unit root and reciprocal (2^-22 relative) and a polynomial arc
// cosine (Abramowitz & Stegun 4.4.46, |error| < 2e-8 rad) in place of the correctly rounded library functions, which are
// 60 % of the table-maker kernel's instructions (ncu r02_v35_tab).  A point within ~1e-6 of a bin edge can land in the
// neighbouring bin; the reference-order kernel (FAST = false) keeps the precise forms and stays comparable with the oracle
// bin by bin.
template <bool FAST> __device__ __forceinline__ float tab_sqrt(float x)
{
    if (FAST) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    return sqrtf(x);
}
template <bool FAST> __device__ __forceinline__ float tab_div(float a, float b)
{
    if (FAST) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return a * r; }
    return a / b;
}
template <bool FAST> __device__ __forceinline__ float tab_acos(float x)
{
    if (!FAST) return acosf(x);
    const float a = fminf(fabsf(x), 1.f);
    float p = -0.0012624911f;
    p = fmaf(p, a, 0.0066700901f); p = fmaf(p, a, -0.0170881256f); p = fmaf(p, a, 0.0308918810f); p = fmaf(p, a, -0.0501743046f);
    p = fmaf(p, a, 0.0889789874f); p = fmaf(p, a, -0.2145988016f); p = fmaf(p, a, 1.5707963050f);
    const float r = tab_sqrt<true>(1.f - a) * p;
    return (x < 0.f) ? 3.14159265359f - r : r;
}

// the point relative to the reference particle: p - ref (with time), its component along the particle's axis, the rest
struct TableFrame {
    float px, py, pz, pw, l, rx, ry, rz, rw, n_rho, rho_perp;
};

template <bool FAST = false>
__device__ inline TableFrame table_frame(const TabulateArgs &tb, float x, float y, float z, float t)
{
    TableFrame f;
    f.px = x - tb.ref_pos[0]; f.py = y - tb.ref_pos[1]; f.pz = z - tb.ref_pos[2]; f.pw = t - tb.ref_pos[3];
    f.l = ((f.px * tb.ref_dir[0] + f.py * tb.ref_dir[1]) + f.pz * tb.ref_dir[2]) + f.pw * tb.ref_dir[3];
    f.rx = f.px - f.l * tb.ref_dir[0]; f.ry = f.py - f.l * tb.ref_dir[1]; f.rz = f.pz - f.l * tb.ref_dir[2]; f.rw = f.pw - f.l * tb.ref_dir[3];
    f.n_rho = tab_sqrt<FAST>(f.rx * f.rx + f.ry * f.ry + f.rz * f.rz);
    f.rho_perp = ((f.rx * tb.ref_perp[0] + f.ry * tb.ref_perp[1]) + f.rz * tb.ref_perp[2]) + f.rw * tb.ref_perp[3];
    return f;
}

// the four coordinates every table has: (r, azimuth, cos polar angle, delay time) or (rho, azimuth, z, delay time)
template <bool FAST = false>
__device__ inline void table_coordinates_4(const TabulateArgs &tb, const TableFrame &f, float c[5])
{
    const float kPiOver180 = 3.14159265359f / 180;
    if (tb.geometry == 0) {
        c[0] = tab_sqrt<FAST>(f.px * f.px + f.py * f.py + f.pz * f.pz);
        const float azimuth = (f.n_rho > 0) ? (FAST ? tab_acos<FAST>(tab_div<FAST>(f.rho_perp, f.n_rho)) * (1.f / kPiOver180) : acosf(f.rho_perp / f.n_rho) / kPiOver180) : 0;
        if (tb.full_azimuth) {
            // cross(rho, perpDir) . dir
            const float cx = f.ry * tb.ref_perp[2] - f.rz * tb.ref_perp[1], cy = f.rz * tb.ref_perp[0] - f.rx * tb.ref_perp[2],
                        cz = f.rx * tb.ref_perp[1] - f.ry * tb.ref_perp[0];
            const float sign = (cx * tb.ref_dir[0] + cy * tb.ref_dir[1]) + cz * tb.ref_dir[2];
            c[1] = (sign > 0) ? 360.f - azimuth : azimuth;
        } else {
            c[1] = azimuth;
        }
        c[2] = (c[0] > 0) ? tab_div<FAST>(f.l, c[0]) : 0;
        c[3] = f.pw - c[0] * tb.min_inv_group_vel;
    } else {
        c[0] = f.n_rho;
        c[1] = (c[0] > 0) ? tab_acos<FAST>(tab_div<FAST>(f.rho_perp, c[0])) : 0;
        c[2] = tb.ref_pos[2] + f.l * tb.ref_dir[2];
        c[3] = f.pw - (f.l + c[0] * tb.tan_theta_c) * 3.33564095f;   // recip_speedOfLight, propagation_kernel.h.cl:149
    }
}

// ... of a four-axis table, unrolled (the coordinates stay in registers).  FAST: the axes of the tables in use are linear
// or quadratic (inverse transform: identity or root) -- selected without a branch; cube roots and general powers take the
// library functions behind a branch that is uniform over the warp (the axes are the table's, not the point's).
template <bool FAST = false>
__device__ inline uint32_t table_bin_index_4(const TabulateArgs &tb, const float c[4])
{
    if (FAST && tb.simple4) {
        // linear and quadratic axes only (the layouts in use): no branch, the axes' numbers in five 16-byte loads
        const float4 scale = *reinterpret_cast<const float4 *>(tb.scale4), shift = *reinterpret_cast<const float4 *>(tb.neg_offset4);
        const float4 root = *reinterpret_cast<const float4 *>(tb.root4);
        const int4 bins = *reinterpret_cast<const int4 *>(tb.n_bins4);
        const uint4 stride = *reinterpret_cast<const uint4 *>(tb.stride4);
        const float v0 = (root.x != 0.f) ? tab_sqrt<true>(c[0]) : c[0], v1 = (root.y != 0.f) ? tab_sqrt<true>(c[1]) : c[1];
        const float v2 = (root.z != 0.f) ? tab_sqrt<true>(c[2]) : c[2], v3 = (root.w != 0.f) ? tab_sqrt<true>(c[3]) : c[3];
        // convert_int_sat_rtn: cvt.rmi.s32.f32 (NaN -> 0, saturating), then Axis::GetIndexCode's clamp to the under- / overflow bins
        const int k0 = min(max(__float2int_rd(fmaf(scale.x, v0, shift.x)), -1), bins.x) + 1;
        const int k1 = min(max(__float2int_rd(fmaf(scale.y, v1, shift.y)), -1), bins.y) + 1;
        const int k2 = min(max(__float2int_rd(fmaf(scale.z, v2, shift.z)), -1), bins.z) + 1;
        const int k3 = min(max(__float2int_rd(fmaf(scale.w, v3, shift.w)), -1), bins.w) + 1;
        return stride.x * static_cast<uint32_t>(k0) + stride.y * static_cast<uint32_t>(k1) + stride.z * static_cast<uint32_t>(k2) +
               stride.w * static_cast<uint32_t>(k3);
    }
    uint32_t index = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const DevAxis &ax = tb.axes[i];
        float v = c[i];
        if (FAST) {
            const int kind = ax.inverse;
            if (kind >= 3) v = (kind == 3) ? cbrtf(v) : powf(v, ax.inv_power);
            else v = (kind == 2) ? tab_sqrt<true>(v) : ((kind == 1) ? 1.f : v);
        } else {
            if (ax.inverse == 1) v = 1.f;
            else if (ax.inverse == 2) v = sqrtf(v);
            else if (ax.inverse == 3) v = cbrtf(v);
            else if (ax.inverse == 4) v = powf(v, ax.inv_power);
        }
        const float f = floorf(ax.scale * v - ax.offset);
        // convert_int_sat_rtn: NaN -> 0, saturation (cvt.rmi.s32.f32 does exactly that)
        int k = FAST ? __float2int_rd(ax.scale * v - ax.offset)
                     : ((f != f) ? 0 : ((f >= 2147483648.f) ? 2147483647 : ((f <= -2147483648.f) ? (-2147483647 - 1) : static_cast<int>(f))));
        k = min(max(k, -1), ax.n_bins) + 1;
        index += ax.stride * static_cast<uint32_t>(k);
    }
    return index;
}

// isOutOfBounds (Axes.cxx:113-123, 151-159)
__device__ inline bool table_out_of_bounds(const TabulateArgs &tb, const float c[5])
{
    return (tb.geometry == 0) ? ((c[3] > tb.max3) || (c[0] > tb.max0)) : (c[3] > tb.max3);
}

// getBinIndex (Axes.cxx:71-93) with Axis::GetIndexCode (Axis.cxx:44-60): convert_int_sat_rtn = floor with saturation
__device__ inline uint32_t table_bin_index(const TabulateArgs &tb, const float c[5])
{
    uint32_t index = 0;
    for (int i = 0; i < tb.ndim; ++i) {
        const DevAxis &ax = tb.axes[i];
        float v = c[i];
        if (ax.inverse == 1) v = 1.f;
        else if (ax.inverse == 2) v = sqrtf(v);
        else if (ax.inverse == 3) v = cbrtf(v);
        else if (ax.inverse == 4) v = powf(v, ax.inv_power);
        const float f = floorf(ax.scale * v - ax.offset);
        int k = (f != f) ? 0 : ((f >= 2147483648.f) ? 2147483647 : ((f <= -2147483648.f) ? (-2147483647 - 1) : static_cast<int>(f)));
        k = min(max(k, -1), ax.n_bins) + 1;
        index += ax.stride * static_cast<uint32_t>(k);
    }
    return index;
}

__device__ inline float table_angular_acceptance(const TabulateArgs &tb, float x)
{
    if (tb.num_angular == 0) return 0.f;
    float v = tb.angular[tb.num_angular - 1];
    for (int i = tb.num_angular - 2; i >= 0; --i) v = tb.angular[i] + x * v;
    return v;
}

} // namespace clsimcu
