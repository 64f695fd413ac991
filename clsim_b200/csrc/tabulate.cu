// tabulate.cu -- the table-maker variant: host side.
//
// Replaces I3CLSimStepToTableConverter (private/clsim/tabulator/I3CLSimStepToTableConverter.cxx) for the step ->
// table path.  The kernel is the reference-order kernel (kernel_reference.cu) with a TabulateArgs block: the
// reference compiles the same propKernel with -DTABULATE.  What is different on the B200: the table (up to a few
// hundred MB: 202 x 38 x 102 x 107 bins in the default spherical layout) stays in HBM and the kernel adds to it
// with float atomics, so nothing comes back per bunch.  The reference ships 8-byte (index, weight) entries to the
// host for every metre of every photon's path and sums them there (…cxx:484-497); its restart protocol for work
// items that run out of entry space disappears with the entry buffers.
//
// The table geometry (Axes / Axis: bin index code, bin edges, bin volumes) is restated from
// private/clsim/tabulator/Axes.cxx and Axis.cxx, the header values from …StepToTableConverter.cxx:96-121, 141-152.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/clsimcuda.h"
#include "mcpe.h"
#include "tables.h"
#include "tabulate.h"

namespace clsimcu {
namespace {

constexpr double kDegree = 3.14159265358979323846 / 180.0;   // I3Units::degree
constexpr double kSpeedOfLight = 0.299792458;                  // I3Constants::c [m/ns]

#define CUDA_OK(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// Axis::Transform / InverseTransform (Axis.cxx:89-133)
double axis_transform(const clsimcu_axis &a, double v) { return a.kind == CLSIMCU_AXIS_POWER ? std::pow(v, static_cast<double>(a.power)) : v; }
double axis_inverse(const clsimcu_axis &a, double v) { return a.kind == CLSIMCU_AXIS_POWER ? std::pow(v, 1. / a.power) : v; }
// Axis::GetBinEdge (Axis.cxx:72-78)
double axis_edge(const clsimcu_axis &a, unsigned i)
{
    const double imin = axis_inverse(a, a.min), imax = axis_inverse(a, a.max);
    const double istep = (imax - imin) / a.n_bins;
    return axis_transform(a, imin + i * istep);
}

double poly5(const double c[5], double x) { return c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4]))); }

// I3CLSimLightSourceToStepConverterUtils::NumberOfPhotonsPerMeter (…Utils.cxx:44-110): the reference integrates with
// GSL QAG to 1e-5; the bias is piecewise linear in wavelength, so Gauss-Legendre node to node is exact to rounding
double photons_per_meter(const double n_phase[5], const clsimcu_wlen_bias *bias, double from_wlen, double to_wlen)
{
    static const double gx[5] = {0., 0.5384693101056831, -0.5384693101056831, 0.9061798459386640, -0.9061798459386640};
    static const double gw[5] = {0.5688888888888889, 0.4786286704993665, 0.4786286704993665, 0.2369268850561891, 0.2369268850561891};
    auto bias_at = [&](double wlen) {
        if (!bias || bias->kind == CLSIMCU_BIAS_CONSTANT) return bias ? bias->value : 1.;
        double whole;
        double frac = std::modf((wlen - bias->x0) / bias->dx, &whole);
        long bin = static_cast<long>(whole);
        if (bin < 0 || (bin == 0 && frac < 0)) { bin = 0; frac = 0; }
        else if (bin >= bias->n - 1) { bin = bias->n - 2; frac = 1; }
        return bias->v[bin] + (bias->v[bin + 1] - bias->v[bin]) * frac;
    };
    auto f = [&](double inv_wlen) {
        const double wlen = 1. / inv_wlen;
        const double n = poly5(n_phase, wlen / 1e-6);
        return bias_at(wlen) * (2. * 3.14159265358979323846 / 137.) * (1. - 1. / (n * n));
    };
    std::vector<double> edges = {1. / to_wlen, 1. / from_wlen};
    if (bias && bias->kind == CLSIMCU_BIAS_TABLE)
        for (int i = 0; i < bias->n; ++i) {
            const double w = bias->x0 + bias->dx * i;
            if (w > from_wlen && w < to_wlen) edges.push_back(1. / w);
        }
    std::sort(edges.begin(), edges.end());
    double total = 0.;
    for (size_t k = 0; k + 1 < edges.size(); ++k) {
        // split every interval further: the integrand is smooth but not polynomial
        const int parts = 8;
        for (int p = 0; p < parts; ++p) {
            const double lo = edges[k] + (edges[k + 1] - edges[k]) * p / parts, hi = edges[k] + (edges[k + 1] - edges[k]) * (p + 1) / parts;
            const double mid = 0.5 * (lo + hi), half = 0.5 * (hi - lo);
            for (int q = 0; q < 5; ++q) total += gw[q] * f(mid + half * gx[q]) * half;
        }
    }
    return total;
}

} // namespace
} // namespace clsimcu

using namespace clsimcu;

struct clsimcu_tabulator {
    clsimcu_engine *engine = nullptr;
    clsimcu_tabulator_config cfg{};
    std::vector<double> angular;
    size_t shape[5] = {0}, strides[5] = {0}, n_bins = 0;
    TabulateArgs args{};          // host copy; reference particle changes per bunch
    TabulateArgs *d_args[4] = {nullptr, nullptr, nullptr, nullptr};   // ring: a launch reads its own copy
    unsigned next_args = 0;
    float *d_table = nullptr, *d_squared = nullptr;
    double spectral_bias_factor = 1., n_group = 0., n_phase = 0.;
    double sum_of_photon_weights = 0.;
    uint64_t num_photons = 0;
    std::mutex mutex;
};

namespace {

void free_tabulator(clsimcu_tabulator *t)
{
    if (!t) return;
    if (t->engine) {
        cudaSetDevice(engine_device(t->engine));
        cudaDeviceSynchronize();
    }
    for (TabulateArgs *p : t->d_args) cudaFree(p);
    cudaFree(t->d_table);
    cudaFree(t->d_squared);
    if (t->engine) clsimcu_destroy(t->engine);
    delete t;
}

// SphericalAxes / CylindricalAxes::GetBinVolume (Axes.cxx:125-140, 161-172)
double bin_volume(const clsimcu_tabulator_config &c, const size_t idx[5])
{
    const clsimcu_axis *a = c.axes;
    if (c.geometry == CLSIMCU_TABLE_SPHERICAL) {
        const double scalefactor = (a[1].max > 180.) ? 1 : 2;
        return ((std::pow(axis_edge(a[0], idx[0] + 1), 3) - std::pow(axis_edge(a[0], idx[0]), 3)) / 3.) * scalefactor * kDegree *
               (axis_edge(a[1], idx[1] + 1) - axis_edge(a[1], idx[1])) * (axis_edge(a[2], idx[2] + 1) - axis_edge(a[2], idx[2]));
    }
    return ((std::pow(axis_edge(a[0], idx[0] + 1), 2) - std::pow(axis_edge(a[0], idx[0]), 2)) / 2.) * 2 *
           (axis_edge(a[1], idx[1] + 1) - axis_edge(a[1], idx[1])) * (axis_edge(a[2], idx[2] + 1) - axis_edge(a[2], idx[2]));
}

} // namespace

extern "C" {

int clsimcu_tabulator_create(const clsimcu_config *scene, const clsimcu_tabulator_config *cfg, clsimcu_tabulator **out)
{
    if (!scene || !cfg || !out) return report_error(CLSIMCU_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->struct_size != static_cast<int32_t>(sizeof(clsimcu_tabulator_config)))
        return report_error(CLSIMCU_ERR_INVALID, "clsimcu_tabulator_config.struct_size does not match this library");
    if (cfg->geometry != CLSIMCU_TABLE_SPHERICAL && cfg->geometry != CLSIMCU_TABLE_CYLINDRICAL) return report_error(CLSIMCU_ERR_INVALID, "unknown table geometry");
    if (cfg->num_axes != 4 && cfg->num_axes != 5) return report_error(CLSIMCU_ERR_INVALID, "a table has 4 axes, or 5 with the impact angle");
    if (!(cfg->step_length > 0.) || !(cfg->reference_area > 0.)) return report_error(CLSIMCU_ERR_INVALID, "step length and reference area must be positive");
    if (cfg->num_axes == 4 && (cfg->num_angular_coefficients < 1 || !cfg->angular_coefficients || cfg->num_angular_coefficients > 24))
        return report_error(CLSIMCU_ERR_INVALID, "a 4-axis table needs the angular acceptance polynomial (1..24 coefficients)");
    for (int i = 0; i < cfg->num_axes; ++i) {
        const clsimcu_axis &a = cfg->axes[i];
        if (a.kind != CLSIMCU_AXIS_LINEAR && a.kind != CLSIMCU_AXIS_POWER) return report_error(CLSIMCU_ERR_INVALID, "unknown axis kind");
        if (a.n_bins == 0 || !(a.max > a.min)) return report_error(CLSIMCU_ERR_INVALID, "an axis needs bins and max > min");
        if (a.kind == CLSIMCU_AXIS_POWER && a.power >= 1 && a.min < 0.) return report_error(CLSIMCU_ERR_INVALID, "a power axis cannot start below zero");
    }
    clsimcu_tabulator *t = new clsimcu_tabulator;
    try {
        t->cfg = *cfg;
        if (cfg->num_axes == 4) t->angular.assign(cfg->angular_coefficients, cfg->angular_coefficients + cfg->num_angular_coefficients);
        t->cfg.angular_coefficients = nullptr;
        // Axes::Axes (Axes.cxx:51-64): every axis has an under- and an overflow bin, last axis is contiguous
        const int nd = cfg->num_axes;
        t->shape[nd - 1] = cfg->axes[nd - 1].n_bins + 2;
        t->strides[nd - 1] = 1;
        for (int i = nd - 2; i >= 0; --i) {
            t->shape[i] = cfg->axes[i].n_bins + 2;
            t->strides[i] = t->strides[i + 1] * t->shape[i + 1];
        }
        t->n_bins = t->strides[0] * t->shape[0];
        if (t->n_bins > 0xffffffffull) throw std::invalid_argument("the table has more than 2^32 bins");

        // the engine behind it: the reference's preamble (…cxx:177-183) as options
        clsimcu_config sc = *scene;
        // four-axis tables run on the persistent kernel unless the caller asks for the reference-order twin; the fifth axis
        // (impact angle) draws two numbers per entry from the work item's stream and stays with the twin
        sc.kernel_mode = (nd == 4 && scene->kernel_mode == CLSIMCU_KERNEL_FAST) ? CLSIMCU_KERNEL_FAST : CLSIMCU_KERNEL_REFERENCE;
        if (sc.kernel_mode == CLSIMCU_KERNEL_FAST && sc.rng_a == nullptr) sc.rng_n = 0;   // the engine sizes and seeds its own streams
        sc.enable_double_buffering = 0;
        sc.stop_detected_photons = 0;
        sc.save_all_photons = 1;
        sc.save_all_photons_prescale = 1.;
        sc.photon_history_entries = 0;
        sc.fixed_number_of_absorption_lengths = 42.;
        sc.output_photons_per_workitem = 1;
        std::memset(&sc.geometry, 0, sizeof sc.geometry);
        if (sc.rng_n == 0 && sc.kernel_mode == CLSIMCU_KERNEL_REFERENCE) sc.rng_n = sc.max_num_workitems;
        if (clsimcu_create(&sc, &t->engine) != CLSIMCU_OK) {
            const std::string msg = clsimcu_last_error();
            free_tabulator(t);
            return report_error(CLSIMCU_ERR_CUDA, msg);
        }

        // GetMinimumRefractiveIndex (…cxx:96-121), including its sampling of wmin + i * (wmax - wmin) for i < 1000
        {
            double best_group = std::numeric_limits<double>::infinity(), best_phase = std::numeric_limits<double>::infinity();
            // medium range (python/MakeIceCubeMediumProperties.py:178-179 by default); the index functions are unbounded
            const double wmin = cfg->min_wavelength > 0. ? cfg->min_wavelength : 265e-9, wmax = cfg->max_wavelength > 0. ? cfg->max_wavelength : 675e-9;
            for (unsigned i = 0; i < 1000; ++i) {
                const double x = (wmin + i * (wmax - wmin)) / 1e-6;
                const double np = poly5(scene->medium.n_phase, x);
                const double ng = np * poly5(scene->medium.n_group, x);
                if (ng > 1 && ng < best_group) {
                    best_group = ng;
                    best_phase = np;
                }
            }
            t->n_group = best_group;
            t->n_phase = best_phase;
        }
        // spectralBiasFactor_ (…cxx:141-152)
        t->spectral_bias_factor = photons_per_meter(scene->medium.n_phase, nullptr, 300e-9, 600e-9) /
                                  photons_per_meter(scene->medium.n_phase, &scene->wlen_bias, cfg->min_wavelength > 0. ? cfg->min_wavelength : 265e-9,
                                                    cfg->max_wavelength > 0. ? cfg->max_wavelength : 675e-9);

        TabulateArgs &a = t->args;
        std::memset(&a, 0, sizeof a);
        a.geometry = cfg->geometry;
        a.ndim = nd;
        a.full_azimuth = (cfg->geometry == CLSIMCU_TABLE_SPHERICAL && cfg->axes[1].max > 180.) ? 1 : 0;
        for (int i = 0; i < nd; ++i) {
            const clsimcu_axis &ax = cfg->axes[i];
            // Axis::GetIndexCode (Axis.cxx:44-60): both numbers pass through ToFloatString
            const double scale = ax.n_bins / (axis_inverse(ax, ax.max) - axis_inverse(ax, ax.min));
            const double offset = scale * axis_inverse(ax, ax.min);
            a.axes[i].scale = float_literal(scale);
            a.axes[i].offset = float_literal(offset);
            a.axes[i].n_bins = static_cast<int>(ax.n_bins);
            a.axes[i].stride = static_cast<uint32_t>(t->strides[i]);
            a.axes[i].inverse = 0;
            if (ax.kind == CLSIMCU_AXIS_POWER) {
                // PowerAxis::GetInverseTransformCode (Axis.cxx:149-171)
                if (ax.power == 0) a.axes[i].inverse = 1;
                else if (ax.power == 1) a.axes[i].inverse = 0;
                else if (ax.power == 2) a.axes[i].inverse = 2;
                else if (ax.power == 3) a.axes[i].inverse = 3;
                else { a.axes[i].inverse = 4; a.axes[i].inv_power = float_literal(1. / ax.power); }
            }
        }
        a.simple4 = (nd == 4) ? 1 : 0;
        for (int i = 0; i < 4 && i < nd; ++i) {
            a.scale4[i] = a.axes[i].scale;
            a.neg_offset4[i] = -a.axes[i].offset;
            a.n_bins4[i] = a.axes[i].n_bins;
            a.stride4[i] = a.axes[i].stride;
            a.root4[i] = (a.axes[i].inverse == 2) ? 1.f : 0.f;
            if (a.axes[i].inverse != 0 && a.axes[i].inverse != 2) a.simple4 = 0;
        }
        a.max0 = float_literal(cfg->axes[0].max);
        a.max3 = float_literal(cfg->axes[3].max);
        a.step_length = float_literal(cfg->step_length);
        a.min_inv_group_vel = float_literal(t->n_group / kSpeedOfLight);
        a.tan_theta_c = float_literal(std::sqrt(t->n_phase * t->n_phase - 1.));
        a.num_angular = static_cast<int>(t->angular.size());
        for (size_t i = 0; i < t->angular.size(); ++i) a.angular[i] = float_literal(t->angular[i]);

        CUDA_OK(cudaSetDevice(engine_device(t->engine)));
        CUDA_OK(cudaMalloc(&t->d_table, t->n_bins * sizeof(float)));
        CUDA_OK(cudaMemset(t->d_table, 0, t->n_bins * sizeof(float)));
        if (cfg->store_squared_weights) {
            CUDA_OK(cudaMalloc(&t->d_squared, t->n_bins * sizeof(float)));
            CUDA_OK(cudaMemset(t->d_squared, 0, t->n_bins * sizeof(float)));
        }
        a.table = t->d_table;
        a.squared = t->d_squared;
        for (TabulateArgs *&p : t->d_args) CUDA_OK(cudaMalloc(&p, sizeof(TabulateArgs)));
        CUDA_OK(cudaDeviceSynchronize());   // (the tables are zeroed on the default stream, the engine launches on a non-blocking one)
    } catch (const std::invalid_argument &ex) {
        const std::string msg = ex.what();
        free_tabulator(t);
        return report_error(CLSIMCU_ERR_INVALID, msg);
    } catch (const std::exception &ex) {
        const std::string msg = ex.what();
        free_tabulator(t);
        return report_error(CLSIMCU_ERR_CUDA, msg);
    }
    *out = t;
    return CLSIMCU_OK;
}

int clsimcu_tabulator_destroy(clsimcu_tabulator *t)
{
    if (!t) return report_error(CLSIMCU_ERR_INVALID, "tabulator is NULL");
    free_tabulator(t);
    return CLSIMCU_OK;
}

int clsimcu_tabulator_enqueue(clsimcu_tabulator *t, const clsimcu_step *steps, size_t n, const clsimcu_reference_particle *ref)
{
    if (!t) return report_error(CLSIMCU_ERR_STATE, "I3CLSimStepToTableConverter is not initialized!");
    if (!steps || n == 0) return CLSIMCU_OK;   // EnqueueSteps ignores an empty series (…cxx:293-294)
    if (!ref) return report_error(CLSIMCU_ERR_INVALID, "reference particle is NULL");
    if (n > engine_max_items(t->engine)) return report_error(CLSIMCU_ERR_INVALID, "Number of steps is greater than maximum number of work items!");
    std::lock_guard<std::mutex> lk(t->mutex);
    // I3CLSimReferenceParticle (…cxx:65-93)
    TabulateArgs a = t->args;
    a.ref_pos[0] = static_cast<float>(ref->x); a.ref_pos[1] = static_cast<float>(ref->y); a.ref_pos[2] = static_cast<float>(ref->z);
    a.ref_pos[3] = static_cast<float>(ref->t);
    a.ref_dir[0] = static_cast<float>(ref->dir_x); a.ref_dir[1] = static_cast<float>(ref->dir_y); a.ref_dir[2] = static_cast<float>(ref->dir_z);
    a.ref_dir[3] = 0.f;
    const double perpz = std::hypot(ref->dir_x, ref->dir_y);
    double px = 1., py = 0., pz = 0.;
    if (perpz > 0.) {
        px = -ref->dir_x * ref->dir_z / perpz;
        py = -ref->dir_y * ref->dir_z / perpz;
        pz = perpz;
        const double norm = std::sqrt(px * px + py * py + pz * pz);   // I3Direction normalises
        px /= norm; py /= norm; pz /= norm;
    }
    a.ref_perp[0] = static_cast<float>(px); a.ref_perp[1] = static_cast<float>(py); a.ref_perp[2] = static_cast<float>(pz); a.ref_perp[3] = 0.f;
    // a launch keeps reading its argument block: with four blocks, wait for the stream when the ring wraps
    TabulateArgs *slot = t->d_args[t->next_args % 4];
    std::string err = engine_copy_on_stream(t->engine, slot, &a, sizeof a, true, (t->next_args % 4) == 3);
    ++t->next_args;
    if (err.empty()) err = engine_launch_tabulate(t->engine, steps, n, slot);
    if (!err.empty()) return report_error(CLSIMCU_ERR_CUDA, err);
    for (size_t i = 0; i < n; ++i) {
        t->num_photons += steps[i].num_photons;
        t->sum_of_photon_weights += static_cast<double>(steps[i].num_photons) * steps[i].weight;
    }
    return CLSIMCU_OK;
}

int clsimcu_tabulator_finish(clsimcu_tabulator *t)
{
    if (!t) return report_error(CLSIMCU_ERR_STATE, "I3CLSimStepToTableConverter is not initialized!");
    const std::string err = engine_copy_on_stream(t->engine, nullptr, nullptr, 0, false, true);
    if (!err.empty()) return report_error(CLSIMCU_ERR_CUDA, err);
    return CLSIMCU_OK;
}

int clsimcu_tabulator_info(clsimcu_tabulator *t, double out[16])
{
    if (!t || !out) return report_error(CLSIMCU_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(t->mutex);
    for (int i = 0; i < 16; ++i) out[i] = 0.;
    out[0] = static_cast<double>(t->n_bins);
    for (int i = 0; i < t->cfg.num_axes; ++i) {
        out[1 + i] = static_cast<double>(t->shape[i]);
        out[6 + i] = static_cast<double>(t->strides[i]);
    }
    out[11] = t->spectral_bias_factor * t->sum_of_photon_weights;   // tableHeader["n_photons"], …cxx:609
    out[12] = t->n_group;
    out[13] = t->n_phase;
    out[14] = t->spectral_bias_factor;
    out[15] = static_cast<double>(t->num_photons);
    return CLSIMCU_OK;
}

int clsimcu_tabulator_get_table(clsimcu_tabulator *t, int normalize, float *bins, float *squared, size_t cap)
{
    if (!t) return report_error(CLSIMCU_ERR_STATE, "I3CLSimStepToTableConverter is not initialized!");
    if (!bins || cap < t->n_bins) return report_error(CLSIMCU_ERR_INVALID, "output buffer is smaller than the table");
    if (squared && !t->d_squared) return report_error(CLSIMCU_ERR_INVALID, "squared weights were not stored");
    std::lock_guard<std::mutex> lk(t->mutex);
    std::string err = engine_copy_on_stream(t->engine, bins, t->d_table, t->n_bins * sizeof(float), false, squared == nullptr);
    if (err.empty() && squared) err = engine_copy_on_stream(t->engine, squared, t->d_squared, t->n_bins * sizeof(float), false, true);
    if (!err.empty()) return report_error(CLSIMCU_ERR_CUDA, err);
    if (normalize) {
        // Normalize() (…cxx:522-556): the first three dimensions are spatial
        const int nd = t->cfg.num_axes;
        const size_t spatial_stride = t->strides[2];
        size_t idx[5];
        for (size_t offset = 0; offset < t->n_bins; offset += spatial_stride) {
            for (int j = 0; j < nd; ++j) {
                const long raw = static_cast<long>(offset / t->strides[j] % t->shape[j]) - 1;
                idx[j] = static_cast<size_t>(std::min(std::max(raw, 0l), static_cast<long>(t->shape[j]) - 3));
            }
            double norm = bin_volume(t->cfg, idx) / (t->cfg.step_length * t->cfg.reference_area);
            for (size_t i = 0; i < spatial_stride; ++i) bins[offset + i] = static_cast<float>(bins[offset + i] / norm);
            if (squared) {
                norm *= norm;
                for (size_t i = 0; i < spatial_stride; ++i) squared[offset + i] = static_cast<float>(squared[offset + i] / norm);
            }
        }
    }
    return CLSIMCU_OK;
}

} // extern "C"
