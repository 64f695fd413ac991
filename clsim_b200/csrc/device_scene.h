// device_scene.h -- the POD block the kernels read: scalars by value (kernel parameter,
// __grid_constant__) and pointers into one table arena in HBM.  Built by the engine from
// SceneTables (tables.h).  Everything the reference bakes into JIT-compiled constants
//  lives here instead.
#pragma once

#include <cstdint>

namespace clsimcu {

constexpr int kTiltLutMaxCells = 64;   // fast kernel: cells of the interval grid of the ice-tilt table

constexpr int kMaxWlenGenerators = 8;
constexpr int kMaxSubdetectors = 9; // sparse_collision_kernel.c.cl:455-457
constexpr uint32_t kFastKernelStepIndexBits = 27; // the fast kernel tags photons with (step index | creating lane << 27)

struct DevWlenGenerator {
    int kind, n;
    float x0, dx, min_val, range, value;
    const float *xs, *density, *cumulative;
};

struct DevCellGrid {
    int num_x, num_y;
    float start_x, start_y, width_x, width_y;
    float inv_width_x, inv_width_y; // fast kernel only
    const uint16_t *cell_to_string;
};

struct DevMedium {
    int num_layers;
    float z0, h, inv_h;
    float kappa, A, B, D, E, alpha, inv_ref_wlen;
    float n_phase[5], n_group[5], c_light;
    int scat_kind;
    float f_sl, one_minus_f_sl, g, g2, sl_beta;
    float inv_f_sl, inv_one_minus_f_sl, inv_2g; // fast kernel: reciprocals (0 where undefined)
    // fast kernel, SL + HG mix folded into constants (computed in double on the host, see upload_tables):
    // with U = the 32-bit draw as a float (u = U 2^-32):
    //   cos_sl = ex2(sl_beta * lg2(U) + sl_off) - 1            == 2 (u / f_sl)^beta - 1
    //   cos_hg = hg_c - hg_w * r^2,  r = 1 / (hg_h0 + hg_h1 U) == ((1 + g^2) - ((1 - g^2) / (1 + g s))^2) / 2g,  s = 2 (1 - u) / (1 - f_sl) - 1
    //   SL where U < mix_split == u < f_sl (both sides scaled by the exact 2^32)
    float sl_off, hg_h0, hg_h1, hg_c, hg_w, mix_split;
    int mix_folded; // the five constants above are set (f_sl in (0,1), g != 0)
    int tilt_nd, tilt_nz;
    float tilt_z0, tilt_dz, tilt_inv_dz, tilt_lnx, tilt_lny;
    float tilt_zr_offset;                  // fast kernel: -tilt_z0 / tilt_dz, so that the z node index is one FMA
    // fast kernel: the interval of the tilt table along the tilt direction from a uniform grid (cell = scale * nr + offset),
    // at most one interior node per cell; tilt_lut_n == 0: no grid (nodes too close together), compare against the nodes
    float tilt_lut_scale, tilt_lut_offset;
    int tilt_lut_n;
    int anisotropy, pre_renorm, post_renorm;
    float l[3], rl[3], azx, azy, neg_azy, B2;
    float pre[9], post[9];
    const float *a_dust400, *delta_tau, *b400; // [num_layers]
    const float *abs_dust, *abs_tau;           // fast kernel: D*aDust400+E, 1+0.01*deltaTau
    const float *tilt_dist, *tilt_corr;
};

struct DevBias {
    int kind, n;
    float x0, dx, value;
    const float *v;
};

struct DevGeometry {
    int num_strings, num_sets, max_layers, num_grids, layer_table_size;
    float om_radius, string_max_radius;
    float tmpl_scale_x, tmpl_scale_y;
    DevCellGrid grids[kMaxSubdetectors];
    const float *string_x, *string_y, *string_min_z, *string_max_z;
    const uint8_t *string_set;
    const uint16_t *set_layer_count;
    const float *set_start_z, *set_layer_height;
    const uint16_t *layer_to_dom;
    const int16_t *tmpl_dx, *tmpl_dy;
    const float *tmpl_z;
    const uint32_t *string_tmpl_start;
    const float *string_mean_x, *string_mean_y;
    // Fast kernel only: pixel map over the xy plane.  near_info[pixel] = 16 x the index of the string
    // nearest to the pixel centre (low 16 bits: at most 4094 strings) | the RANGE of the pixel, stored as the upper 16
    // bits of an fp32 (rounded down): a photon anywhere in the pixel can fly that far before any
    // OTHER string can come within string_max_radius of it.  The fast kernel cuts flights at that
    // range, so a segment only ever has to be tested against the one named string (exactly, from
    // the photon's own position).  Range +inf with string index num_strings marks pixels where the strings
    // are too dense for the pixel size: there every segment takes the reference's cell walk.  Points outside the map
    // clamp to the border pixels (the bound stays valid: projection onto the map rectangle is
    // non-expansive and every string lies inside it).
    int near_nx, near_ny;
    float near_x0, near_y0, near_inv_pixel;
    float near_off_x, near_off_y; // -near_x0 * near_inv_pixel, -near_y0 * near_inv_pixel
    const uint32_t *near_info;
    // index -> ID rewrite on the device (…ConverterOpenCL.cxx:1565-1602 does it on the host)
    const int16_t *string_index_to_id;
    const uint32_t *dom_id_offset; // per string into dom_ids
    const uint16_t *dom_ids;
};

struct DevScene {
    DevMedium medium;
    DevWlenGenerator generators[kMaxWlenGenerators];
    int num_generators;
    DevBias bias;
    DevGeometry geo;
    int stop_detected, save_all, fixed_abs, pancake, history_entries;
    float prescale, fixed_abs_lens, pancake_factor, inv_pancake_factor;
    int generic_transforms;   // test hook (environment CLSIMCU_GENERIC_TRANSFORMS, read when the engine is created): no block-matrix form
};

// Per-launch arguments.
// Table-maker variant of the reference-order kernel (-DTABULATE, propagation_kernel.c.cl:226-304): the binning
// code the reference generates from its Axes (private/clsim/tabulator/Axes.cxx:71-93, Axis.cxx:44-60) as data.
struct alignas(16) DevAxis {
    float scale, offset;   // index = clamp(floor(scale * inverse_transform(x) - offset), -1, n_bins) + 1
    int n_bins;
    uint32_t stride;       // (the four numbers every point needs: one 16-byte load)
    int inverse;           // 0: x (linear, power 1)   1: constant 1 (power 0)   2: sqrt   3: cbrt   4: pow(x, inv_power)
    float inv_power;
};
struct TabulateArgs {
    int geometry, ndim, full_azimuth;
    DevAxis axes[5];
    float max0, max3;                    // isOutOfBounds (Axes.cxx:113-123, 151-159)
    float step_length;                   // VOLUME_MODE_STEP
    float min_inv_group_vel, tan_theta_c;
    int num_angular;
    float angular[24];                   // getAngularAcceptance, Horner form of I3CLSimFunctionPolynomial.cxx:139-153
    alignas(16) float ref_pos[4];        // I3CLSimReferenceParticle
    alignas(16) float ref_dir[4];
    alignas(16) float ref_perp[4];
    float *table, *squared;              // HBM, added to atomically
    // persistent kernel, four-axis tables whose axes are all linear or quadratic (`simple4`): the axes again, one array per
    // quantity, so that the four bin indices of a point take five 16-byte loads and no branch
    int simple4;
    alignas(16) float scale4[4];
    alignas(16) float neg_offset4[4];
    alignas(16) int n_bins4[4];
    alignas(16) uint32_t stride4[4];
    alignas(16) float root4[4];          // 1: the inverse transform of this axis is the square root, 0: the identity
};

struct LaunchArgs {
    const void *steps;         // clsimcu_step[num_steps]
    uint32_t num_steps;
    uint32_t max_hits;         // capacity of photons[]
    void *photons;             // clsimcu_photon[max_hits]
    float *history;            // [max_hits][history_entries][4] or nullptr
    float *history_ring;       // fast kernel with photon history: [history_entries][resident threads][4], the lanes' scatter-point rings
    uint32_t *hit_counter;     // device counter (keeps counting past max_hits, quirk 10)
    unsigned long long *stats; // [0] photons created, [1] segments
    uint32_t *work_counter;    // fast kernel: next step to hand out
    uint64_t *rng_x;           // RNG streams used by this launch
    uint32_t *rng_a;
    const struct DevScene *scene_dev; // the scene again, in global memory (out-of-line device functions read it there)
    uint32_t rng_creation_offset;     // fast kernel: thread t creates photons from stream rng_creation_offset + t
    uint64_t *rng_tag_x;       // optional (save-all replay): per record, creation and propagation RNG states
    uint32_t *rng_tag_a;
    int count_stats;
    uint32_t step_chunks;      // fast kernel: parts a step is cut into when work is handed out (set by launch_fast_kernel)
    const TabulateArgs *tabulate;  // reference-order kernel only: table-maker variant (device pointer) or nullptr
};

// kernel launchers (defined in kernel_reference.cu / kernel_fast.cu)
int launch_reference_kernel(const DevScene &scene, const LaunchArgs &args, void *stream);
int launch_fast_kernel(const DevScene &scene, const LaunchArgs &args, int grid_blocks, void *stream);
bool fast_kernel_supports(const DevScene &scene, const char **why);
bool fast_kernel_smem_is_the_problem(const DevScene &scene);
void fast_kernel_geometry(int device, int *grid_blocks, int *threads_per_block);

} // namespace clsimcu
