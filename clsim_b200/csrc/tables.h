// tables.h -- host-side flattening of the model description into POD tables.
//
// This is the B200 engine's answer to the reference's run-time code generators
// (private/opencl/I3CLSimHelperGenerateGeometrySource.cxx,
//  private/opencl/I3CLSimHelperGenerateMediumPropertiesSource{,_Optimizers}.cxx and the
//  GetOpenCLFunction() members of the function / random_value classes): instead of printing
// OpenCL C text that a JIT turns into constants, the same numbers are computed once and
// laid out as flat arrays that are uploaded to HBM and staged into shared memory by the
// kernels.  Values are rounded exactly the way the reference's literals are
// (double -> "%.10e" text -> float), see float_literal().
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/clsimcuda.h"

namespace clsimcu {

// double -> the float an OpenCL compiler reads from the reference's printed literal
// (private/clsim/I3CLSimHelperToFloatString.h:37-60)
float float_literal(double value);

struct WlenGeneratorTable {
    int kind = 0;
    int n = 0;
    float x0 = 0.f, dx = 0.f;        // equal spacing
    float min_val = 0.f, range = 0.f; // no dispersion
    float value = 0.f;               // constant
    std::vector<float> xs, density, cumulative;
};

struct CellGridTable {
    int num_x = 1, num_y = 1;
    float start_x = 0.f, start_y = 0.f, width_x = 0.f, width_y = 0.f;
    std::vector<uint16_t> cell_to_string; // [y*num_x + x], 0xFFFF = empty
};

struct GeometryTables {
    int num_strings = 0;
    float om_radius = 0.f;
    float string_max_radius = 0.f;
    std::vector<float> string_x, string_y, string_min_z, string_max_z, string_radius;
    std::vector<uint8_t> string_set;
    int num_sets = 0, max_layers = 0;
    std::vector<uint16_t> set_layer_count;
    std::vector<float> set_start_z, set_layer_height;
    std::vector<uint16_t> layer_to_dom; // [set*max_layers + layer], padded to a multiple of 64
    std::vector<CellGridTable> grids;   // one per subdetector
    int max_doms_per_string = 0;
    float tmpl_scale_x = 0.f, tmpl_scale_y = 0.f;
    std::vector<int16_t> tmpl_dx, tmpl_dy;
    std::vector<float> tmpl_z;
    std::vector<uint32_t> string_tmpl_start;
    std::vector<float> string_mean_x, string_mean_y;
    std::vector<int32_t> string_index_to_id;
    std::vector<std::vector<uint32_t>> dom_index_to_id;
};

struct MediumTables {
    int num_layers = 0;
    float z0 = 0.f, h = 0.f;
    float kappa = 0.f, A = 0.f, B = 0.f, D = 0.f, E = 0.f, alpha = 0.f;
    float inv_ref_wlen = 0.f;  // literal of 1/(400 nm)
    std::vector<float> a_dust400, delta_tau, b400;
    float n_phase[5] = {0, 0, 0, 0, 0}, n_group[5] = {0, 0, 0, 0, 0};
    float c_light = 0.f;
    int scat_kind = 0;
    float f_sl = 0.f, one_minus_f_sl = 0.f, g = 0.f, g2 = 0.f, sl_beta = 0.f;
    int tilt_nd = 0, tilt_nz = 0;
    std::vector<float> tilt_dist, tilt_corr;
    float tilt_z0 = 0.f, tilt_dz = 0.f, tilt_lnx = 0.f, tilt_lny = 0.f;
    bool anisotropy = false;
    float l[3] = {0, 0, 0}, rl[3] = {0, 0, 0}, azx = 0.f, azy = 0.f, neg_azy = 0.f, B2 = 0.f;
    float pre[9] = {0}, post[9] = {0};
    bool pre_renorm = false, post_renorm = false;
};

struct BiasTable {
    int kind = 0, n = 0;
    float x0 = 0.f, dx = 0.f, value = 1.f;
    std::vector<float> v;
};

struct SceneTables {
    MediumTables medium;
    std::vector<WlenGeneratorTable> generators;
    BiasTable bias;
    GeometryTables geometry;
    bool has_geometry = false;
    bool stop_detected = false, save_all = false, fixed_abs = false, pancake = false;
    float prescale = 0.f, fixed_abs_lens = 0.f, pancake_factor = 1.f;
    int history_entries = 0;
};

// Throws std::runtime_error (messages follow the reference's where one exists).
void build_scene_tables(const clsimcu_config &config, SceneTables &out);

// JSON text, same schema as the oracle's, for table parity tests.
std::string describe_scene_tables(const SceneTables &tables);

// xy pixel map of the fast kernel's collision pre-test (device_scene.h, DevGeometry::near_info): per pixel the byte
// offset of the nearest string's 16-byte record (low 16 bits) and the pixel's RANGE as the upper 16 bits of an fp32,
// rounded down.  Host code, no GPU needed: the engine uploads it, the CPU tests check its guarantees.
struct CollisionMap {
    int nx = 0, ny = 0;
    float x0 = 0.f, y0 = 0.f, pixel = 0.f, inv_pixel = 0.f, off_x = 0.f, off_y = 0.f;
    std::vector<uint32_t> info;
};
CollisionMap build_collision_map(const GeometryTables &g, int pixel_budget);
std::string describe_collision_map(const GeometryTables &g, const CollisionMap &m);

// Safe-prime MWC multipliers, rows [first, first+n) of the descending sequence from
// 4294967118 (private/make_safeprimes/main.cxx:32-104).  Multi-threaded; memoised on disk in
// the reference's binary "safeprimes_base32" format when cache_path is non-empty
// (private/opencl/mwcrng_init.h:67-97 reads the same format).
void safeprime_multipliers(uint64_t first, uint64_t n, uint32_t *out, const std::string &cache_path);

// x[] seeds from `seed` under the rejection rule of mwcrng_init.h:107-113.
void seed_rng_states(uint64_t seed, const uint32_t *a, uint64_t *x, size_t n);

} // namespace clsimcu
