/*
 * clsim_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, not product).
 *
 * The oracle is a scalar fp32 restatement, for CPUs, of the reference's step->photon
 * path: the static OpenCL files resources/kernels/{mwcrng_kernel,propagation_kernel.c,
 * sparse_collision_kernel.c}.cl plus everything the C++ code generators append
 * (private/opencl/I3CLSimHelperGenerate{Geometry,MediumProperties}Source*.cxx and the
 * GetOpenCLFunction() of the function / random_value classes).  It follows the
 * reference's *precise* math path (useNativeMath=false, the setting the reference
 * uses on CPU devices, python/traysegments/common.py:45).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  Nothing under clsim_b200/ links or imports it.
 *
 * Parity status: component tables are pinned against reference data (rnd.txt safe
 * primes, ice tables through the reference's own Python loaders, the ppc formulas
 * restated in resources/tests/ scripts).  Whole-kernel output is UNPINNED: the reference
 * tree holds no golden propKernel output and neither OpenCL nor IceTray exist in this
 * image.
 *
 * The configuration structs mirror the model-level description the reference objects
 * hold; the layout is declared here independently of include/clsimcuda.h and a test
 * asserts that both agree, so one Python description can drive both.
 */
#ifndef CLSIM_ORACLE_H_INCLUDED
#define CLSIM_ORACLE_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_step {          /* propagation_kernel.h.cl:52-63 */
    float pos_and_time[4];
    float dir_and_length_and_beta[4];
    uint32_t num_photons;
    float weight;
    uint32_t identifier;
    uint8_t source_type;
    uint8_t dummy1;
    uint16_t dummy2;
} oracle_step;

typedef struct oracle_photon {        /* propagation_kernel.h.cl:65-81 */
    float pos_and_time[4];
    float dir[2];
    float wavelength;
    float cherenkov_dist;
    uint32_t num_scatters;
    float weight;
    uint32_t identifier;
    int16_t string_id;
    uint16_t om_id;
    float start_pos_and_time[4];
    float start_dir[2];
    float group_velocity;
    float dist_in_abs_lens;
} oracle_photon;

typedef struct oracle_wlen_generator {
    int32_t kind;         /* 0 interp equal, 1 interp unequal, 2 no dispersion, 3 constant */
    int32_t n;
    double x0, dx;
    const double *x;
    const double *y;
    double from_wlen, to_wlen;
    double value;
} oracle_wlen_generator;

typedef struct oracle_wlen_bias {
    int32_t kind;         /* 0 constant, 1 table */
    int32_t n;
    double x0, dx;
    const double *v;
    double value;
} oracle_wlen_bias;

typedef struct oracle_medium {
    int32_t num_layers;
    int32_t scat_kind;    /* 0 mixed SL+HG, 1 HG, 2 SL */
    double layers_zstart, layers_height;
    double kappa, A, B, D, E;
    const double *a_dust400;
    const double *delta_tau;
    double alpha;
    const double *b400;
    double n_phase[5];
    double n_group[5];
    double f_sl, mean_cos;
    int32_t tilt_num_dist, tilt_num_z;
    const double *tilt_dist;
    const double *tilt_corr;
    double tilt_z0, tilt_dz, tilt_azimuth;
    int32_t has_anisotropy, pre_renormalize, post_renormalize, reserved0;
    double aniso_azimuth, aniso_along, aniso_perp;
    double pre_matrix[9];
    double post_matrix[9];
} oracle_medium;

typedef struct oracle_geometry {
    int32_t num_doms;
    int32_t reserved0;
    const int32_t *string_id;
    const uint32_t *dom_id;
    const double *x, *y, *z;
    const int32_t *subdetector;
    double om_radius;
} oracle_geometry;

typedef struct oracle_config {
    int32_t struct_size;
    int32_t device;               /* ignored */
    int32_t kernel_mode;          /* ignored */
    int32_t enable_double_buffering; /* ignored */
    int32_t stop_detected_photons;
    int32_t save_all_photons;
    int32_t photon_history_entries;
    int32_t num_wlen_generators;
    double save_all_photons_prescale;
    double fixed_number_of_absorption_lengths;
    double pancake_factor;
    uint64_t max_num_workitems;   /* ignored */
    uint32_t workgroup_size;      /* ignored */
    uint32_t output_photons_per_workitem; /* ignored */
    const oracle_wlen_generator *wlen_generators;
    oracle_wlen_bias wlen_bias;
    oracle_medium medium;
    oracle_geometry geometry;
    uint64_t rng_n;               /* ignored: RNG state is passed per call */
    const uint32_t *rng_a;
    const uint64_t *rng_x;
    uint64_t rng_seed;
    uint64_t rng_first_multiplier;
} oracle_config;

typedef struct oracle_scene oracle_scene;

size_t oracle_sizeof_config(void);
const char *oracle_last_error(void);

/* "Compile": build every table the reference's generators would print. */
oracle_scene *oracle_scene_create(const oracle_config *config);
void oracle_scene_destroy(oracle_scene *scene);

/* propKernel over n work-items, work-item i using RNG stream (rng_x[i], rng_a[i])
 * (propagation_kernel.c.cl:406-913).  Hits are emitted in (work-item, emission)
 * order -- the reference's atom_inc order is unspecified.  Photons beyond `cap`
 * are counted but dropped (quirk 10).  string_id/om_id hold the real IDs, as
 * after GetConversionResult.  rng_x is updated in place.
 * history: NULL or cap * photon_history_entries * 4 floats, raw ring layout like
 * the device buffer.  stats[0]=photons created, [1]=segments, [2]=layer crossings,
 * [3]=rng draws.  Returns the number of hits counted (may exceed cap). */
uint64_t oracle_propagate(const oracle_scene *scene, const oracle_step *steps, size_t n,
                          uint64_t *rng_x, const uint32_t *rng_a,
                          oracle_photon *out, size_t cap, float *history,
                          int num_threads, uint64_t stats[4]);

/* Table-maker variant (-DTABULATE, propagation_kernel.c.cl:226-304, 755-785): propKernel over n work-items with
 * savePath instead of the DOMs; every (index, weight) entry is added to bins[] (and weight^2 to squared[], if not
 * NULL) in double precision.  The scene must have been created with save_all_photons and a fixed number of
 * absorption lengths.  reference = {x, y, z, t, dir_x, dir_y, dir_z} of the reference particle.  Returns the
 * number of entries. */
typedef struct oracle_table_config {
    int32_t geometry;            /* 0 spherical, 1 cylindrical */
    int32_t num_axes;            /* 4 or 5 */
    int32_t axis_kind[5];        /* 0 linear, 1 power */
    uint32_t axis_power[5];
    uint32_t axis_bins[5];
    int32_t num_angular_coefficients;
    double axis_min[5], axis_max[5];
    double step_length;
    double n_group, n_phase;     /* minimum refractive indices (…StepToTableConverter.cxx:96-121) */
    const double *angular_coefficients;
} oracle_table_config;
uint64_t oracle_tabulate(const oracle_scene *scene, const oracle_table_config *config, const oracle_step *steps, size_t n,
                         uint64_t *rng_x, const uint32_t *rng_a, const double reference[7], double *bins, double *squared);

/* One photon of `step` from RNG state (*x, a).  traj: NULL or room for
 * max_points * 8 floats {x,y,z,t,dx,dy,dz,abs_lens_left} recorded at creation
 * and after every segment.  Returns 1 if a record was written to *out (hit, or
 * absorption in save-all mode with prescale passed), else 0. */
int oracle_propagate_single_photon(const oracle_scene *scene, const oracle_step *step,
                                   uint64_t *x, uint32_t a, oracle_photon *out,
                                   float *traj, int max_points, int *num_points);

/* Same, for a photon whose creation draws come from the stream (x_create, a_create) and whose
 * propagation draws come from the stream (x_propagate, a_propagate): the B200 fast kernel keeps
 * two MWC streams per lane, one for photon creation and one for propagation. */
int oracle_propagate_single_photon_split(const oracle_scene *scene, const oracle_step *step,
                                         uint64_t x_create, uint32_t a_create, uint64_t x_propagate, uint32_t a_propagate,
                                         oracle_photon *out, float *traj, int max_points, int *num_points);

/* MWC RNG (mwcrng_kernel.cl:12-28): n draws of [0,1) from (*x, a). */
void oracle_rng_uniform_co(uint64_t *x, uint32_t a, float *out, size_t n);

/* Safe-prime multiplier rows [first, first+n) (make_safeprimes/main.cxx:32-104).
 * n2/n1 may be NULL. */
int oracle_safeprimes(uint64_t first, uint64_t n, uint32_t *a, uint64_t *n2, uint64_t *n1);
/* x[] seeds under the rejection rule of mwcrng_init.h:107-113 from a splitmix64 stream. */
void oracle_rng_seed_states(uint64_t seed, const uint32_t *a, uint64_t *x, size_t n);

/* Geometry tables as JSON text (same schema as clsimcu_describe_tables). */
int oracle_describe_tables(const oracle_scene *scene, char *buf, size_t cap, size_t *needed);

/* Per-function evaluators, the analogue of the reference's testers
 * (private/test/I3CLSim*Tester.cxx).  which:
 *   0 getPhaseRefIndex(layer,wlen)     1 getGroupVelocity(layer,wlen)
 *   2 getScatteringLength(layer,wlen)  3 getAbsorptionLength(layer,wlen)
 *   4 getWavelengthBias(wlen)
 * in: n rows of (layer, wlen) as floats. */
void oracle_eval_wlen_function(const oracle_scene *scene, int which, const float *layer_wlen,
                               float *out, size_t n);
/* which: 0 getTiltZShift(pos)  1 getDirectionalAbsLenCorrFactor(dir); in: n rows xyz. */
void oracle_eval_scalar_field(const oracle_scene *scene, int which, const float *xyz, float *out, size_t n);
/* which: 0 transformDirectionPreScatter  1 ...PostScatter; in/out: n rows xyz. */
void oracle_eval_vector_transform(const oracle_scene *scene, int which, const float *xyz, float *out, size_t n);
/* which: 0 makeScatteringCosAngle, 1+k generateWavelength_k; n samples from (*x, a). */
void oracle_sample(const oracle_scene *scene, int which, uint64_t *x, uint32_t a, float *out, size_t n);
/* scatterDirectionByAngle (propagation_kernel.c.cl:83-129): rows (cosa, sina, dx,dy,dz, rnd) -> xyz */
void oracle_scatter_direction(const float *in6, float *out3, size_t n);

#ifdef __cplusplus
}
#endif
#endif
