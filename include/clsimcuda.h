/*
 * clsimcuda.h -- C ABI of libclsimcuda.so, the B200 (sm_100a) step -> photon engine.
 *
 * This is the drop-in boundary for clsim's step-to-photon hot path.  Every entry
 * point below is what a reference-side I3CLSimStepToPhotonConverterCUDA (C++,
 * INTEGRATION.md) binds; each cites the reference interface it replaces.  All
 * citations are relative to the reference tree (claudiok/clsim).
 *
 * Conventions: extern "C", plain pointers and sizes, no exceptions cross the
 * boundary.  Every function returns CLSIMCU_OK (0) or a negative error code and
 * leaves a human-readable message in clsimcu_last_error() (thread-local).
 * Units follow the reference (I3Units): metres, nanoseconds, radians; wavelengths
 * in METRES (400 nm == 400e-9).
 *
 * The configuration structs carry the *model-level* numbers in double precision,
 * exactly what the reference's description objects hold (I3CLSimMediumProperties,
 * I3CLSimSimpleGeometry, I3CLSimRandomValue*, I3CLSimFunction*).  The library does
 * the flattening the reference does in its code generators
 * (private/opencl/I3CLSimHelperGenerate{Geometry,MediumProperties}Source*.cxx):
 * float-literal rounding, cell grids, string sets, z layers, quantised DOM
 * positions, cumulative wavelength tables.
 */
#ifndef CLSIMCUDA_H_INCLUDED
#define CLSIMCUDA_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLSIMCU_OK                 0
#define CLSIMCU_ERR_INVALID       -1   /* bad argument / violated precondition        */
#define CLSIMCU_ERR_UNSUPPORTED   -2   /* model class or option outside the hot path   */
#define CLSIMCU_ERR_CUDA          -3   /* CUDA runtime failure (no CPU fallback)       */
#define CLSIMCU_ERR_STATE         -4   /* wrong life-cycle state                       */
#define CLSIMCU_ERR_INTERRUPTED   -5   /* engine is shutting down                      */

/* ---- wire formats ------------------------------------------------------- */

/* 48-byte step record; bit-identical to struct I3CLSimStep
 * (resources/kernels/propagation_kernel.h.cl:52-63,
 *  public/clsim/I3CLSimStep.h:68-155). */
typedef struct clsimcu_step {
    float x, y, z, t;                 /* posAndTime                       */
    float theta, phi, length, beta;   /* dirAndLengthAndBeta              */
    uint32_t num_photons;
    float weight;
    uint32_t identifier;
    uint8_t source_type;              /* 0 = Cherenkov, >=1 = flasher #   */
    uint8_t dummy1;
    uint16_t dummy2;
} clsimcu_step;

/* 80-byte photon record; bit-identical to struct I3CLSimPhoton
 * (resources/kernels/propagation_kernel.h.cl:65-81,
 *  public/clsim/I3CLSimPhoton.h:67-213).  string_id / om_id are the real IDs
 * (the index -> ID rewrite of I3CLSimStepToPhotonConverterOpenCL.cxx:1565-1619
 * is done on the device). */
typedef struct clsimcu_photon {
    float x, y, z, t;                 /* position relative to the hit DOM, time */
    float theta, phi;                 /* direction of travel                    */
    float wavelength;
    float cherenkov_dist;
    uint32_t num_scatters;
    float weight;
    uint32_t identifier;
    int16_t string_id;
    uint16_t om_id;
    float start_x, start_y, start_z, start_t;
    float start_theta, start_phi;
    float group_velocity;
    float dist_in_abs_lens;
} clsimcu_photon;

/* ---- configuration (model-level, double precision) ---------------------- */

/* A wavelength generator (reference: I3CLSimRandomValue subclasses handed to
 * SetWlenGenerators, public/clsim/I3CLSimStepToPhotonConverter.h:90-100). */
#define CLSIMCU_WLEN_INTERP_EQUAL    0  /* I3CLSimRandomValueInterpolatedDistribution(xFirst,xSpacing,y) */
#define CLSIMCU_WLEN_INTERP_UNEQUAL  1  /* I3CLSimRandomValueInterpolatedDistribution(x,y)               */
#define CLSIMCU_WLEN_NO_DISPERSION   2  /* I3CLSimRandomValueWlenCherenkovNoDispersion(from,to)          */
#define CLSIMCU_WLEN_CONSTANT        3  /* I3CLSimRandomValueConstant(value)                             */
typedef struct clsimcu_wlen_generator {
    int32_t kind;
    int32_t n;            /* entries in y (and x)                               */
    double x0, dx;        /* INTERP_EQUAL                                       */
    const double *x;      /* INTERP_UNEQUAL                                     */
    const double *y;      /* INTERP_*: un-normalised density at the nodes       */
    double from_wlen;     /* NO_DISPERSION                                      */
    double to_wlen;
    double value;         /* CONSTANT                                           */
} clsimcu_wlen_generator;

/* Wavelength bias (reference: SetWlenBias; I3CLSimFunctionFromTable in equal
 * spacing mode, or I3CLSimFunctionConstant). */
#define CLSIMCU_BIAS_CONSTANT 0
#define CLSIMCU_BIAS_TABLE    1
typedef struct clsimcu_wlen_bias {
    int32_t kind;
    int32_t n;
    double x0, dx;
    const double *v;
    double value;         /* CONSTANT */
} clsimcu_wlen_bias;

/* Layered medium (reference: I3CLSimMediumProperties filled the way
 * python/MakeIceCubeMediumProperties.py:166-230 fills it). */
#define CLSIMCU_SCAT_MIXED_SL_HG 0   /* I3CLSimRandomValueMixed(f, SimplifiedLiu(g), HenyeyGreenstein(g)) */
#define CLSIMCU_SCAT_HG          1   /* I3CLSimRandomValueHenyeyGreenstein(g)                             */
#define CLSIMCU_SCAT_SL          2   /* I3CLSimRandomValueSimplifiedLiu(g)                                */
typedef struct clsimcu_medium {
    int32_t num_layers;
    int32_t scat_kind;
    double layers_zstart;
    double layers_height;
    /* I3CLSimFunctionAbsLenIceCube per layer (kappa,A,B,D,E equal in all layers) */
    double kappa, A, B, D, E;
    const double *a_dust400;     /* [num_layers] */
    const double *delta_tau;     /* [num_layers] */
    /* I3CLSimFunctionScatLenIceCube per layer (alpha equal in all layers) */
    double alpha;
    const double *b400;          /* [num_layers] */
    /* I3CLSimFunctionRefIndexIceCube, phase + group override (layer independent) */
    double n_phase[5];
    double n_group[5];
    /* scattering angle distribution */
    double f_sl;                 /* fractionOfFirstDistribution (MIXED) */
    double mean_cos;             /* g */
    /* ice tilt: I3CLSimScalarFieldIceTiltZShift, or constant 0 when tilt_num_dist == 0 */
    int32_t tilt_num_dist;
    int32_t tilt_num_z;
    const double *tilt_dist;     /* [tilt_num_dist] distancesFromOriginAlongTilt      */
    const double *tilt_corr;     /* [tilt_num_dist][tilt_num_z] zCorrections           */
    double tilt_z0, tilt_dz;     /* firstZCoordinate, zCoordinateSpacing               */
    double tilt_azimuth;         /* directionOfTiltAzimuth [rad]                       */
    /* anisotropy: I3CLSimScalarFieldAnisotropyAbsLenScaling + 2x I3CLSimVectorTransformMatrix */
    int32_t has_anisotropy;
    int32_t pre_renormalize, post_renormalize;
    int32_t reserved0;
    double aniso_azimuth;        /* anisotropyDirAzimuth [rad] */
    double aniso_along;          /* magnitudeAlongDir          */
    double aniso_perp;           /* magnitudePerpToDir         */
    double pre_matrix[9];        /* row major                  */
    double post_matrix[9];
} clsimcu_medium;

/* Flat DOM list (reference: I3CLSimSimpleGeometry, public/clsim/I3CLSimSimpleGeometry.h). */
typedef struct clsimcu_geometry {
    int32_t num_doms;
    int32_t reserved0;
    const int32_t *string_id;    /* [num_doms] */
    const uint32_t *dom_id;      /* [num_doms] */
    const double *x, *y, *z;     /* [num_doms] */
    const int32_t *subdetector;  /* [num_doms] rank of the subdetector NAME in sorted order
                                    (the reference keys a std::set<std::string>) */
    double om_radius;            /* includes the oversize factor */
} clsimcu_geometry;

#define CLSIMCU_KERNEL_FAST       0  /* persistent kernel, per-lane RNG streams (default)        */
#define CLSIMCU_KERNEL_REFERENCE  1  /* one thread per step, the reference's RNG-stream mapping,
                                        precise math; all options supported                      */

typedef struct clsimcu_config {
    int32_t struct_size;         /* sizeof(clsimcu_config), ABI check                          */
    int32_t device;              /* CUDA ordinal (SetDevice)                                   */
    int32_t kernel_mode;         /* CLSIMCU_KERNEL_*                                           */
    int32_t enable_double_buffering;      /* SetEnableDoubleBuffering                          */
    int32_t stop_detected_photons;        /* SetStopDetectedPhotons                            */
    int32_t save_all_photons;             /* SetSaveAllPhotons                                 */
    int32_t photon_history_entries;       /* SetPhotonHistoryEntries                           */
    int32_t num_wlen_generators;
    double save_all_photons_prescale;     /* SetSaveAllPhotonsPrescale                         */
    double fixed_number_of_absorption_lengths; /* NaN = sample (SetFixedNumberOfAbsorptionLengths) */
    double pancake_factor;                /* SetDOMPancakeFactor                               */
    uint64_t max_num_workitems;  /* largest bunch accepted by clsimcu_enqueue (SetMaxNumWorkitems) */
    uint32_t workgroup_size;     /* bunch-size granularity advertised (SetWorkgroupSize); 0 -> 1 */
    uint32_t output_photons_per_workitem; /* output capacity factor; 0 -> 10 like the reference */
    const clsimcu_wlen_generator *wlen_generators;
    clsimcu_wlen_bias wlen_bias;
    clsimcu_medium medium;
    clsimcu_geometry geometry;   /* ignored when save_all_photons                              */
    /* MWC RNG (private/opencl/mwcrng_init.h:26-117).  rng_n streams.
       rng_a/rng_x given: used verbatim.  rng_a == NULL: multipliers are the
       safe-prime sequence (private/make_safeprimes/main.cxx) starting at row
       rng_first_multiplier, x[] drawn from rng_seed under the reference's
       rejection rule. */
    uint64_t rng_n;
    const uint32_t *rng_a;
    const uint64_t *rng_x;
    uint64_t rng_seed;
    uint64_t rng_first_multiplier;
} clsimcu_config;

typedef struct clsimcu_engine clsimcu_engine;

/* ---- life cycle ---------------------------------------------------------- */

/* Replaces: the setter block + Compile + Initialize of
 * I3CLSimStepToPhotonConverterOpenCL (private/opencl/I3CLSimStepToPhotonConverterOpenCL.cxx:217-388,
 * 485-548; factory private/clsim/I3CLSimModuleHelper.cxx:303-372).  Builds the tables,
 * uploads them, seeds the RNG, allocates device/pinned buffers and starts the
 * worker thread.  Fails (never falls back) when no CUDA device is usable. */
int clsimcu_create(const clsimcu_config *config, clsimcu_engine **engine);

/* Replaces: ~I3CLSimStepToPhotonConverterOpenCL (…OpenCL.cxx:110-145): interrupts and
 * joins the worker, frees device memory. */
int clsimcu_destroy(clsimcu_engine *engine);

/* ---- the hot calls -------------------------------------------------------- */

/* Replaces: EnqueueSteps (…OpenCL.cxx:1525-1544).  Copies n 48-byte steps out of
 * the caller's buffer; blocks while 5 bunches are already queued.  Errors like the
 * reference: n == 0, n > max_num_workitems, n % workgroup_size != 0. */
int clsimcu_enqueue(clsimcu_engine *engine, const clsimcu_step *steps, size_t n, uint32_t identifier);

/* Replaces: GetConversionResult (…OpenCL.cxx:1604-1619).  Blocks until a bunch is
 * done.  *photons is never NULL on success (I3CLSimClientModule.cxx:589 requires
 * it); *history is NULL unless photon_history_entries > 0, else
 * n_photons * photon_history_entries float4 rows already in forward order with
 * unused rows NaN (…OpenCL.cxx:940-989).  Release with clsimcu_release_result. */
typedef struct clsimcu_result {
    uint32_t identifier;
    uint32_t reserved0;
    size_t num_photons;
    clsimcu_photon *photons;
    float *history;              /* [num_photons][photon_history_entries][4] or NULL */
    uint64_t num_photons_generated; /* sum of step.num_photons of the bunch */
    uint64_t num_hits_counted;   /* device counter; > num_photons means truncated (…OpenCL.cxx:1027-1032) */
    void *opaque;
    /* photon -> MCPE conversion on the device (clsimcu_attach_mcpe_converter below); NULL / 0 otherwise */
    struct clsimcu_mcpe *mcpes;
    size_t num_mcpes;
} clsimcu_result;
int clsimcu_get_result(clsimcu_engine *engine, clsimcu_result *result);
int clsimcu_release_result(clsimcu_engine *engine, clsimcu_result *result);

/* Replaces: QueueSize / MorePhotonsAvailable (…OpenCL.cxx:1546-1562). */
int clsimcu_queue_size(clsimcu_engine *engine, size_t *size);
int clsimcu_more_photons_available(clsimcu_engine *engine, int *available);

/* Replaces: GetWorkgroupSize / GetMaxNumWorkitems (sizing handshake,
 * I3CLSimServer.cxx:95-115). */
int clsimcu_workgroup_size(clsimcu_engine *engine, size_t *size);
int clsimcu_max_num_workitems(clsimcu_engine *engine, size_t *size);

/* Replaces: GetStatistics (…OpenCL.cxx:1621-1640).  out[0..7] =
 * TotalDeviceTime[ns], TotalHostTime[ns], NumKernelCalls, TotalNumPhotonsGenerated,
 * TotalNumPhotonsAtDOMs, AverageDeviceTimePerPhoton, AverageHostTimePerPhoton,
 * DeviceUtilization. */
int clsimcu_get_statistics(clsimcu_engine *engine, double out[8]);

/* ---- device-resident path (no reference equivalent; measurement only) ------ */

/* Upload a bunch once, then run the propagation kernel `repeat` times on it with
 * inputs resident in HBM.  Timed with CUDA events on the engine's stream.
 * Outputs: kernel milliseconds (sum over repeats), photons generated and hits
 * counted over all repeats.  The RNG streams advance between repeats. */
int clsimcu_upload_resident(clsimcu_engine *engine, const clsimcu_step *steps, size_t n);
int clsimcu_run_resident(clsimcu_engine *engine, int repeat, double *kernel_ms,
                         uint64_t *photons_generated, uint64_t *hits_counted,
                         uint64_t *segments);
/* Copy the hits of the last resident run back (at most cap records). */
int clsimcu_download_resident(clsimcu_engine *engine, clsimcu_photon *out, size_t cap, size_t *n);
/* ... and, with photon_history_entries > 0, their scatter-point histories: cap * photon_history_entries rows of
 * four floats in forward order, unused rows NaN, like clsimcu_result::history (…OpenCL.cxx:940-989). */
int clsimcu_download_resident_history(clsimcu_engine *engine, float *out, size_t cap, size_t *n);

/* ---- test hooks ------------------------------------------------------------- */

/* RNG state of the first n streams (reference keeps it on the device between
 * launches, propagation_kernel.c.cl:458-459, 911-912). */
int clsimcu_rng_get(clsimcu_engine *engine, uint64_t *x, uint32_t *a, size_t n);
int clsimcu_rng_set(clsimcu_engine *engine, const uint64_t *x, const uint32_t *a, size_t n);

/* Flattened geometry tables (what GenerateGeometrySource emits as constants),
 * serialised as a JSON text for comparison with the oracle's builder. */
int clsimcu_describe_tables(clsimcu_engine *engine, char *buf, size_t cap, size_t *needed);

/* Table building without a GPU (host only): same JSON as above. */
int clsimcu_describe_tables_from_config(const clsimcu_config *config, char *buf, size_t cap, size_t *needed);

/* The fast kernel's xy collision map for a given pixel budget (host only, no GPU needed), as a JSON text: per pixel
 * the nearest string (16 x its index, low 16 bits) and the range within which no other string can be touched (upper 16
 * bits of an fp32).  No counterpart in the reference (its cell grids, sparse_collision_kernel.c.cl:194-460, are what
 * the map prunes); exported so that the guarantees the kernel relies on can be tested without a device. */
int clsimcu_describe_collision_map_from_config(const clsimcu_config *config, int32_t pixel_budget, char *buf, size_t cap, size_t *needed);

/* Safe-prime MWC multipliers (private/make_safeprimes/main.cxx:32-104): writes the
 * rows [first, first+n) of the descending sequence that starts at 4294967118. */
int clsimcu_safeprime_multipliers(uint64_t first, uint64_t n, uint32_t *a);

/* Start states x[0..n) of the MWC streams with multipliers a[0..n), as clsimcu_create draws them from config->rng_seed
 * (host only, no GPU needed).  The acceptance rule is init_MWC_RNG's (private/opencl/mwcrng_init.h:107-113: x != 0,
 * upper half < a - 1, lower half < 0xffffffff, two 32-bit draws per candidate); the draws come from splitmix64 in place
 * of the reference's I3RandomService (IceTray's GSL service, un-vendored). */
int clsimcu_seed_rng_states(uint64_t seed, const uint32_t *a, uint64_t *x, uint64_t n);

/* In SAVE_ALL mode with kernel FAST: per saved photon i the two MWC stream states it was made
 * from, so a CPU checker can replay single photons.  The fast kernel keeps two streams per
 * lane, one for photon creation and one for propagation:
 *   x[2*i]   state of the creation stream before the photon was created,   a[2*i]   its multiplier,
 *   x[2*i+1] state of the propagation stream when the photon started,      a[2*i+1] its multiplier.
 * Parallel to the last result of clsimcu_download_resident; x and a have room for 2*cap entries. */
int clsimcu_download_resident_rng_tags(clsimcu_engine *engine, uint64_t *x, uint32_t *a, size_t cap);

/* ---- photon -> MCPE on the device -------------------------- */

/* One photo-electron: I3MCPE(ParticleID, npe, time) (simclasses/I3MCPE.h as used at
 * private/clsim/dom/I3PhotonToMCPEConverter.cxx:521, 668) together with the OMKey of the
 * series it belongs to.  `identifier` is the step identifier of the photon; the caller maps
 * it to the particle's major/minor ID (private/clsim/I3CLSimClientModule.cxx:385-412). */
typedef struct clsimcu_mcpe {
    int16_t string_id;
    uint16_t om_id;
    float time;
    uint32_t npe;
    uint32_t identifier;
} clsimcu_mcpe;

/* CLSIMCU_MCPE_INLOOP: I3CLSimPhotonToMCPEConverterForDOMs::Convert
 *   (private/clsim/dom/I3PhotonToMCPEConverter.cxx:602-669): p = weight * acceptance[OM](wavelength)
 *   * angular(-dir.z); the photon position (relative to the DOM) must lie within 3 cm of the
 *   165.1 mm sphere; time unchanged.
 * CLSIMCU_MCPE_MODULE: I3PhotonToMCPEConverter::Convert (…cxx:395-523): cos = -(dir . dom_dir),
 *   p = weight * acceptance(wavelength) * angular(cos) * efficiency[OM]; position check only for
 *   pancake_factor == 1 against dom_radius * oversize_factor; time += (p . dir) * (1 - pancake /
 *   oversize) / groupVelocity with p = DOM centre - photon position.
 * Both: weight < 0 and p > 1 are fatal in the reference and are errors here; weight == 0 is
 * skipped; the photon survives when p > u, u uniform in [0,1). */
#define CLSIMCU_MCPE_INLOOP 0
#define CLSIMCU_MCPE_MODULE 1
typedef struct clsimcu_mcpe_config {
    int32_t struct_size;          /* sizeof(clsimcu_mcpe_config) */
    int32_t device;
    int32_t flavour;              /* CLSIMCU_MCPE_* */
    int32_t num_acceptances;      /* wavelength acceptance curves (I3CLSimFunctionFromTable / Constant) */
    const clsimcu_wlen_bias *acceptances;
    int32_t num_doms;
    int32_t num_angular_coefficients;
    const int32_t *string_id;     /* [num_doms] */
    const uint32_t *dom_id;       /* [num_doms] */
    const uint8_t *acceptance_of_dom;    /* [num_doms] index into acceptances; NULL = all 0 */
    const double *efficiency_of_dom;     /* [num_doms] MODULE: relative DOM efficiency x SPE compensation; NULL = all 1 */
    const double *angular_coefficients;  /* I3CLSimFunctionPolynomial(coefficients), argument cos(angle) */
    double dom_dir[3];            /* MODULE: PMT axis, (0,0,-1) for IceCube */
    double dom_radius;            /* without oversize (165.1 mm) */
    double oversize_factor, pancake_factor;
    int32_t only_warn_about_positions;   /* OnlyWarnAboutInvalidPhotonPositions */
    int32_t reserved0;
    /* MWC streams of the thinning draws: rows [rng_first_multiplier, +num streams) of the
       safe-prime table, states drawn from rng_seed like clsimcu_config's */
    uint64_t rng_seed;
    uint64_t rng_first_multiplier;
} clsimcu_mcpe_config;

typedef struct clsimcu_mcpe_converter clsimcu_mcpe_converter;

int clsimcu_mcpe_create(const clsimcu_mcpe_config *config, clsimcu_mcpe_converter **converter);
int clsimcu_mcpe_destroy(clsimcu_mcpe_converter *converter);

/* Converts a photon series held in HOST memory (copies in, one kernel, copies out).  uniforms:
 * NULL = the converter's MWC streams; else one explicit draw per photon (test hook: makes the
 * survivors a pure function of the inputs).  Survivors are written in photon order; *n_out may
 * exceed cap (then only cap records were written). */
int clsimcu_mcpe_convert(clsimcu_mcpe_converter *converter, const clsimcu_photon *photons, size_t n, const float *uniforms,
                         clsimcu_mcpe *out, size_t cap, size_t *n_out);

/* Number of MWC streams of a converter and its states (test hook: photon j is thinned with
 * draw number j / streams of stream j % streams). */
int clsimcu_mcpe_rng_get(clsimcu_mcpe_converter *converter, uint64_t *x, uint32_t *a, size_t cap, size_t *streams);

/* Runs the conversion on the engine's stream right after every propagation launch, on the hits
 * while they are still in HBM: results then carry `mcpes` (and, unless keep_photons, no photons:
 * 16 bytes per surviving photo-electron cross PCIe instead of 80 per photon; the docs ask for
 * exactly this, resources/docs/clsim_server_worklist.txt:85-124).  Call before the first
 * clsimcu_enqueue; the converter must outlive the engine and be on the same device. */
int clsimcu_attach_mcpe_converter(clsimcu_engine *engine, clsimcu_mcpe_converter *converter, int keep_photons);

/* ---- light source -> steps on the device ---------------------- */

/* One entry of the reference's step generation queue: what EnqueueLightSource leaves behind for
 * MakeSteps (CascadeStepData_t / MuonStepData_t, public/clsim/I3CLSimLightSourceToStepConverterPPC.h;
 * filled at private/clsim/I3CLSimLightSourceToStepConverterPPC.cxx:336-366, 437-471).  The yield
 * (number of photons, hence steps) is the caller's business, exactly as in the reference, where it
 * needs I3RandomService and sim-services' shower parameterisation; the per-step work -- the part
 * that scales with the light yield -- is done on the device:
 *   CLSIMCU_SOURCE_CASCADE             longitudinal position pb * Gamma(pa) [m] along the axis, direction
 *                                      smeared (…PPC.cxx:523-537, 785-819)
 *   CLSIMCU_SOURCE_TRACK_CASCADE_LIKE  position uniform in [0, length), direction smeared (…cxx:546-556)
 *   CLSIMCU_SOURCE_TRACK_MUON_LIKE     one step of the whole length at the vertex (…cxx:557-563, 821-842)
 * An entry makes num_steps steps of photons_per_step photons plus, if photons_in_last_step > 0, one
 * more with that many (…cxx:566-607). */
#define CLSIMCU_SOURCE_CASCADE            0
#define CLSIMCU_SOURCE_TRACK_CASCADE_LIKE 1
#define CLSIMCU_SOURCE_TRACK_MUON_LIKE    2
typedef struct clsimcu_step_source {
    double x, y, z, t;            /* particle vertex [m], time [ns]     */
    double dir_x, dir_y, dir_z;   /* direction of travel (unit vector)  */
    double length;                /* TRACK_*: track length [m]          */
    double pa, pb;                /* CASCADE: gamma shape, scale [m]    */
    uint64_t num_steps;
    uint32_t photons_per_step;
    uint32_t photons_in_last_step;
    uint32_t identifier;
    int32_t kind;                 /* CLSIMCU_SOURCE_* */
} clsimcu_step_source;

typedef struct clsimcu_step_generator_config {
    int32_t struct_size;          /* sizeof(clsimcu_step_generator_config) */
    int32_t device;
    double angular_a, angular_b;  /* cascade angular smearing, 0.39 and 2.61 in the reference (…PPC.cxx:105) */
    /* MWC streams, one per generator thread: rows [rng_first_multiplier, +streams) of the safe-prime
       table, states drawn from rng_seed like clsimcu_config's */
    uint64_t rng_seed;
    uint64_t rng_first_multiplier;
} clsimcu_step_generator_config;

typedef struct clsimcu_step_generator clsimcu_step_generator;

int clsimcu_stepgen_create(const clsimcu_step_generator_config *config, clsimcu_step_generator **generator);
int clsimcu_stepgen_destroy(clsimcu_step_generator *generator);

/* Makes the steps of n queue entries and copies them to the host: entry after entry, each entry's
 * partial last step at the end of its run.  *n_out = total number of steps (may exceed cap).  With T
 * streams, step j draws from stream j % T, the steps of one stream in ascending order
 * (clsimcu_stepgen_rng_get: test hook for replaying that). */
int clsimcu_stepgen_generate(clsimcu_step_generator *generator, const clsimcu_step_source *sources, size_t n,
                             clsimcu_step *out, size_t cap, size_t *n_out);
int clsimcu_stepgen_rng_get(clsimcu_step_generator *generator, uint64_t *x, uint32_t *a, size_t cap, size_t *streams);

/* EnqueueSteps for steps that do not exist yet: the bunch is generated on the device from the n queue
 * entries (at most 65536 per bunch) and propagated without ever visiting the host: a few KB go over
 * PCIe instead of 48 bytes per step.  The total number of steps must respect max_num_workitems and
 * the workgroup size like clsimcu_enqueue.  Results come back through clsimcu_get_result. */
int clsimcu_enqueue_sources(clsimcu_engine *engine, clsimcu_step_generator *generator, const clsimcu_step_source *sources,
                            size_t n, uint32_t identifier);

/* ---- table-maker variant -------------------------------------- */

/* The -DTABULATE build of the reference kernel (resources/kernels/propagation_kernel.c.cl:226-304, 755-785)
 * behind I3CLSimStepToTableConverter (private/clsim/tabulator/I3CLSimStepToTableConverter.cxx): photons are
 * propagated for a fixed 42 absorption lengths through DOM-free ice, and every `step_length` metres along
 * their path the detection probability exp(-absorption lengths so far) * angular acceptance is added to the
 * bin of a photon-density table around a reference particle.
 * Re-designed for the B200: the table lives in HBM and the kernel adds to it with float atomics.  The
 * reference's per-work-item entry buffers, its "out of space, restart the photon" protocol and the host loop
 * that sums the entries (…StepToTableConverter.cxx:375-520) have no counterpart: the result is the same sum
 * in a different order. */
#define CLSIMCU_AXIS_LINEAR 0     /* clsim::tabulator::LinearAxis (private/clsim/tabulator/Axis.cxx:80-111) */
#define CLSIMCU_AXIS_POWER  1     /* clsim::tabulator::PowerAxis  (…Axis.cxx:113-171) */
typedef struct clsimcu_axis {
    int32_t kind;
    uint32_t power;               /* POWER */
    double min, max;
    uint32_t n_bins;
    uint32_t reserved0;
} clsimcu_axis;

#define CLSIMCU_TABLE_SPHERICAL   0   /* SphericalAxes: radius, azimuth [deg], cos(polar), delay time (+ impact cos) */
#define CLSIMCU_TABLE_CYLINDRICAL 1   /* CylindricalAxes: perpendicular distance, azimuth [rad], depth, delay time (+ impact cos) */
typedef struct clsimcu_tabulator_config {
    int32_t struct_size;          /* sizeof(clsimcu_tabulator_config) */
    int32_t geometry;             /* CLSIMCU_TABLE_* */
    int32_t num_axes;             /* 4, or 5 to tabulate the impact angle as well (TABULATE_IMPACT_ANGLE) */
    int32_t store_squared_weights;
    clsimcu_axis axes[5];
    double step_length;           /* VOLUME_MODE_STEP, 1 m in the reference (stepLength_) */
    double reference_area;        /* domArea_, used by the normalisation */
    int32_t num_angular_coefficients;   /* getAngularAcceptance: I3CLSimFunctionPolynomial (4-axis tables) */
    int32_t reserved0;
    const double *angular_coefficients;
    double min_wavelength, max_wavelength;   /* mediumProperties->GetMin/MaxWavelength(); 0 = 265 nm / 675 nm */
} clsimcu_tabulator_config;

/* I3CLSimReferenceParticle(source) (…StepToTableConverter.cxx:65-93): position, time and direction of the
 * particle the coordinates refer to */
typedef struct clsimcu_reference_particle {
    double x, y, z, t;
    double dir_x, dir_y, dir_z;
} clsimcu_reference_particle;

typedef struct clsimcu_tabulator clsimcu_tabulator;

/* `scene`: wavelength generators, wavelength acceptance (wlen_bias), medium, RNG and device as for
 * clsimcu_create; geometry and the photon options are ignored (the tabulator sets SAVE_ALL_PHOTONS,
 * prescale 1 and 42 fixed absorption lengths like the reference's preamble, …cxx:177-183). */
int clsimcu_tabulator_create(const clsimcu_config *scene, const clsimcu_tabulator_config *config, clsimcu_tabulator **tabulator);
int clsimcu_tabulator_destroy(clsimcu_tabulator *tabulator);

/* Replaces: EnqueueSteps(steps, reference) (…cxx:290-302) + the harvester thread's launch: propagates the
 * bunch and adds its photons to the table.  n <= scene->max_num_workitems.  Returns when the launch is queued;
 * clsimcu_tabulator_finish (Finish, …cxx:304-312) waits for everything. */
int clsimcu_tabulator_enqueue(clsimcu_tabulator *tabulator, const clsimcu_step *steps, size_t n, const clsimcu_reference_particle *reference);
int clsimcu_tabulator_finish(clsimcu_tabulator *tabulator);

/* Table geometry and header values: out[0] = number of bins (over-/underflow included), out[1..5] = shape,
 * out[6..10] = strides, out[11] = n_photons (spectralBiasFactor * sum of photon weights), out[12] = n_group,
 * out[13] = n_phase (minimum refractive indices, …cxx:96-121), out[14] = spectralBiasFactor, out[15] = photons
 * propagated so far. */
int clsimcu_tabulator_info(clsimcu_tabulator *tabulator, double out[16]);

/* Copies the table (and the squared weights, if stored and asked for) to the host after Finish.  normalize != 0
 * applies Normalize() (…cxx:522-556) to the copy: bin volume / (step_length * reference_area). */
int clsimcu_tabulator_get_table(clsimcu_tabulator *tabulator, int normalize, float *bins, float *squared_weights, size_t cap);

/* Number of usable CUDA devices (reference: I3CLSimOpenCLDevice::GetAllDevices,
 * private/opencl/I3CLSimOpenCLDevice.cxx, as used by python/traysegments/common.py:10-77).
 * 0 devices is an error: there is no CPU fallback. */
int clsimcu_device_count(int *count);

const char *clsimcu_last_error(void);
const char *clsimcu_version(void);
size_t clsimcu_sizeof_config(void);

#ifdef __cplusplus
}
#endif
#endif /* CLSIMCUDA_H_INCLUDED */
