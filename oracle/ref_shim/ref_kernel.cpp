// ref_kernel.cpp -- the reference's own propKernel, compiled for the host.  TEST INFRASTRUCTURE (oracle/_ref).
//
// What this is: resources/kernels/{mwcrng_kernel, propagation_kernel.h, propagation_kernel.c,
// sparse_collision_kernel.h, sparse_collision_kernel.c}.cl of /root/reference, read where they lie at BUILD time,
// passed through translate.py (one syntax rewrite, vector literals; the build log lists the 13 lines) and compiled
// by g++ under opencl_c_shim.inc.  It exists to PIN oracle/clsim_oracle.cpp -- the hand-written restatement the
// GPU parity tests check against -- to the reference's kernel text: tests/test_ref_kernel.py asserts that both
// produce bit-identical hit lists and final RNG states on the same steps and the same (x, a) streams.
//
// What it is not: the run-time generated parts of the reference's program (wavelength generators, bias, medium
// functions, geometry tables) come from the oracle's restatements (see ref_variant.inc); they are pinned
// separately against reference data (tests/test_golden_inputs.py).  Nothing under clsim_b200/ links this.
//
// The oracle's translation unit is included for its scene, table builders and generated-function restatements;
// its extern "C" entry points are compiled in as well (the library is loaded RTLD_LOCAL by ctypes).
#include "../clsim_oracle.cpp"

#include <cmath>
#include <cstdint>
#include <cstring>

// ---- variants: the #ifdef axes of the kernel text.  name = options
#define REF_NS ref_stop_pancake_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_NO_FLASHER
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_tilt
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_NO_FLASHER
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_flasher_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_flasher_tilt
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#include "ref_variant.inc"

#define REF_NS ref_stop_flasher_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_stop_flasher_tilt
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#include "ref_variant.inc"

#define REF_NS ref_stop_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_NO_FLASHER
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_stop_tilt
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_NO_FLASHER
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_history_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_HISTORY
#define REF_OPT_NO_FLASHER
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_history_tilt
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_HISTORY
#define REF_OPT_NO_FLASHER
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_fixedabs_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_FIXED_ABS
#define REF_OPT_NO_FLASHER
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_nonstop_pancake_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_PANCAKE
#define REF_OPT_NO_FLASHER
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_nonstop_pancake_tilt
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_PANCAKE
#define REF_OPT_NO_FLASHER
#include "ref_variant.inc"

#define REF_NS ref_saveall_tiltconst
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_SAVE_ALL
#define REF_OPT_NO_FLASHER
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_saveall_tilt
#define REF_OPT_SUBDETECTORS 1
#define REF_OPT_SAVE_ALL
#define REF_OPT_NO_FLASHER
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_tiltconst_sub9
#define REF_OPT_SUBDETECTORS 9
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_NO_FLASHER
#define REF_OPT_TILT_CONSTANT
#include "ref_variant.inc"

#define REF_NS ref_stop_pancake_tilt_sub9
#define REF_OPT_SUBDETECTORS 9
#define REF_OPT_STOP
#define REF_OPT_PANCAKE
#define REF_OPT_NO_FLASHER
#include "ref_variant.inc"

namespace {

typedef uint32_t (*RunFn)(const oracle_scene *, const oracle_step *, size_t, size_t, size_t, uint64_t *, uint32_t *, oracle_photon *, uint32_t,
                          float *);
struct Variant {
    const char *name;
    bool stop, save_all, history, fixed_abs, pancake, flasher, tilt;
    int subdetectors;
    RunFn run;
};
#define V(ns, stop, saveall, hist, fixed, pancake, flasher, tilt) {#ns, stop, saveall, hist, fixed, pancake, flasher, tilt, 1, &ns::run_work_items}
#define V9(ns, stop, saveall, hist, fixed, pancake, flasher, tilt) {#ns, stop, saveall, hist, fixed, pancake, flasher, tilt, 9, &ns::run_work_items}
const Variant kVariants[] = {
    V(ref_stop_pancake_tiltconst, true, false, false, false, true, false, false),
    V(ref_stop_pancake_tilt, true, false, false, false, true, false, true),
    V(ref_stop_pancake_flasher_tiltconst, true, false, false, false, true, true, false),
    V(ref_stop_pancake_flasher_tilt, true, false, false, false, true, true, true),
    V(ref_stop_flasher_tiltconst, true, false, false, false, false, true, false),
    V(ref_stop_flasher_tilt, true, false, false, false, false, true, true),
    V(ref_stop_tiltconst, true, false, false, false, false, false, false),
    V(ref_stop_tilt, true, false, false, false, false, false, true),
    V(ref_stop_pancake_history_tiltconst, true, false, true, false, true, false, false),
    V(ref_stop_pancake_history_tilt, true, false, true, false, true, false, true),
    V(ref_stop_pancake_fixedabs_tiltconst, true, false, false, true, true, false, false),
    V(ref_nonstop_pancake_tiltconst, false, false, false, false, true, false, false),
    V(ref_nonstop_pancake_tilt, false, false, false, false, true, false, true),
    V(ref_saveall_tiltconst, false, true, false, false, false, false, false),
    V(ref_saveall_tilt, false, true, false, false, false, false, true),
    V9(ref_stop_pancake_tiltconst_sub9, true, false, false, false, true, false, false),
    V9(ref_stop_pancake_tilt_sub9, true, false, false, false, true, false, true),
};
#undef V
#undef V9

const Variant *find_variant(const oracle_scene &sc, bool flasher)
{
    const bool tilt = sc.med.tiltND > 0;
    for (const Variant &v : kVariants) {
        if (v.stop != sc.stop || v.save_all != sc.saveAll || v.history != (sc.history > 0) || v.fixed_abs != sc.fixedAbs || v.tilt != tilt) continue;
        if (!sc.saveAll && v.pancake != sc.pancake) continue;
        // a kernel compiled with flasher support also runs Cherenkov steps; prefer the exact match
        if (v.flasher != flasher) continue;
        if (!sc.saveAll && static_cast<size_t>(v.subdetectors) < sc.geo.cells.size()) continue;   // first fit: the smallest unrolling that covers the scene
        return &v;
    }
    return nullptr;
}

} // namespace

extern "C" {

// The variant of the reference kernel (its set of preprocessor options) that `scene` with these steps selects, or
// NULL when that combination was not compiled into this library.
const char *ref_variant_name(const oracle_scene *scene, int with_flasher)
{
    const Variant *v = find_variant(*scene, with_flasher != 0);
    return v ? v->name : nullptr;
}

// propKernel of the reference over n work-items (one launch): same contract as oracle_propagate.  Hits come out
// in (work-item, emission) order; string/DOM indices are rewritten to IDs as the reference's host code does after
// the launch (I3CLSimStepToPhotonConverterOpenCL.cxx:1565-1602).  with_flasher: compile-time -DNO_FLASHER absent.
// Returns the hit counter (may exceed cap), or UINT64_MAX when the variant does not exist.
uint64_t ref_propagate(const oracle_scene *scene, const oracle_step *steps, size_t n, uint64_t *rng_x, const uint32_t *rng_a,
                       oracle_photon *out, size_t cap, float *history, int with_flasher, int num_threads)
{
    const oracle_scene &sc = *scene;
    const Variant *v = find_variant(sc, with_flasher != 0);
    if (!v) {
        g_last_error = "oracle/_ref: this combination of kernel options was not compiled";
        return UINT64_MAX;
    }
    std::vector<uint32_t> a(rng_a, rng_a + n);   // the kernel stores a[] back
    auto rewrite_ids = [&](oracle_photon *p, size_t count) {
        if (sc.saveAll) return;
        for (size_t k = 0; k < count; ++k) {
            const unsigned short s = static_cast<unsigned short>(p[k].string_id), d = p[k].om_id;
            p[k].string_id = static_cast<int16_t>(sc.geo.stringIndexToID.at(s));
            p[k].om_id = static_cast<uint16_t>(sc.geo.domIndexToID.at(s).at(d));
        }
    };
    if (num_threads <= 1) {
        const uint32_t count = v->run(scene, steps, 0, n, n, rng_x, a.data(), out, static_cast<uint32_t>(cap), history);
        rewrite_ids(out, std::min<size_t>(count, cap));
        return count;
    }
    // chunks of work-items with their own output buffer each, concatenated in order: the same list as one thread's
    const size_t chunk = 256, numChunks = (n + chunk - 1) / chunk;
    const int H = sc.history;
    std::vector<std::vector<oracle_photon>> hits(numChunks);
    std::vector<std::vector<float>> hist(numChunks);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads)
#endif
    for (long long c = 0; c < static_cast<long long>(numChunks); ++c) {
        const size_t lo = c * chunk, hi = std::min(n, lo + chunk);
        uint64_t photons = 0;
        for (size_t i = lo; i < hi; ++i) photons += steps[i].num_photons;
        // (stop mode: at most one hit per photon; non-stop mode may exceed this and then drops like any full buffer)
        const size_t room = std::min<uint64_t>(std::max<uint64_t>(photons, 16) * (sc.stop || sc.saveAll ? 1 : 4), 0xffffffffu);
        hits[c].resize(room);
        if (H > 0) hist[c].resize(room * 4 * H);
        const uint32_t count = v->run(scene, steps, lo, hi, n, rng_x, a.data(), hits[c].data(), static_cast<uint32_t>(room), H > 0 ? hist[c].data() : nullptr);
        hits[c].resize(std::min<size_t>(count, room));
    }
    uint64_t count = 0;
    for (size_t c = 0; c < numChunks; ++c) {
        for (size_t k = 0; k < hits[c].size(); ++k) {
            if (count < cap && out) {
                out[count] = hits[c][k];
                if (H > 0 && history) std::memcpy(history + count * 4 * H, &hist[c][k * 4 * H], sizeof(float) * 4 * H);
            }
            ++count;
        }
    }
    rewrite_ids(out, std::min<size_t>(count, cap));
    return count;
}

} // extern "C"
