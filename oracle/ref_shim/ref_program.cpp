// ref_program.cpp -- ONE complete program of the reference, assembled the way I3CLSimStepToPhotonConverterOpenCL
// assembles it (private/opencl/I3CLSimStepToPhotonConverterOpenCL.cxx:655-667) and compiled for the host.
// TEST INFRASTRUCTURE (oracle/_ref).  Nothing under clsim_b200/ links this.
//
// program.cl.inc (beside this file's include path, written at test time by oracle/pyoracle.py RefProgram) is, in
// this order: the preamble (#defines of the options), resources/kernels/mwcrng_kernel.cl, the wavelength
// generators, the wavelength bias and the medium properties as the reference's own C++ generators write them
// (oracle/_ref/libclsim_ref_medium.so: those generators compiled unmodified), the geometry as its geometry
// generator writes it (libclsim_ref_geometry.so), and resources/kernels/{propagation_kernel.h,
// sparse_collision_kernel.h, sparse_collision_kernel.c, propagation_kernel.c}.cl -- every byte of it text the
// reference ships or generates, passed through translate.py's one rewrite (vector literals).  Unlike
// ref_kernel.cpp, NOTHING in this program is supplied by the oracle's restatements.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace prog {
#include "opencl_c_shim.inc"
#include "opencl_c_shim_generated.inc"
#include "program.cl.inc"
} // namespace prog

#ifndef REF_PROGRAM_HAS_TILT
#define REF_PROGRAM_HAS_TILT 1
#endif

extern "C" {

// same meaning of `which` and the same array layouts as oracle_eval_* / oracle_sample (oracle/clsim_oracle.cpp)
void prog_eval_wlen_function(int which, const float *in, float *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        const unsigned layer = static_cast<unsigned>(in[2 * i]);
        const float w = in[2 * i + 1];
        switch (which) {
        case 0: out[i] = prog::getPhaseRefIndex(layer, w); break;
        case 1: out[i] = prog::getGroupVelocity(layer, w); break;
        case 2: out[i] = prog::getScatteringLength(layer, w); break;
        case 3: out[i] = prog::getAbsorptionLength(layer, w); break;
        default: out[i] = prog::getWavelengthBias(w); break;
        }
    }
}

void prog_eval_scalar_field(int which, const float *xyz, float *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        const prog::float4 v(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
        if (which == 0) {
#ifdef getTiltZShift_IS_CONSTANT
            out[i] = getTiltZShift_IS_CONSTANT;
#else
            out[i] = prog::getTiltZShift(v);
#endif
        } else {
            out[i] = prog::getDirectionalAbsLenCorrFactor(v);
        }
    }
}

void prog_eval_vector_transform(int which, const float *xyz, float *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        prog::float4 v(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
        if (which == 0) prog::transformDirectionPreScatter(&v);
        else prog::transformDirectionPostScatter(&v);
        out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
    }
}

void prog_sample(int which, uint64_t *x, uint32_t a, float *out, size_t n)
{
    prog::ulong rx = *x;
    prog::uint ra = a;
    for (size_t i = 0; i < n; ++i) {
        if (which == 0) out[i] = prog::makeScatteringCosAngle(&rx, &ra);
        else out[i] = prog::generateWavelength(static_cast<prog::uint>(which - 1), &rx, &ra);
    }
    *x = rx;
}

#ifndef SAVE_ALL_PHOTONS
void prog_dom_position(unsigned short string_index, unsigned short dom_index, float *xyz)
{
    prog::geometryGetDomPosition(string_index, dom_index, xyz, xyz + 1, xyz + 2);
}
#endif

// propKernel over work-items [0, n) in index order on the calling thread: one launch.  Hit records keep the kernel's
// string / DOM INDICES (the caller rewrites them to IDs as the reference's host code does after the launch).
// Returns the hit counter (may exceed cap).
uint32_t prog_propagate(const void *steps, size_t n, uint64_t *rng_x, uint32_t *rng_a, void *out, uint32_t cap, const unsigned short *layer_to_om,
                        float *history)
{
    prog::uint hitIndex = 0;
    for (size_t i = 0; i < n; ++i) {
        prog::ocl_work_item.global_id = i;
        prog::ocl_work_item.global_size = n;
        prog::propKernel(&hitIndex, cap,
#ifndef SAVE_ALL_PHOTONS
                         const_cast<unsigned short *>(layer_to_om),
#endif
                         reinterpret_cast<prog::I3CLSimStep *>(const_cast<void *>(steps)), reinterpret_cast<prog::I3CLSimPhoton *>(out),
#ifdef SAVE_PHOTON_HISTORY
                         reinterpret_cast<prog::float4 *>(history),
#endif
                         reinterpret_cast<prog::ulong *>(rng_x), rng_a);
    }
    (void)history; (void)layer_to_om;
    return hitIndex;
}

} // extern "C"
