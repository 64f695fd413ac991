// Stand-in for the OpenCL C++ header the reference vendors as public/clsim/cl.hpp (it needs CL/cl.h, which does not
// exist in this image): the generator only uses the scalar typedefs.
#ifndef CLSIM_REF_SHIM_CL_HPP
#define CLSIM_REF_SHIM_CL_HPP
#include <cstdint>
typedef uint16_t cl_ushort;
typedef int16_t cl_short;
typedef uint32_t cl_uint;
typedef int32_t cl_int;
typedef float cl_float;
typedef double cl_double;
#endif
