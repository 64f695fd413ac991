// Stand-in for IceTray's umbrella header (un-vendored), just enough for the reference's geometry source generator
// (private/opencl/I3CLSimHelperGenerateGeometrySource.cxx) to compile unmodified: logging macros and the pointer
// typedef macro.  Test infrastructure (oracle/_ref), not product code.
#ifndef CLSIM_REF_SHIM_I3TRAYHEADERS_H
#define CLSIM_REF_SHIM_I3TRAYHEADERS_H
#include <cmath>     // (NAN, std::isnan: the real umbrella header brings <cmath> in)
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <iostream>   // (the real umbrella header brings the standard containers and streams in)
#include <limits>
#include <map>
#include <set>
#include <vector>

#include "boost/shared_ptr.hpp"
#include <memory>
#include <stdexcept>
#include <string>

namespace ref_shim {
template <class... A> inline void fatal(const char *fmt, A... a)
{
    char buf[1024];
    std::snprintf(buf, sizeof buf, fmt, a...);
    throw std::runtime_error(buf);
}
inline void fatal(const char *msg) { throw std::runtime_error(msg); }
} // namespace ref_shim

#define log_fatal(...) ref_shim::fatal(__VA_ARGS__)
#define log_error(...) ((void)0)
#define log_warn(...) ((void)0)
#define log_info(...) ((void)0)
#define log_notice(...) ((void)0)
#define log_debug(...) ((void)0)
#define log_trace(...) ((void)0)

#define I3_POINTER_TYPEDEFS(C)               \
    typedef boost::shared_ptr<C> C##Ptr;     \
    typedef boost::shared_ptr<const C> C##ConstPtr
#endif
