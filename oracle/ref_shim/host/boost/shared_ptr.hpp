// Stand-in: boost::shared_ptr is std::shared_ptr here (I3_POINTER_TYPEDEFS in icetray/I3TrayHeaders.h uses the same type).
#ifndef CLSIM_REF_SHIM_SHARED_PTR_HPP
#define CLSIM_REF_SHIM_SHARED_PTR_HPP
#include <memory>
namespace boost {
using std::const_pointer_cast;
using std::dynamic_pointer_cast;
using std::make_shared;
using std::shared_ptr;
using std::static_pointer_cast;
} // namespace boost
#endif
