// Stand-in for the sliver of boost::filesystem private/opencl/mwcrng_init.h uses: path, operator/, exists, string().
#ifndef CLSIM_REF_SHIM_FILESYSTEM_HPP
#define CLSIM_REF_SHIM_FILESYSTEM_HPP
#include <filesystem>
namespace boost { namespace filesystem {
using std::filesystem::exists;
using std::filesystem::path;
}} // namespace boost::filesystem
#endif
