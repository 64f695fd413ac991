// Stand-in for boost::lexical_cast (boost is not in this image): number -> string through a stream, as boost does.
#ifndef CLSIM_REF_SHIM_LEXICAL_CAST_HPP
#define CLSIM_REF_SHIM_LEXICAL_CAST_HPP
#include <sstream>
#include <string>
namespace boost {
template <class Target, class Source> inline Target lexical_cast(const Source &v)
{
    std::stringstream s;
    s << v;
    Target out;
    s >> out;
    return out;
}
} // namespace boost
#endif
