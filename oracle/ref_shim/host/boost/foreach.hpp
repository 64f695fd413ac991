// Stand-in for BOOST_FOREACH: the range-based for of C++11.
#ifndef CLSIM_REF_SHIM_FOREACH_HPP
#define CLSIM_REF_SHIM_FOREACH_HPP
#define BOOST_FOREACH(decl, range) for (decl : range)
#endif
