// Stand-in for boost/static_assert.hpp.
#ifndef CLSIM_REF_SHIM_STATIC_ASSERT_HPP
#define CLSIM_REF_SHIM_STATIC_ASSERT_HPP
#define BOOST_STATIC_ASSERT(...) static_assert(__VA_ARGS__, #__VA_ARGS__)
#endif
