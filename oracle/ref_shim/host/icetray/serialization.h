// Stand-in for icetray/serialization.h (IceTray's copy of boost::serialization, un-vendored).  The reference's classes
// declare `template <class Archive> void serialize(Archive &, unsigned)` members and register themselves with macros;
// nothing is serialized in oracle/_ref, so the member templates are never instantiated and the macros expand to nothing.
#ifndef CLSIM_REF_SHIM_SERIALIZATION_H
#define CLSIM_REF_SHIM_SERIALIZATION_H
#include "icetray/I3TrayHeaders.h"
namespace icecube { namespace serialization {
class access;
template <class T> inline int make_nvp(const char *, T &) { return 0; }
template <class T> inline int make_nvp(const char *, const T &) { return 0; }
template <class Base, class Derived> inline Base &base_object(Derived &d) { return d; }
template <class Base, class Derived> inline const Base &base_object(const Derived &d) { return d; }
}} // namespace icecube::serialization
using icecube::serialization::base_object;
using icecube::serialization::make_nvp;
#define I3_SERIALIZABLE(T)
#define I3_SPLIT_SERIALIZABLE(T)
#define I3_CLASS_VERSION(T, V)
#define I3_SERIALIZATION_SPLIT_MEMBER() template <class Archive> void serialize(Archive &, unsigned) {}
#endif
