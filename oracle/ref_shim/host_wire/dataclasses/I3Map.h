// Stand-in for dataclasses/I3Map.h (the key types the reference's series maps are keyed by: icetray/OMKey.h).
#ifndef CLSIM_REF_SHIM_I3MAP_H
#define CLSIM_REF_SHIM_I3MAP_H
#include <map>
#include "icetray/serialization.h"
#include "icetray/OMKey.h"
template <class K, class V> class I3Map : public I3FrameObject, public std::map<K, V> {};
#endif
