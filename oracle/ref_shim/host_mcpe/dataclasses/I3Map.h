// Stand-in for dataclasses/I3Map.h / I3Vector.h: standard containers that are frame objects.
#ifndef CLSIM_REF_SHIM_MCPE_I3MAP_H
#define CLSIM_REF_SHIM_MCPE_I3MAP_H
#include <map>
#include <string>
#include <vector>
#include "icetray/I3FrameObject.h"
#include "icetray/OMKey.h"
template <class K, class V> class I3Map : public I3FrameObject, public std::map<K, V> {};
template <class T> class I3Vector : public I3FrameObject, public std::vector<T> {};
typedef I3Map<std::string, double> I3MapStringDouble;
I3_POINTER_TYPEDEFS(I3MapStringDouble);
#endif
