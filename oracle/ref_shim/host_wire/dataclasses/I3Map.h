// Stand-in for dataclasses/I3Map.h (+ the two key types the reference's series maps are keyed by).
#ifndef CLSIM_REF_SHIM_I3MAP_H
#define CLSIM_REF_SHIM_I3MAP_H
#include <map>
#include "icetray/serialization.h"
template <class K, class V> class I3Map : public I3FrameObject, public std::map<K, V> {};
struct OMKey {
    int string;
    unsigned om;
    bool operator<(const OMKey &o) const { return string != o.string ? string < o.string : om < o.om; }
};
struct ModuleKey {
    int string;
    unsigned om;
    bool operator<(const ModuleKey &o) const { return string != o.string ? string < o.string : om < o.om; }
};
#endif
