// Stand-in for dataclasses/I3Vector.h: a std::vector that is a frame object and has a serialize() member template (the
// reference specialises it for its step and photon series).
#ifndef CLSIM_REF_SHIM_I3VECTOR_H
#define CLSIM_REF_SHIM_I3VECTOR_H
#include <vector>
#include "icetray/serialization.h"
template <class T> class I3Vector : public I3FrameObject, public std::vector<T> {
public:
    I3Vector() {}
    explicit I3Vector(std::size_t n) : std::vector<T>(n) {}
    I3Vector(std::size_t n, const T &v) : std::vector<T>(n, v) {}
    template <class Iterator> I3Vector(Iterator first, Iterator last) : std::vector<T>(first, last) {}
    template <class Archive> void serialize(Archive &ar, unsigned version);
};
#endif
