// Stand-in for dataclasses/physics/I3ParticleID.h.
#ifndef CLSIM_REF_SHIM_I3PARTICLEID_H
#define CLSIM_REF_SHIM_I3PARTICLEID_H
#include <cstdint>
struct I3ParticleID {
    uint64_t majorID;
    int32_t minorID;
    I3ParticleID() : majorID(0), minorID(0) {}
    I3ParticleID(uint64_t M, int32_t m) : majorID(M), minorID(m) {}
    bool operator<(const I3ParticleID &o) const { return majorID != o.majorID ? majorID < o.majorID : minorID < o.minorID; }
};
#endif
