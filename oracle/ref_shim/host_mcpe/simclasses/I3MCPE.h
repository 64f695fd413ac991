// Stand-in for simclasses/I3MCPE.h: one photo-electron (who made it, how many, when).
#ifndef CLSIM_REF_SHIM_I3MCPE_H
#define CLSIM_REF_SHIM_I3MCPE_H
#include <vector>
#include "dataclasses/I3Map.h"
#include "dataclasses/physics/I3ParticleID.h"
struct I3MCPE {
    I3ParticleID ID;
    uint32_t npe;
    double time;
    I3MCPE() : npe(0), time(0) {}
    I3MCPE(const I3ParticleID &id, uint32_t n, double t) : ID(id), npe(n), time(t) {}
};
typedef std::vector<I3MCPE> I3MCPESeries;
typedef I3Map<OMKey, I3MCPESeries> I3MCPESeriesMap;
I3_POINTER_TYPEDEFS(I3MCPESeriesMap);
#endif
