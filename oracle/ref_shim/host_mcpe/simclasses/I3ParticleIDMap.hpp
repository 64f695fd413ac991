// Stand-in for simclasses/I3ParticleIDMap.hpp: per DOM, which particle made which photo-electrons (filled by hit merging only).
#ifndef CLSIM_REF_SHIM_I3PARTICLEIDMAP_HPP
#define CLSIM_REF_SHIM_I3PARTICLEIDMAP_HPP
#include <map>
#include <vector>
#include "dataclasses/I3Map.h"
#include "dataclasses/physics/I3ParticleID.h"
typedef std::map<I3ParticleID, std::vector<uint32_t> > ParticlePulseIndexMap;
typedef I3Map<OMKey, ParticlePulseIndexMap> I3ParticleIDMap;
I3_POINTER_TYPEDEFS(I3ParticleIDMap);
#endif
