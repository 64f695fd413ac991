// Stand-in for sim-services/MCPEMCPulseTools.hpp (un-vendored): hit merging is off by default and NOT part of what is pinned;
// the function exists so that the converter's source compiles, and refuses to run.
#ifndef CLSIM_REF_SHIM_MCPEMCPULSETOOLS_HPP
#define CLSIM_REF_SHIM_MCPEMCPULSETOOLS_HPP
#include "simclasses/I3MCPE.h"
#include "simclasses/I3ParticleIDMap.hpp"
namespace MCHitMerging {
inline ParticlePulseIndexMap extractPIDInfoandMerge(I3MCPESeries &, const OMKey &)
{
    ref_shim::fatal("MCHitMerging (sim-services) is not vendored by the reference: MergeHits cannot be pinned");
    return ParticlePulseIndexMap();
}
} // namespace MCHitMerging
#endif
