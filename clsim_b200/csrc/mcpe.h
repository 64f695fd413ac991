// mcpe.h -- internal seam between the engine (engine.cu) and the photon -> MCPE conversion (mcpe.cu).
#ifndef CLSIMCU_MCPE_H_INCLUDED
#define CLSIMCU_MCPE_H_INCLUDED

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/clsimcuda.h"

namespace clsimcu {

// engine.cu: leaves `msg` in clsimcu_last_error() and returns `code`
int report_error(int code, const std::string &msg);
// engine.cu: where the safe-prime multiplier table is memoised
std::string prime_cache_file();

// device counters of one conversion: [0] survivors, then what the reference treats as fatal
enum McpeCounter { kMcpeSurvivors = 0, kMcpeNegativeWeight, kMcpeProbabilityAboveOne, kMcpeBadPosition, kMcpeUnknownDom, kMcpeCounters = 8 };

struct McpeLaunch {
    const clsimcu_photon *photons;   // device
    const uint32_t *count;           // device: number of photons (the propagation kernel's hit counter), or NULL
    uint32_t max_count;              // cap on *count (capacity of `photons`), or the count itself when count == NULL
    const float *uniforms;           // device: explicit draws, or NULL for the converter's MWC streams
    clsimcu_mcpe *out;               // device
    uint32_t cap;
    uint32_t *counters;              // device, kMcpeCounters words, zeroed by the caller
};

int mcpe_device(const clsimcu_mcpe_converter *c);
// one kernel launch on `stream`; throws std::runtime_error on a CUDA error.  Launches on one converter must be
// ordered by the caller (its MWC states advance).
void mcpe_enqueue(clsimcu_mcpe_converter *c, const McpeLaunch &l, cudaStream_t stream);
// the reference's fatal conditions, from the counters of a finished conversion; empty when there was none
std::string mcpe_error_text(const clsimcu_mcpe_converter *c, const uint32_t *counters);

} // namespace clsimcu

#endif
