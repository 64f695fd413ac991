// stepgen.h -- internal seam between the engine (engine.cu) and the step generator (stepgen.cu).
#ifndef CLSIMCU_STEPGEN_H_INCLUDED
#define CLSIMCU_STEPGEN_H_INCLUDED

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/clsimcuda.h"

namespace clsimcu {

constexpr size_t kMaxSourcesPerBunch = 65536;

struct StepGenLaunch {
    const clsimcu_step_source *sources;   // device
    const uint64_t *first_step;           // device, num_sources + 1 entries
    uint32_t num_sources;
    uint64_t total;
    clsimcu_step *out;                    // device
};

int stepgen_device(const clsimcu_step_generator *g);
// one kernel launch on `stream`; throws std::runtime_error on a CUDA error.  Launches on one generator must be
// ordered by the caller (its MWC states advance).
void stepgen_enqueue(clsimcu_step_generator *g, const StepGenLaunch &l, cudaStream_t stream);
std::string stepgen_layout(const clsimcu_step_source *sources, size_t n, std::vector<uint64_t> &first_step, uint64_t *photons);

} // namespace clsimcu

#endif
