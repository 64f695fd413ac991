// The fast kernel is not part of the host check (inline PTX, warp collectives, shared memory): its entry points exist so that
// engine.cu links, and say so when called.  TEST INFRASTRUCTURE (tests/hostcheck).
#include <cuda_runtime.h>

#include "device_scene.h"

thread_local HostcheckIdx threadIdx, blockIdx, blockDim, gridDim;

namespace clsimcu {
int launch_fast_kernel(const DevScene &, const LaunchArgs &, int, void *) { return -3; }
bool fast_kernel_supports(const DevScene &, const char **why)
{
    if (why) *why = "the host check build holds no fast kernel";
    return false;
}
bool fast_kernel_smem_is_the_problem(const DevScene &) { return false; }
void fast_kernel_geometry(int, int *grid_blocks, int *threads_per_block)
{
    *grid_blocks = 1;
    *threads_per_block = 32;
}
} // namespace clsimcu
