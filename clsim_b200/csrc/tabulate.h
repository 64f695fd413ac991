// tabulate.h -- internal seam between the engine (engine.cu) and the table-maker variant (tabulate.cu).
#ifndef CLSIMCU_TABULATE_H_INCLUDED
#define CLSIMCU_TABULATE_H_INCLUDED

#include <cstddef>
#include <string>

#include "../../include/clsimcuda.h"
#include "device_scene.h"

namespace clsimcu {

int engine_device(const clsimcu_engine *e);
size_t engine_max_items(const clsimcu_engine *e);
// copies the bunch to the device and launches the reference-order kernel in table mode on the engine's compute
// stream; returns an error text or an empty string
std::string engine_launch_tabulate(clsimcu_engine *e, const clsimcu_step *steps, size_t n, const TabulateArgs *d_tab);
// a copy ordered with the launches (and, with `wait`, a stream synchronisation)
std::string engine_copy_on_stream(clsimcu_engine *e, void *dst, const void *src, size_t bytes, bool to_device, bool wait);

} // namespace clsimcu

#endif
