// Stand-in for public/clsim/I3CLSimStep.h (it is an I3FrameObject with boost serialization): only the name is needed by
// the inline utilities compiled in oracle/ref_shim/ref_stepgen_utils.cpp.
#ifndef CLSIM_REF_SHIM_I3CLSIMSTEP_H
#define CLSIM_REF_SHIM_I3CLSIMSTEP_H
struct I3CLSimStep;
#endif
