// Stand-in for public/clsim/function/I3CLSimFunction.h (serializable IceTray class): only the name is needed by the
// declarations in I3CLSimLightSourceToStepConverterUtils.h.
#ifndef CLSIM_REF_SHIM_I3CLSIMFUNCTION_H
#define CLSIM_REF_SHIM_I3CLSIMFUNCTION_H
#include "icetray/I3TrayHeaders.h"
class I3CLSimFunction;
I3_POINTER_TYPEDEFS(I3CLSimFunction);
#endif
