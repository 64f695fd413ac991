// Stand-in for public/clsim/I3CLSimLightSourceParameterization.h: not used by the parts of I3CLSimModuleHelper.cxx that are run.
#ifndef CLSIM_REF_SHIM_LIGHT_SOURCE_PARAMETERIZATION_H
#define CLSIM_REF_SHIM_LIGHT_SOURCE_PARAMETERIZATION_H
#endif
