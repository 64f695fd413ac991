// Stand-in for icetray/I3ConditionalModule.h: a module that always runs.
#ifndef CLSIM_REF_SHIM_I3CONDITIONALMODULE_H
#define CLSIM_REF_SHIM_I3CONDITIONALMODULE_H
#include "icetray/I3Module.h"
class I3ConditionalModule : public I3Module {
public:
    explicit I3ConditionalModule(const I3Context &context) : I3Module(context) {}
};
#endif
