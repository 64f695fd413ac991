// Stand-in for icetray/I3FrameObject.h: the base of everything that can sit in an IceTray frame.
#ifndef CLSIM_REF_SHIM_I3FRAMEOBJECT_H
#define CLSIM_REF_SHIM_I3FRAMEOBJECT_H
#include "icetray/I3TrayHeaders.h"
class I3FrameObject {
public:
    virtual ~I3FrameObject() {}
};
#endif
