// Stand-in for dataclasses/status/I3DetectorStatus.h: the PMT high voltage per DOM.
#ifndef CLSIM_REF_SHIM_I3DETECTORSTATUS_H
#define CLSIM_REF_SHIM_I3DETECTORSTATUS_H
#include <map>
#include "icetray/I3FrameObject.h"
#include "icetray/OMKey.h"
struct I3DOMStatus {
    double pmtHV;
    I3DOMStatus() : pmtHV(0) {}
};
class I3DetectorStatus : public I3FrameObject {
public:
    std::map<OMKey, I3DOMStatus> domStatus;
};
I3_POINTER_TYPEDEFS(I3DetectorStatus);
#endif
