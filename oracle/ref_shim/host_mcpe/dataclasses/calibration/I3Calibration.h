// Stand-in for dataclasses/calibration/I3Calibration.h: the two numbers of a DOM's calibration the converter reads.
#ifndef CLSIM_REF_SHIM_I3CALIBRATION_H
#define CLSIM_REF_SHIM_I3CALIBRATION_H
#include <cmath>
#include <map>
#include "icetray/I3FrameObject.h"
#include "icetray/OMKey.h"
struct SPEChargeDistribution {
    double compensation_factor;
    SPEChargeDistribution() : compensation_factor(NAN) {}
};
class I3DOMCalibration {
public:
    I3DOMCalibration() : eff_(NAN) {}
    double GetRelativeDomEff() const { return eff_; }
    void SetRelativeDomEff(double v) { eff_ = v; }
    SPEChargeDistribution GetCombinedSPEChargeDistribution() const { return spe_; }
    void SetCompensationFactor(double v) { spe_.compensation_factor = v; }
private:
    double eff_;
    SPEChargeDistribution spe_;
};
class I3Calibration : public I3FrameObject {
public:
    std::map<OMKey, I3DOMCalibration> domCal;
};
I3_POINTER_TYPEDEFS(I3Calibration);
#endif
