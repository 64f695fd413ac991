// Stand-in for public/clsim/I3CLSimStepToPhotonConverterOpenCL.h (needs an OpenCL runtime): the methods
// I3CLSimModuleHelper::initializeOpenCL calls, so that private/clsim/I3CLSimModuleHelper.cxx compiles unmodified.  Only the
// wavelength-generator factories of that file are run in oracle/_ref.
#ifndef CLSIM_REF_SHIM_CONVERTER_OPENCL_H
#define CLSIM_REF_SHIM_CONVERTER_OPENCL_H
#include <cstddef>
#include <cstdint>
#include <vector>
#include "icetray/I3TrayHeaders.h"
#include "phys-services/I3RandomService.h"
#include "clsim/I3CLSimMediumProperties.h"
#include "clsim/I3CLSimOpenCLDevice.h"
#include "clsim/I3CLSimSimpleGeometryFromI3Geometry.h"
class I3CLSimStepToPhotonConverterOpenCL {
public:
    I3CLSimStepToPhotonConverterOpenCL(I3RandomServicePtr, bool) {}
    void SetDevice(const I3CLSimOpenCLDevice &) {}
    void SetWlenGenerators(const std::vector<I3CLSimRandomValueConstPtr> &) {}
    void SetWlenBias(I3CLSimFunctionConstPtr) {}
    void SetMediumProperties(I3CLSimMediumPropertiesConstPtr) {}
    void SetGeometry(I3CLSimSimpleGeometryFromI3GeometryPtr) {}
    void SetEnableDoubleBuffering(bool) {}
    void SetDoublePrecision(bool) {}
    void SetStopDetectedPhotons(bool) {}
    void SetSaveAllPhotons(bool) {}
    void SetSaveAllPhotonsPrescale(double) {}
    void SetFixedNumberOfAbsorptionLengths(double) {}
    void SetDOMPancakeFactor(double) {}
    void SetPhotonHistoryEntries(uint32_t) {}
    void Compile() {}
    std::size_t GetMaxWorkgroupSize() const { return 1; }
    void SetWorkgroupSize(std::size_t) {}
    std::size_t GetWorkgroupSize() const { return 1; }
    void SetMaxNumWorkitems(std::size_t) {}
    void Initialize() {}
};
I3_POINTER_TYPEDEFS(I3CLSimStepToPhotonConverterOpenCL);
#endif
