// Stand-in for public/clsim/I3CLSimOpenCLDevice.h (needs an OpenCL runtime): what initializeOpenCL asks a device.
#ifndef CLSIM_REF_SHIM_OPENCL_DEVICE_H
#define CLSIM_REF_SHIM_OPENCL_DEVICE_H
#include <cstdint>
#include <string>
class I3CLSimOpenCLDevice {
public:
    bool GetUseNativeMath() const { return false; }
    uint32_t GetApproximateNumberOfWorkItems() const { return 1; }
    std::string GetPlatformName() const { return std::string(); }
    std::string GetDeviceName() const { return std::string(); }
    bool IsGPU() const { return false; }
};
#endif
