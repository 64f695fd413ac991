// Stand-in for phys-services/I3RandomService.h (un-vendored): the interface the step-converter utilities call.
#ifndef CLSIM_REF_SHIM_I3RANDOMSERVICE_H
#define CLSIM_REF_SHIM_I3RANDOMSERVICE_H
#include "icetray/I3TrayHeaders.h"
class I3RandomService {
public:
    virtual ~I3RandomService() {}
    virtual unsigned int Integer(unsigned int imax) = 0;
    virtual double Uniform(double x = 1) = 0;
};
I3_POINTER_TYPEDEFS(I3RandomService);
#endif
