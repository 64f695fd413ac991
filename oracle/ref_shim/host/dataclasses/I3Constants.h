// Stand-in for dataclasses/I3Constants.h (un-vendored): the IceCube unit system has metres and nanoseconds as 1.
#ifndef CLSIM_REF_SHIM_I3CONSTANTS_H
#define CLSIM_REF_SHIM_I3CONSTANTS_H
namespace I3Units {
static const double meter = 1.0, m = meter, millimeter = 1e-3 * meter, mm = millimeter;
}
#endif
