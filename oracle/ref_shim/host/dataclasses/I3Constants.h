// Stand-in for dataclasses/I3Constants.h (un-vendored): the speed of light in the IceCube unit system (m / ns).
#ifndef CLSIM_REF_SHIM_I3CONSTANTS_H
#define CLSIM_REF_SHIM_I3CONSTANTS_H
#include "icetray/I3Units.h"
namespace I3Constants {
static const double c = 2.99792458e8 * I3Units::m / (I3Units::second);
}
#endif
