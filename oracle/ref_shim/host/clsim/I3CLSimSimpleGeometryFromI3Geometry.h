// Stand-in for public/clsim/I3CLSimSimpleGeometryFromI3Geometry.h (built from IceTray's I3Geometry, un-vendored): the name only.
#ifndef CLSIM_REF_SHIM_GEOMETRY_FROM_I3_H
#define CLSIM_REF_SHIM_GEOMETRY_FROM_I3_H
#include "icetray/I3TrayHeaders.h"
class I3CLSimSimpleGeometryFromI3Geometry {};
I3_POINTER_TYPEDEFS(I3CLSimSimpleGeometryFromI3Geometry);
#endif
