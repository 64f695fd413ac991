// Stand-in for dataclasses/physics/I3Particle.h (un-vendored): the step-converter utilities only name the type.
#ifndef CLSIM_REF_SHIM_I3PARTICLE_H
#define CLSIM_REF_SHIM_I3PARTICLE_H
#include "icetray/I3TrayHeaders.h"
class I3Particle {};
#endif
