// Stand-in for dataclasses/I3Position.h: three cartesian coordinates.
#ifndef CLSIM_REF_SHIM_I3POSITION_H
#define CLSIM_REF_SHIM_I3POSITION_H
#include "icetray/I3TrayHeaders.h"
class I3Position {
public:
    I3Position() : x_(0), y_(0), z_(0) {}
    I3Position(double x, double y, double z) : x_(x), y_(y), z_(z) {}
    double GetX() const { return x_; }
    double GetY() const { return y_; }
    double GetZ() const { return z_; }
private:
    double x_, y_, z_;
};
I3_POINTER_TYPEDEFS(I3Position);
#endif
