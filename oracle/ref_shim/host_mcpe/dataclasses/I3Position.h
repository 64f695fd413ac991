// Stand-in for dataclasses/I3Position.h and I3Direction.h: cartesian vectors.  An I3Direction is the unit vector a particle
// TRAVELS along (GetX/GetY/GetZ); IceTray's zenith / azimuth (where it comes from) are not used on this path.
#ifndef CLSIM_REF_SHIM_MCPE_I3POSITION_H
#define CLSIM_REF_SHIM_MCPE_I3POSITION_H
#include <cmath>
#include "icetray/I3TrayHeaders.h"
class I3Direction {
public:
    I3Direction() : x_(0), y_(0), z_(1) {}
    I3Direction(double x, double y, double z)
    {
        const double r = std::sqrt(x * x + y * y + z * z);
        x_ = x / r; y_ = y / r; z_ = z / r;
    }
    void SetThetaPhi(double theta, double phi) { x_ = std::sin(theta) * std::cos(phi); y_ = std::sin(theta) * std::sin(phi); z_ = std::cos(theta); }
    double GetX() const { return x_; }
    double GetY() const { return y_; }
    double GetZ() const { return z_; }
private:
    double x_, y_, z_;
};
class I3Position {
public:
    I3Position() : x_(0), y_(0), z_(0) {}
    I3Position(double x, double y, double z) : x_(x), y_(y), z_(z) {}
    double GetX() const { return x_; }
    double GetY() const { return y_; }
    double GetZ() const { return z_; }
    double Magnitude() const { return std::sqrt(x_ * x_ + y_ * y_ + z_ * z_); }
    I3Position operator-(const I3Position &o) const { return I3Position(x_ - o.x_, y_ - o.y_, z_ - o.z_); }
    double operator*(const I3Direction &d) const { return x_ * d.GetX() + y_ * d.GetY() + z_ * d.GetZ(); }
private:
    double x_, y_, z_;
};
#endif
