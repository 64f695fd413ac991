// Stand-in for dataclasses/I3Direction.h.  IceCube stores a direction as the zenith and azimuth of where the particle
// COMES FROM; theta and phi are the polar angles of where it GOES: theta = pi - zenith, phi = azimuth + pi (mod 2 pi).
// (The convention is IceTray's, un-vendored; the wire test does not depend on it -- only the raw theta / phi setters are used.)
#ifndef CLSIM_REF_SHIM_I3DIRECTION_H
#define CLSIM_REF_SHIM_I3DIRECTION_H
#include <cmath>
#include "icetray/I3TrayHeaders.h"
class I3Direction {
public:
    I3Direction() : theta_(0), phi_(0) {}
    I3Direction(double x, double y, double z)
    {
        const double r = std::sqrt(x * x + y * y + z * z);
        theta_ = (r > 0) ? std::acos(z / r) : 0.;
        phi_ = std::atan2(y, x);
        if (phi_ < 0) phi_ += 2 * M_PI;
    }
    void SetThetaPhi(double theta, double phi) { theta_ = theta; phi_ = phi; }
    double CalcTheta() const { return theta_; }
    double CalcPhi() const { return phi_; }
    double GetZenith() const { return M_PI - theta_; }
    double GetAzimuth() const { return std::fmod(phi_ + M_PI, 2 * M_PI); }
    double GetX() const { return std::sin(theta_) * std::cos(phi_); }
    double GetY() const { return std::sin(theta_) * std::sin(phi_); }
    double GetZ() const { return std::cos(theta_); }
private:
    double theta_, phi_;
};
I3_POINTER_TYPEDEFS(I3Direction);
#endif
