// Stand-in for icetray/I3Units.h (un-vendored).  The IceCube unit system: metre, nanosecond, GeV and radian are 1.
// Only the length and angle units are exercised by the classes pinned in oracle/_ref; the derived mechanical units
// (used by the ANTARES water models, which are compiled but not evaluated) follow the CLHEP construction.
#ifndef CLSIM_REF_SHIM_I3UNITS_H
#define CLSIM_REF_SHIM_I3UNITS_H
namespace I3Units {
static const double meter = 1.0, m = meter, centimeter = 1e-2 * meter, cm = centimeter, millimeter = 1e-3 * meter, mm = millimeter,
                    micrometer = 1e-6 * meter, nanometer = 1e-9 * meter, kilometer = 1e3 * meter, km = kilometer;
static const double meter2 = meter * meter, m2 = meter2, meter3 = meter * meter * meter, m3 = meter3, cm3 = cm * cm * cm;
static const double radian = 1.0, rad = radian, degree = (3.14159265358979323846 / 180.0) * radian, deg = degree;
static const double nanosecond = 1.0, ns = nanosecond, second = 1e9 * nanosecond, s = second;
static const double eV = 1e-9, GeV = 1.0, e_SI = 1.60217646e-19, joule = eV / e_SI;
static const double kilogram = joule * second * second / (meter * meter), kg = kilogram, gram = 1e-3 * kilogram, g = gram;
static const double newton = joule / meter, pascal = newton / m2, bar = 100000 * pascal;
static const double kelvin = 1.0;
static const double perCent = 0.01, perThousand = 0.001, perMillion = 0.000001;
} // namespace I3Units
#endif
