// Stand-in for dataclasses/geometry/{I3Geometry,I3OMGeo,I3ModuleGeo}.h: where a PMT / a module sits and where it looks.
#ifndef CLSIM_REF_SHIM_I3GEOMETRY_H
#define CLSIM_REF_SHIM_I3GEOMETRY_H
#include "dataclasses/I3Map.h"
#include "dataclasses/I3Position.h"
struct I3OMGeo {
    I3Position position;
    I3Direction direction;
    I3Direction GetDirection() const { return direction; }
};
class I3ModuleGeo {
public:
    I3ModuleGeo() : radius_(0) {}
    I3ModuleGeo(const I3Position &p, const I3Direction &d, double r) : pos_(p), dir_(d), radius_(r) {}
    const I3Position &GetPos() const { return pos_; }
    I3Direction GetDir() const { return dir_; }
    double GetRadius() const { return radius_; }
private:
    I3Position pos_;
    I3Direction dir_;
    double radius_;
};
typedef I3Map<OMKey, I3OMGeo> I3OMGeoMap;
typedef I3Map<ModuleKey, I3ModuleGeo> I3ModuleGeoMap;
I3_POINTER_TYPEDEFS(I3OMGeoMap);
I3_POINTER_TYPEDEFS(I3ModuleGeoMap);
#endif
