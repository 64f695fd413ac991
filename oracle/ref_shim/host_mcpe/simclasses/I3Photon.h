// Stand-in for simclasses/I3Photon.h and I3CompressedPhoton.h: a detected photon as clsim's client module fills it
// (private/clsim/I3CLSimClientModule.cxx:330-360): position, direction of travel, time, wavelength, weight, group velocity.
#ifndef CLSIM_REF_SHIM_I3PHOTON_H
#define CLSIM_REF_SHIM_I3PHOTON_H
#include "dataclasses/I3Map.h"
#include "dataclasses/I3Position.h"
#include "dataclasses/physics/I3ParticleID.h"
class I3CompressedPhoton {
public:
    I3CompressedPhoton() : time_(0), weight_(0), wavelength_(0), groupVelocity_(0) {}
    double GetTime() const { return time_; }
    double GetWeight() const { return weight_; }
    double GetWavelength() const { return wavelength_; }
    double GetGroupVelocity() const { return groupVelocity_; }
    const I3Position &GetPos() const { return pos_; }
    const I3Direction &GetDir() const { return dir_; }
    I3ParticleID GetParticleID() const { return id_; }
    void SetTime(double v) { time_ = v; }
    void SetWeight(double v) { weight_ = v; }
    void SetWavelength(double v) { wavelength_ = v; }
    void SetGroupVelocity(double v) { groupVelocity_ = v; }
    void SetPos(const I3Position &p) { pos_ = p; }
    void SetDir(const I3Direction &d) { dir_ = d; }
    void SetParticleID(const I3ParticleID &id) { id_ = id; }
private:
    double time_, weight_, wavelength_, groupVelocity_;
    I3Position pos_;
    I3Direction dir_;
    I3ParticleID id_;
};
class I3Photon : public I3CompressedPhoton {
public:
    I3Photon() : numScattered_(0) {}
    const I3Position &GetStartPos() const { return startPos_; }
    uint32_t GetNumScattered() const { return numScattered_; }
    void SetStartPos(const I3Position &p) { startPos_ = p; }
    void SetNumScattered(uint32_t n) { numScattered_ = n; }
private:
    I3Position startPos_;
    uint32_t numScattered_;
};
typedef I3Vector<I3Photon> I3PhotonSeries;
typedef I3Map<ModuleKey, I3PhotonSeries> I3PhotonSeriesMap;
typedef I3Vector<I3CompressedPhoton> I3CompressedPhotonSeries;
typedef I3Map<ModuleKey, I3CompressedPhotonSeries> I3CompressedPhotonSeriesMap;
I3_POINTER_TYPEDEFS(I3PhotonSeriesMap);
I3_POINTER_TYPEDEFS(I3CompressedPhotonSeriesMap);
#endif
