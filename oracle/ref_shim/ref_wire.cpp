// ref_wire.cpp -- the reference's own step / photon records and their serialization code, compiled for the host.
// TEST INFRASTRUCTURE (oracle/_ref/libclsim_ref_wire.so).  Nothing under clsim_b200/ links this.
//
// public/clsim/I3CLSimStep.h and I3CLSimPhoton.h (the records, with the reference's own copy of the OpenCL platform types,
// public/clsim/fake_cl_platform.h, -DI3CLSIM_WITHOUT_OPENCL) and private/clsim/I3CLSimStep.cxx, I3CLSimPhoton.cxx (the
// serialize() members of the records and of their series) are compiled unmodified; the archives they write to are the
// stand-ins of oracle/ref_shim/host_wire/icetray/serialization.h (which says what is the reference's and what is restated).
//
// Pins: (T1, T2) the byte layout of a record filled through the reference's SETTERS against clsimcu_step / clsimcu_photon;
// (f1) the body of a serialized series -- which values, of which C++ types, in which order -- against clsim_b200/wire.py.
#include <cstdint>
#include <cstring>
#include <string>

#include "clsim/I3CLSimStep.h"
#include "clsim/I3CLSimPhoton.h"

namespace {
thread_local std::string g_error;
thread_local std::string g_bytes;
}

extern "C" {

const char *ref_wire_error() { return g_error.c_str(); }
uint32_t ref_wire_step_size() { return sizeof(I3CLSimStep); }
uint32_t ref_wire_photon_size() { return sizeof(I3CLSimPhoton); }
uint32_t ref_wire_step_version() { return i3clsimstep_version_; }
uint32_t ref_wire_photon_version() { return i3clsimphoton_version_; }

// One step through the reference's setters -> its 48 bytes.  f: x y z t theta phi length beta weight; u: num id sourceType dummy1 dummy2
void ref_wire_make_step(const float *f, const uint32_t *u, void *out)
{
    I3CLSimStep s;
    std::memset(&s, 0xEE, sizeof s);   // (a byte no setter writes would show)
    s.SetPosX(f[0]); s.SetPosY(f[1]); s.SetPosZ(f[2]); s.SetTime(f[3]);
    s.SetDirTheta(f[4]); s.SetDirPhi(f[5]); s.SetLength(f[6]); s.SetBeta(f[7]);
    s.SetWeight(f[8]);
    s.SetNumPhotons(u[0]); s.SetID(u[1]); s.SetSourceType(static_cast<uint8_t>(u[2])); s.SetDummy1(static_cast<uint8_t>(u[3]));
    s.SetDummy2(static_cast<uint16_t>(u[4]));
    std::memcpy(out, &s, sizeof s);
}

// One photon through the reference's setters -> its 80 bytes.  f: x y z t theta phi wavelength cherenkovDist weight startX startY
// startZ startT startTheta startPhi groupVelocity distInAbsLens; u: numScatters id stringID(int16) omID
void ref_wire_make_photon(const float *f, const int64_t *u, void *out)
{
    I3CLSimPhoton p;
    std::memset(&p, 0xEE, sizeof p);
    p.SetPosX(f[0]); p.SetPosY(f[1]); p.SetPosZ(f[2]); p.SetTime(f[3]);
    p.SetDirTheta(f[4]); p.SetDirPhi(f[5]); p.SetWavelength(f[6]); p.SetCherenkovDist(f[7]); p.SetWeight(f[8]);
    p.SetStartPosX(f[9]); p.SetStartPosY(f[10]); p.SetStartPosZ(f[11]); p.SetStartTime(f[12]);
    p.SetStartDirTheta(f[13]); p.SetStartDirPhi(f[14]); p.SetGroupVelocity(f[15]); p.SetDistInAbsLens(f[16]);
    p.SetNumScatters(static_cast<uint32_t>(u[0])); p.SetID(static_cast<uint32_t>(u[1])); p.SetStringID(static_cast<int16_t>(u[2]));
    p.SetOMID(static_cast<uint16_t>(u[3]));
    std::memcpy(out, &p, sizeof p);
}

// I3Vector<I3CLSimStep>::serialize(portable_binary_oarchive) of n records given as raw bytes -> the body.  Returns its length;
// *out points to it (valid until the next call on this thread).
uint64_t ref_wire_write_steps(const void *records, uint64_t n, const char **out)
{
    I3CLSimStepSeries series(n);
    if (n) std::memcpy(&series[0], records, n * sizeof(I3CLSimStep));
    portable_binary_oarchive ar;
    series.serialize(ar, 0);
    g_bytes.swap(ar.bytes);
    *out = g_bytes.data();
    return g_bytes.size();
}

uint64_t ref_wire_write_photons(const void *records, uint64_t n, const char **out)
{
    I3CLSimPhotonSeries series(n);
    if (n) std::memcpy(&series[0], records, n * sizeof(I3CLSimPhoton));
    portable_binary_oarchive ar;
    series.serialize(ar, 0);
    g_bytes.swap(ar.bytes);
    *out = g_bytes.data();
    return g_bytes.size();
}

// I3Vector<I3CLSimStep>::serialize(portable_binary_iarchive) of a body -> records (at most cap are copied out).  Returns the
// number of records, or -1 with ref_wire_error() set (log_fatal of the reference, or a short stream); *left = bytes not consumed.
int64_t ref_wire_read_steps(const char *body, uint64_t len, void *records, uint64_t cap, uint64_t *left)
{
    try {
        I3CLSimStepSeries series;
        portable_binary_iarchive ar(body, len);
        series.serialize(ar, 0);
        if (left) *left = ar.left();
        const uint64_t k = series.size() < cap ? series.size() : cap;
        if (k) std::memcpy(records, &series[0], k * sizeof(I3CLSimStep));
        return static_cast<int64_t>(series.size());
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

int64_t ref_wire_read_photons(const char *body, uint64_t len, void *records, uint64_t cap, uint64_t *left)
{
    try {
        I3CLSimPhotonSeries series;
        portable_binary_iarchive ar(body, len);
        series.serialize(ar, 0);
        if (left) *left = ar.left();
        const uint64_t k = series.size() < cap ? series.size() : cap;
        if (k) std::memcpy(records, &series[0], k * sizeof(I3CLSimPhoton));
        return static_cast<int64_t>(series.size());
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// the getters of a record: f and u as in ref_wire_make_step
void ref_wire_read_step_fields(const void *record, float *f, uint32_t *u)
{
    I3CLSimStep s;
    std::memcpy(&s, record, sizeof s);
    f[0] = s.GetPosX(); f[1] = s.GetPosY(); f[2] = s.GetPosZ(); f[3] = s.GetTime();
    f[4] = s.GetDirTheta(); f[5] = s.GetDirPhi(); f[6] = s.GetLength(); f[7] = s.GetBeta(); f[8] = s.GetWeight();
    u[0] = s.GetNumPhotons(); u[1] = s.GetID(); u[2] = s.GetSourceType(); u[3] = s.GetDummy1(); u[4] = s.GetDummy2();
}

} // extern "C"
