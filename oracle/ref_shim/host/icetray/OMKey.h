// Stand-in for icetray/OMKey.h (+ ModuleKey): (string, om[, pmt]) keys ordered like IceTray's (string, then om, then pmt).
#ifndef CLSIM_REF_SHIM_OMKEY_H
#define CLSIM_REF_SHIM_OMKEY_H
#include <ostream>
class OMKey {
public:
    OMKey() : s_(0), o_(0), p_(0) {}
    OMKey(int s, unsigned o, unsigned char p = 0) : s_(s), o_(o), p_(p) {}
    int GetString() const { return s_; }
    unsigned GetOM() const { return o_; }
    unsigned char GetPMT() const { return p_; }
    bool operator<(const OMKey &k) const { return s_ != k.s_ ? s_ < k.s_ : (o_ != k.o_ ? o_ < k.o_ : p_ < k.p_); }
    bool operator==(const OMKey &k) const { return s_ == k.s_ && o_ == k.o_ && p_ == k.p_; }
private:
    int s_;
    unsigned o_;
    unsigned char p_;
};
inline std::ostream &operator<<(std::ostream &os, const OMKey &k) { return os << "OMKey(" << k.GetString() << "," << k.GetOM() << "," << int(k.GetPMT()) << ")"; }
class ModuleKey {
public:
    ModuleKey() : s_(0), o_(0) {}
    ModuleKey(int s, unsigned o) : s_(s), o_(o) {}
    int GetString() const { return s_; }
    unsigned GetOM() const { return o_; }
    bool operator<(const ModuleKey &k) const { return s_ != k.s_ ? s_ < k.s_ : o_ < k.o_; }
private:
    int s_;
    unsigned o_;
};
#endif
