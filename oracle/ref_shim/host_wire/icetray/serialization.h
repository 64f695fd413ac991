// Stand-in for icetray/serialization.h with WORKING archives, for oracle/_ref/libclsim_ref_wire.so only: the reference's
// I3CLSimStep.cxx / I3CLSimPhoton.cxx are compiled unmodified and their serialize() members are run against these classes.
// TEST INFRASTRUCTURE.
//
// What is real here and what is not: the SEQUENCE of what gets written (class version, count, blob, the order and the
// C++ types of the fields) is the reference's code.  The ENCODING of one integer / float / raw block is this header's
// restatement of the archive IceTray uses (icecube::archive::portable_binary_[io]archive = the "eos portable archive"
// on boost.serialization, un-vendored): an integer is one signed size byte followed by that many bytes of the value,
// least significant first (zero: the single byte 0); a float travels as the integer of its IEEE-754 bits; a binary
// object is copied verbatim; a base-class sub-object is its tracking flag and class version, here two zero bytes for the
// I3FrameObject base (stated as an assumption in clsim_b200/wire.py as well).
#ifndef CLSIM_REF_SHIM_WIRE_SERIALIZATION_H
#define CLSIM_REF_SHIM_WIRE_SERIALIZATION_H
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "icetray/I3TrayHeaders.h"
#include "icetray/I3FrameObject.h"

namespace icecube {
namespace serialization {
class access {
public:
    template <class Archive, class T> static void serialize(Archive &ar, T &t, unsigned version) { t.serialize(ar, version); }
    template <class Archive, class T> static void save(Archive &ar, const T &t, unsigned version) { t.save(ar, version); }
    template <class Archive, class T> static void load(Archive &ar, T &t, unsigned version) { t.load(ar, version); }
};
template <class T> inline T &make_nvp(const char *, T &v) { return v; }
template <class T> inline const T &make_nvp(const char *, const T &v) { return v; }
struct binary_object {
    void *data;
    std::size_t size;
};
inline binary_object make_binary_object(void *p, std::size_t n) { return binary_object{p, n}; }
inline binary_object make_binary_object(const void *p, std::size_t n) { return binary_object{const_cast<void *>(p), n}; }
template <class Base> struct base_ref {
    Base *base;
};
template <class Base, class Derived> inline base_ref<Base> base_object(Derived &d) { return base_ref<Base>{static_cast<Base *>(&d)}; }
} // namespace serialization

namespace archive {
class portable_binary_oarchive {
public:
    std::string bytes;
    template <class T> portable_binary_oarchive &operator<<(const T &v) { put(v); return *this; }
    template <class T> portable_binary_oarchive &operator&(const T &v) { put(v); return *this; }

private:
    template <class T> typename std::enable_if<std::is_integral<T>::value>::type put(const T &v)
    {
        typedef typename std::make_unsigned<T>::type U;
        if (v == 0) { bytes.push_back(0); return; }
        const bool negative = std::is_signed<T>::value && v < 0;
        U mag = negative ? static_cast<U>(~static_cast<U>(v)) + 1u : static_cast<U>(v);   // (eos stores the two's complement of negatives;
        U raw = static_cast<U>(v);                                                        //  only unsigned values occur on this path)
        int n = 0;
        for (U t = mag; t; t >>= 8) ++n;
        bytes.push_back(static_cast<char>(negative ? -n : n));
        for (int i = 0; i < n; ++i) bytes.push_back(static_cast<char>((raw >> (8 * i)) & 0xff));
    }
    void put(const float &v)
    {
        uint32_t bits;
        std::memcpy(&bits, &v, 4);
        put(bits);
    }
    void put(const serialization::binary_object &b) { bytes.append(static_cast<const char *>(b.data), b.size); }
    void put(const serialization::base_ref<I3FrameObject> &) { bytes.append(2, '\0'); }
};

class portable_binary_iarchive {
public:
    portable_binary_iarchive(const char *p, std::size_t n) : p_(p), end_(p + n) {}
    template <class T> portable_binary_iarchive &operator>>(T &v) { get(v); return *this; }
    template <class T> portable_binary_iarchive &operator>>(const T &v) { get(const_cast<T &>(v)); return *this; }   // (binary_object, base_ref: rvalues)
    template <class T> portable_binary_iarchive &operator&(T &v) { get(v); return *this; }
    std::size_t left() const { return static_cast<std::size_t>(end_ - p_); }

private:
    const char *p_, *end_;
    void need(std::size_t n) const
    {
        if (left() < n) throw std::runtime_error("portable_binary_iarchive: input stream error");
    }
    template <class T> typename std::enable_if<std::is_integral<T>::value>::type get(T &v)
    {
        need(1);
        const int n = static_cast<signed char>(*p_++);
        if (n == 0) { v = 0; return; }
        if (n < 0 && !std::is_signed<T>::value) throw std::runtime_error("portable_binary_iarchive: negative value for an unsigned field");
        const int len = n < 0 ? -n : n;
        if (static_cast<std::size_t>(len) > sizeof(T)) throw std::runtime_error("portable_binary_iarchive: integer does not fit the field");
        need(len);
        typedef typename std::make_unsigned<T>::type U;
        U raw = n < 0 ? static_cast<U>(~static_cast<U>(0)) : 0;
        for (int i = 0; i < len; ++i) {
            raw &= ~(static_cast<U>(0xff) << (8 * i));
            raw |= static_cast<U>(static_cast<unsigned char>(*p_++)) << (8 * i);
        }
        v = static_cast<T>(raw);
    }
    void get(float &v)
    {
        uint32_t bits;
        get(bits);
        std::memcpy(&v, &bits, 4);
    }
    void get(serialization::binary_object &b)
    {
        need(b.size);
        std::memcpy(b.data, p_, b.size);
        p_ += b.size;
    }
    void get(serialization::base_ref<I3FrameObject> &)
    {
        need(2);
        p_ += 2;
    }
};
} // namespace archive
} // namespace icecube

using icecube::archive::portable_binary_iarchive;
using icecube::archive::portable_binary_oarchive;
using icecube::serialization::base_object;
using icecube::serialization::make_nvp;

#define I3_SERIALIZABLE(T)
#define I3_SPLIT_SERIALIZABLE(T)
#define I3_CLASS_VERSION(T, V)
#define I3_SERIALIZATION_SPLIT_MEMBER() template <class Archive> void serialize(Archive &, unsigned) {}
#endif
