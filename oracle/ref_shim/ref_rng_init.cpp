// ref_rng_init.cpp -- the reference's RNG seeding (private/opencl/mwcrng_init.h: init_MWC_RNG) compiled unmodified.
// TEST INFRASTRUCTURE (oracle/_ref/libclsim_ref_rng.so).  Reads a multiplier table in the reference's text format, draws
// the start states from an I3RandomService; here that service hands out, 32 bits at a time, the 64-bit values the caller
// supplies (the un-vendored I3GSLRandomService is replaced by splitmix64 in oracle and product; the RULE is what is pinned).
#include <cstdint>
#include <cstddef>
#include <limits>

#include "opencl/mwcrng_init.h"

namespace {
class SuppliedBits : public I3RandomService {
public:
    SuppliedBits(const uint64_t *v, size_t n) : v_(v), n_(n), at_(0) {}
    unsigned int Integer(unsigned int) override
    {
        if (at_ >= 2 * n_) throw std::runtime_error("out of supplied random values");
        const uint64_t r = v_[at_ / 2];
        const unsigned int out = (at_ % 2 == 0) ? static_cast<unsigned int>(r >> 32) : static_cast<unsigned int>(r);
        ++at_;
        return out;
    }
    double Uniform(double) override { throw std::runtime_error("not used"); }
    size_t used() const { return (at_ + 1) / 2; }
private:
    const uint64_t *v_;
    size_t n_, at_;
};
}

extern "C" int ref_init_mwc_rng(uint64_t *x, uint32_t *a, uint32_t n, const char *safeprimes_file, const uint64_t *values, size_t num_values, size_t *used)
{
    try {
        boost::shared_ptr<SuppliedBits> rng(new SuppliedBits(values, num_values));
        const int rc = init_MWC_RNG(x, a, n, rng, safeprimes_file);
        if (used) *used = rng->used();
        return rc;
    } catch (const std::exception &) {
        return -1;
    }
}
