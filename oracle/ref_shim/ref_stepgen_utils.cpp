// oracle/_ref/libclsim_ref_stepgen.so: the inline samplers of the reference's step generator
// (private/clsim/I3CLSimLightSourceToStepConverterUtils.h: mwcRngRandomNumber_co/oc, gammaDistributedNumber,
// scatterDirectionByAngle, mwcRngInitState), compiled unmodified from where the header lies under /root/reference.
// tests/test_stepgen_oracle.py holds oracle/stepgen_oracle.py -- the checker of csrc/stepgen.cu -- against them,
// bit for bit.  Test infrastructure; nothing under clsim_b200/ links or loads this.
#include <cstdint>

#include "clsim/I3CLSimLightSourceToStepConverterUtils.h"   // -I $(REFERENCE)/private

namespace U = I3CLSimLightSourceToStepConverterUtils;

namespace {
// an I3RandomService that hands out a prepared list of 32-bit integers (for mwcRngInitState)
class ListService : public I3RandomService {
public:
    const uint32_t *values;
    size_t n, at = 0;
    ListService(const uint32_t *v, size_t count) : values(v), n(count) {}
    unsigned int Integer(unsigned int) override { return at < n ? values[at++] : 1u; }
    double Uniform(double) override { return 0.5; }
};
} // namespace

extern "C" {

double ref_mwc_co(uint64_t *state, uint32_t a) { return U::mwcRngRandomNumber_co(*state, a); }
double ref_mwc_oc(uint64_t *state, uint32_t a) { return U::mwcRngRandomNumber_oc(*state, a); }
double ref_gamma_distributed(double shape, uint64_t *state, uint32_t a) { return U::gammaDistributedNumber(shape, *state, a); }
void ref_scatter_direction_by_angle(double cosa, double sina, double *xyz, double random_value)
{
    U::scatterDirectionByAngle(cosa, sina, xyz[0], xyz[1], xyz[2], random_value);
}
// -> the state the reference would start a stream of multiplier `a` from, drawing from `values`; *used = integers consumed
uint64_t ref_mwc_init_state(const uint32_t *values, size_t n, uint32_t a, size_t *used)
{
    auto service = std::make_shared<ListService>(values, n);
    const uint64_t x = U::mwcRngInitState(service, a);
    *used = service->at;
    return x;
}

} // extern "C"
