// test_models.h -- the small self-contained detector / ice / step models the C++ tests share (no GCD file,
// no ice tables on disk: the shapes of resources/scripts/benchmark.py:63-114 and
// python/MakeIceCubeMediumProperties.py:166-230 with synthetic numbers).
#ifndef CLSIM_TEST_MODELS_H_INCLUDED
#define CLSIM_TEST_MODELS_H_INCLUDED

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>

#include "I3CLSimStepToPhotonConverterCUDA.h"

static const double nm = 1e-9, deg = M_PI / 180.0;

// tests/test_hostcheck.py runs these programs against the CUDA sources compiled for the host (tests/hostcheck), which hold no
// fast kernel: with CLSIM_HOSTCHECK=1 in the environment the converters take the reference-order kernel and smaller bunches.
// On a GPU (the variable unset) nothing changes.
static inline bool test_on_host_check() { return std::getenv("CLSIM_HOSTCHECK") != nullptr; }
static inline bool test_native_math() { return !test_on_host_check(); }
static inline std::size_t test_sized(std::size_t on_gpu, std::size_t on_host) { return test_on_host_check() ? on_host : on_gpu; }

// 24-DOM ring (resources/scripts/benchmark.py:63-114)
static inline I3CLSimSimpleGeometryConstPtr make_ring_geometry(double oversize)
{
    const double radius = 120.0;
    const double dirs[8][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}};
    std::shared_ptr<I3CLSimSimpleGeometryUserConfigurable> g(new I3CLSimSimpleGeometryUserConfigurable(0.16510 * oversize, 24));
    std::size_t at = 0;
    for (int s = 0; s < 8; ++s) {
        const double len = std::sqrt(dirs[s][0] * dirs[s][0] + dirs[s][1] * dirs[s][1]);
        const double dz[3] = {radius, 0.0, -radius};
        for (int d = 0; d < 3; ++d, ++at) {
            g->SetStringID(at, s + 1);
            g->SetDomID(at, d + 1);
            g->SetPosX(at, dirs[s][0] / len * radius);
            g->SetPosY(at, dirs[s][1] / len * radius);
            g->SetPosZ(at, dz[d]);
            g->SetSubdetector(at, "Unknown");
        }
    }
    return g;
}

// A 12-layer ice in the shape MakeIceCubeMediumProperties.py:166-230 builds (values of the order of SpiceMie's)
static inline I3CLSimMediumPropertiesConstPtr make_medium(bool with_tilt_and_anisotropy)
{
    const uint32_t layers = 12;
    std::shared_ptr<I3CLSimMediumProperties> m(new I3CLSimMediumProperties(0.9216, layers, -60.0, 10.0, -870.0, 1940.0));
    m->SetForcedMinWlen(265 * nm);
    m->SetForcedMaxWlen(675 * nm);
    const double kappa = 1.08410680294, A = 6954.09033203, B = 6617.75439453, alpha = 0.898608505726;
    I3CLSimFunctionConstPtr phase(new I3CLSimFunctionRefIndexIceCube("phase"));
    I3CLSimFunctionConstPtr group(new I3CLSimFunctionRefIndexIceCube("group"));
    for (uint32_t l = 0; l < layers; ++l) {
        const double be400 = 0.020 + 0.004 * std::sin(0.9 * l), adust = 0.006 + 0.002 * std::cos(0.7 * l), dtau = 5.0 + 0.3 * l;
        const double g = 0.9;
        m->SetAbsorptionLength(l, I3CLSimFunctionConstPtr(new I3CLSimFunctionAbsLenIceCube(kappa, A, B, std::pow(400.0, kappa), 0.0, adust, dtau)));
        m->SetScatteringLength(l, I3CLSimFunctionConstPtr(new I3CLSimFunctionScatLenIceCube(alpha, be400 / (1.0 - g))));
        m->SetPhaseRefractiveIndex(l, phase);
        m->SetGroupRefractiveIndexOverride(l, group);
    }
    I3CLSimRandomValueConstPtr sl(new I3CLSimRandomValueSimplifiedLiu(0.9)), hg(new I3CLSimRandomValueHenyeyGreenstein(0.9));
    m->SetScatteringCosAngleDistribution(I3CLSimRandomValueConstPtr(new I3CLSimRandomValueMixed(0.45, sl, hg)));
    if (with_tilt_and_anisotropy) {
        std::vector<double> dist = {-500.0, -100.0, 0.0, 150.0, 400.0}, zc;
        std::vector<std::vector<double> > corr(dist.size());
        for (int k = 0; k < 20; ++k) zc.push_back(-70.0 + 7.5 * k);
        for (std::size_t i = 0; i < dist.size(); ++i)
            for (int k = 0; k < 20; ++k) corr[i].push_back(0.004 * dist[i] * std::cos(0.2 * k));
        m->SetIceTiltZShift(I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldIceTiltZShift(dist, zc, corr, 225.0 * deg)));
        m->SetDirectionalAbsorptionLengthCorrection(I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldAnisotropyAbsLenScaling(216.0 * deg, 0.04, -0.08)));
        // GetSpiceLeaAnisotropyTransforms.py:39-101: T^T diag(k1,k2,1/(k1 k2))^{+-1} T with T the rotation by the azimuth
        const double k1 = std::exp(0.04), k2 = std::exp(-0.08), kz = 1.0 / (k1 * k2), ca = std::cos(216.0 * deg), sa = std::sin(216.0 * deg);
        double pre[9], post[9];
        const double T[9] = {ca, sa, 0, -sa, ca, 0, 0, 0, 1};
        const double kpre[3] = {k1, k2, kz}, kpost[3] = {1 / k1, 1 / k2, 1 / kz};
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                pre[3 * r + c] = post[3 * r + c] = 0;
                for (int j = 0; j < 3; ++j) {
                    pre[3 * r + c] += T[3 * j + r] * kpre[j] * T[3 * j + c];
                    post[3 * r + c] += T[3 * j + r] * kpost[j] * T[3 * j + c];
                }
            }
        m->SetPreScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformMatrix(pre, true)));
        m->SetPostScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformMatrix(post, true)));
    }
    return m;
}

static inline I3CLSimFunctionConstPtr make_bias()
{
    // shape of GetIceCubeDOMAcceptance (43 values, 260 nm + 10 nm * i), numbers synthetic
    std::vector<double> v;
    for (int i = 0; i < 43; ++i) v.push_back(0.02 + 0.11 * std::exp(-0.5 * std::pow((i - 14) / 7.0, 2)));
    return I3CLSimFunctionConstPtr(new I3CLSimFunctionFromTable(260 * nm, 10 * nm, v));
}

// makeCherenkovWavelengthGenerator with a table bias (I3CLSimModuleHelper.cxx:224-256): spectrum on the bias grid
static inline I3CLSimRandomValueConstPtr make_generator(const I3CLSimFunctionConstPtr &bias, const I3CLSimMediumPropertiesConstPtr &medium)
{
    auto t = std::dynamic_pointer_cast<const I3CLSimFunctionFromTable>(bias);
    std::vector<double> y;
    for (std::size_t i = 0; i < t->GetNumEntries(); ++i) {
        const double w = t->GetEntryWavelength(i), n = medium->GetPhaseRefractiveIndex(0)->GetValue(w);
        y.push_back(t->GetEntryValue(i) * (2.0 * M_PI / 137.0) / (w * w) * (1.0 - 1.0 / (n * n)));
    }
    return I3CLSimRandomValueConstPtr(new I3CLSimRandomValueInterpolatedDistribution(t->GetFirstWavelength(), t->GetWavelengthStepping(), y));
}

static inline I3CLSimStepSeriesPtr make_steps(std::size_t n, uint32_t photons, uint32_t id, unsigned seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    I3CLSimStepSeriesPtr s(new I3CLSimStepSeries(n));
    for (std::size_t i = 0; i < n; ++i) {
        I3CLSimStep &st = (*s)[i];
        std::memset(&st, 0, sizeof(st));
        st.SetPosX(static_cast<float>(30.0 * u(rng)));
        st.SetPosY(static_cast<float>(30.0 * u(rng)));
        st.SetPosZ(static_cast<float>(30.0 * u(rng)));
        st.SetTime(static_cast<float>(100.0 + 50.0 * u(rng)));
        double x, y, z, r2;
        do { x = u(rng); y = u(rng); z = u(rng); r2 = x * x + y * y + z * z; } while (r2 > 1.0 || r2 < 1e-4);
        st.SetDir(x, y, z);
        st.SetLength(1.0f);
        st.SetBeta(1.0f);
        st.SetNumPhotons(photons);
        st.SetWeight(1.0f);
        st.SetID(id);
        st.SetSourceType(0);
    }
    return s;
}


#endif // CLSIM_TEST_MODELS_H_INCLUDED
