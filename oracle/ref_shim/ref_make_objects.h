// ref_make_objects.h -- the reference's description objects (medium, wavelength generators, bias) put together from the oracle's
// plain parameter structs, the way the reference's Python puts them together.  TEST INFRASTRUCTURE (oracle/_ref): shared by
// ref_medium.cpp (text generators) and ref_icetray_mode.cpp (the product's converter class compiled against the reference's headers).
// Uses the reference's class HEADERS only; include after them.
#ifndef CLSIM_REF_SHIM_MAKE_OBJECTS_H
#define CLSIM_REF_SHIM_MAKE_OBJECTS_H
#include <algorithm>
#include <stdexcept>
#include <vector>

#include "../clsim_oracle.h"
#include "clsim/I3CLSimMediumProperties.h"
#include "clsim/function/I3CLSimFunctionAbsLenIceCube.h"
#include "clsim/function/I3CLSimFunctionConstant.h"
#include "clsim/function/I3CLSimFunctionFromTable.h"
#include "clsim/function/I3CLSimFunctionRefIndexIceCube.h"
#include "clsim/function/I3CLSimFunctionScatLenIceCube.h"
#include "clsim/function/I3CLSimScalarFieldAnisotropyAbsLenScaling.h"
#include "clsim/function/I3CLSimScalarFieldConstant.h"
#include "clsim/function/I3CLSimScalarFieldIceTiltZShift.h"
#include "clsim/function/I3CLSimVectorTransformConstant.h"
#include "clsim/function/I3CLSimVectorTransformMatrix.h"
#include "clsim/random_value/I3CLSimRandomValueConstant.h"
#include "clsim/random_value/I3CLSimRandomValueHenyeyGreenstein.h"
#include "clsim/random_value/I3CLSimRandomValueInterpolatedDistribution.h"
#include "clsim/random_value/I3CLSimRandomValueMixed.h"
#include "clsim/random_value/I3CLSimRandomValueSimplifiedLiu.h"
#include "clsim/random_value/I3CLSimRandomValueWlenCherenkovNoDispersion.h"
#include "icetray/I3Units.h"

namespace {

// python/MakeIceCubeMediumProperties.py:185-244 with the arguments the oracle's parameter struct carries
I3CLSimMediumPropertiesPtr make_medium(const oracle_medium &m, const double *tilt_z)
{
    // rock and air levels are not part of the generated text (the step generators read them); IceCube's where the layers
    // fit between them, else the layer range itself, as the class defaults have it (I3CLSimMediumProperties.cxx:43-48)
    const double top = m.layers_zstart + m.num_layers * m.layers_height;
    I3CLSimMediumPropertiesPtr med(new I3CLSimMediumProperties(0.9216 * I3Units::g / I3Units::cm3, static_cast<uint32_t>(m.num_layers), m.layers_zstart,
                                                               m.layers_height, std::min(-870. * I3Units::m, m.layers_zstart),
                                                               std::max(1940. * I3Units::m, top)));
    med->SetForcedMinWlen(265. * I3Units::nanometer);
    med->SetForcedMaxWlen(675. * I3Units::nanometer);

    I3CLSimRandomValueConstPtr scat;
    if (m.scat_kind == 1)
        scat = I3CLSimRandomValueConstPtr(new I3CLSimRandomValueHenyeyGreenstein(m.mean_cos));
    else if (m.scat_kind == 2)
        scat = I3CLSimRandomValueConstPtr(new I3CLSimRandomValueSimplifiedLiu(m.mean_cos));
    else
        scat = I3CLSimRandomValueConstPtr(new I3CLSimRandomValueMixed(m.f_sl, I3CLSimRandomValueConstPtr(new I3CLSimRandomValueSimplifiedLiu(m.mean_cos)),
                                                                      I3CLSimRandomValueConstPtr(new I3CLSimRandomValueHenyeyGreenstein(m.mean_cos))));
    med->SetScatteringCosAngleDistribution(scat);

    if (!m.has_anisotropy) {
        med->SetDirectionalAbsorptionLengthCorrection(I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldConstant(1.)));
        med->SetPreScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformConstant()));
        med->SetPostScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformConstant()));
    } else {
        // python/util/__init__.py GetSpiceLeaAnisotropyTransforms: the matrices are its numpy products, passed through
        med->SetDirectionalAbsorptionLengthCorrection(
            I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldAnisotropyAbsLenScaling(m.aniso_azimuth, m.aniso_along, m.aniso_perp)));
        I3Matrix pre(3, 3), post(3, 3);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                pre(i, j) = m.pre_matrix[3 * i + j];
                post(i, j) = m.post_matrix[3 * i + j];
            }
        med->SetPreScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformMatrix(pre, m.pre_renormalize != 0)));
        med->SetPostScatterDirectionTransform(I3CLSimVectorTransformConstPtr(new I3CLSimVectorTransformMatrix(post, m.post_renormalize != 0)));
    }

    if (m.tilt_num_dist > 0) {
        // python/util/__init__.py GetIceTiltZShift: distances, equally spaced z coordinates, corrections[dist][z]
        std::vector<double> dist(m.tilt_dist, m.tilt_dist + m.tilt_num_dist), zs(m.tilt_num_z);
        for (int k = 0; k < m.tilt_num_z; ++k) zs[k] = tilt_z ? tilt_z[k] : m.tilt_z0 + k * m.tilt_dz;   // (the struct carries first + mean spacing)
        I3Matrix corr(m.tilt_num_dist, m.tilt_num_z);
        for (int j = 0; j < m.tilt_num_dist; ++j)
            for (int k = 0; k < m.tilt_num_z; ++k) corr(j, k) = m.tilt_corr[j * m.tilt_num_z + k];
        med->SetIceTiltZShift(I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldIceTiltZShift(dist, zs, corr, m.tilt_azimuth)));
    } else {
        med->SetIceTiltZShift(I3CLSimScalarFieldConstPtr(new I3CLSimScalarFieldConstant(0.)));
    }

    I3CLSimFunctionConstPtr phase(new I3CLSimFunctionRefIndexIceCube("phase", m.n_phase[0], m.n_phase[1], m.n_phase[2], m.n_phase[3], m.n_phase[4],
                                                                     m.n_group[0], m.n_group[1], m.n_group[2], m.n_group[3], m.n_group[4]));
    I3CLSimFunctionConstPtr group(new I3CLSimFunctionRefIndexIceCube("group", m.n_phase[0], m.n_phase[1], m.n_phase[2], m.n_phase[3], m.n_phase[4],
                                                                     m.n_group[0], m.n_group[1], m.n_group[2], m.n_group[3], m.n_group[4]));
    for (int i = 0; i < m.num_layers; ++i) {
        med->SetPhaseRefractiveIndex(i, phase);
        med->SetGroupRefractiveIndexOverride(i, group);
        med->SetAbsorptionLength(i, I3CLSimFunctionConstPtr(new I3CLSimFunctionAbsLenIceCube(m.kappa, m.A, m.B, m.D, m.E, m.a_dust400[i], m.delta_tau[i])));
        med->SetScatteringLength(i, I3CLSimFunctionConstPtr(new I3CLSimFunctionScatLenIceCube(m.alpha, m.b400[i])));
    }
    return med;
}

// the objects private/clsim/I3CLSimModuleHelper.cxx:75-330 ends up with, from the oracle's description of each
I3CLSimRandomValueConstPtr make_generator(const oracle_wlen_generator &g)
{
    switch (g.kind) {
    case 0:
        return I3CLSimRandomValueConstPtr(new I3CLSimRandomValueInterpolatedDistribution(g.x0, g.dx, std::vector<double>(g.y, g.y + g.n)));
    case 1:
        return I3CLSimRandomValueConstPtr(
            new I3CLSimRandomValueInterpolatedDistribution(std::vector<double>(g.x, g.x + g.n), std::vector<double>(g.y, g.y + g.n)));
    case 2:
        return I3CLSimRandomValueConstPtr(new I3CLSimRandomValueWlenCherenkovNoDispersion(g.from_wlen, g.to_wlen));
    case 3:
        return I3CLSimRandomValueConstPtr(new I3CLSimRandomValueConstant(g.value));
    }
    throw std::runtime_error("unknown wavelength generator kind");
}

// the wavelength bias as the reference's tray segments hand it over: a table on equal bins, or a constant
inline I3CLSimFunctionConstPtr make_bias(const oracle_wlen_bias &b)
{
    if (b.kind == 0) return I3CLSimFunctionConstPtr(new I3CLSimFunctionConstant(b.value));
    return I3CLSimFunctionConstPtr(new I3CLSimFunctionFromTable(b.x0, b.dx, std::vector<double>(b.v, b.v + b.n)));
}

} // namespace
#endif
