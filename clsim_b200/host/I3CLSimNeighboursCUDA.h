// I3CLSimNeighboursCUDA.h -- C++ host classes for the two neighbours of the step -> photon path that run on the
// device, over the C ABI of libclsimcuda:
//
//   I3CLSimPhotonToMCPEConverterCUDA   same constructor arguments and Convert semantics as the reference's
//       I3CLSimPhotonToMCPEConverterForDOMs (public/clsim/dom/I3PhotonToMCPEConverter.h:156-165,
//       private/clsim/dom/I3PhotonToMCPEConverter.cxx:595-669), on whole photon series; AttachTo() runs it in the
//       converter's stream so that GetConversionResultWithMCPEs() hands back photo-electrons.
//   I3CLSimStepGeneratorCUDA           MakeSteps of I3CLSimLightSourceToStepConverterPPC
//       (private/clsim/I3CLSimLightSourceToStepConverterPPC.cxx:523-640) for entries of its step generation
//       queue (CascadeStepData_t / MuonStepData_t); EnqueueInto() makes the bunch on the device.
//
// Inside IceTray compile with -DCLSIM_CUDA_IN_ICETRAY; stand-alone the stand-ins of clsim_compat.h are used.
#ifndef I3CLSIMNEIGHBOURSCUDA_H_INCLUDED
#define I3CLSIMNEIGHBOURSCUDA_H_INCLUDED

#include <cstdint>
#include <map>
#include <vector>

#include "clsimcuda.h"
#include "I3CLSimStepToPhotonConverterCUDA.h"

#ifdef CLSIM_CUDA_IN_ICETRAY
#include "icetray/OMKey.h"
#include "clsim/function/I3CLSimFunctionConstant.h"
#include "clsim/function/I3CLSimFunctionFromTable.h"
#include "clsim/function/I3CLSimFunctionPolynomial.h"
#else
// icetray/OMKey.h, the part used here
struct OMKey {
    OMKey(int string = 0, unsigned om = 0, unsigned char pmt = 0) : string_(string), om_(om), pmt_(pmt) {}
    int GetString() const { return string_; }
    unsigned GetOM() const { return om_; }
    bool operator<(const OMKey &o) const { return string_ != o.string_ ? string_ < o.string_ : (om_ != o.om_ ? om_ < o.om_ : pmt_ < o.pmt_); }

private:
    int string_;
    unsigned om_;
    unsigned char pmt_;
};

// public/clsim/function/I3CLSimFunctionPolynomial.h, the range-less form the DOM angular sensitivity uses
struct I3CLSimFunctionPolynomial : public I3CLSimFunction {
    explicit I3CLSimFunctionPolynomial(const std::vector<double> &coeffs) : coefficients_(coeffs) {}
    double GetValue(double x) const override
    {
        // private/clsim/function/I3CLSimFunctionPolynomial.cxx:86-102
        if (coefficients_.empty()) return 0.;
        double sum = coefficients_[0], multiplier = 1.;
        for (std::size_t i = 1; i < coefficients_.size(); ++i) {
            multiplier *= x;
            sum += coefficients_[i] * multiplier;
        }
        return sum;
    }
    const std::vector<double> &GetCoefficients() const { return coefficients_; }

private:
    std::vector<double> coefficients_;
};
#endif

class I3CLSimPhotonToMCPEConverterCUDA {
public:
    // `randomSeed` stands for the I3RandomServicePtr of the reference: the thinning draws come from MWC streams
    // on the device (rows [firstRNGMultiplierRow, ...) of the safe-prime table)
    I3CLSimPhotonToMCPEConverterCUDA(uint64_t randomSeed, const std::map<OMKey, I3CLSimFunctionConstPtr> &wavelengthAcceptance,
                                     I3CLSimFunctionConstPtr angularAcceptance, int device = 0, uint64_t firstRNGMultiplierRow = 0);
    ~I3CLSimPhotonToMCPEConverterCUDA();
    I3CLSimPhotonToMCPEConverterCUDA(const I3CLSimPhotonToMCPEConverterCUDA &) = delete;
    I3CLSimPhotonToMCPEConverterCUDA &operator=(const I3CLSimPhotonToMCPEConverterCUDA &) = delete;

    // Convert over a series; `uniforms` (one per photon) replaces the device's draws when given
    std::vector<clsimcu_mcpe> Convert(const I3CLSimPhotonSeries &photons, const std::vector<float> *uniforms = nullptr);
    // the conversion runs behind every propagation launch of `converter` (call after Initialize, before EnqueueSteps)
    void AttachTo(I3CLSimStepToPhotonConverterCUDA &converter, bool keepPhotons = false);
    clsimcu_mcpe_converter *handle() { return handle_; }

private:
    clsimcu_mcpe_converter *handle_;
};

class I3CLSimStepGeneratorCUDA {
public:
    // the two kinds of entries of I3CLSimLightSourceToStepConverterPPC::stepGenerationQueue_
    struct Source {
        double x, y, z, time;             // particle vertex
        double dirX, dirY, dirZ;          // particle direction (unit)
        uint32_t particleIdentifier;
        uint64_t photonsPerStep, numSteps, numPhotonsInLastStep;
        bool isCascade;                   // CascadeStepData_t: pa, pb;  else MuonStepData_t: stepIsCascadeLike, length
        double pa, pb;
        bool stepIsCascadeLike;
        double length;
    };
    explicit I3CLSimStepGeneratorCUDA(uint64_t randomSeed, int device = 0, uint64_t firstRNGMultiplierRow = 0, double angularDistA = 0.39,
                                      double angularDistB = 2.61);
    ~I3CLSimStepGeneratorCUDA();
    I3CLSimStepGeneratorCUDA(const I3CLSimStepGeneratorCUDA &) = delete;
    I3CLSimStepGeneratorCUDA &operator=(const I3CLSimStepGeneratorCUDA &) = delete;

    // MakeSteps: the steps of these entries, made on the device, copied back
    I3CLSimStepSeriesPtr MakeSteps(const std::vector<Source> &sources);
    // EnqueueSteps on `converter` for the bunch these entries describe; it never exists on the host.  Returns its size.
    std::size_t EnqueueInto(I3CLSimStepToPhotonConverterCUDA &converter, const std::vector<Source> &sources, uint32_t identifier);

private:
    static std::vector<clsimcu_step_source> Flatten(const std::vector<Source> &sources);
    clsimcu_step_generator *handle_;
};

#endif // I3CLSIMNEIGHBOURSCUDA_H_INCLUDED
