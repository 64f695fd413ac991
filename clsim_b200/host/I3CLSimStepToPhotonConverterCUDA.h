// I3CLSimStepToPhotonConverterCUDA -- the reference's step -> photon converter interface
// (public/clsim/I3CLSimStepToPhotonConverter.h:67-192) implemented on libclsimcuda (include/clsimcuda.h),
// i.e. on hand-written sm_100a kernels.  Drop-in for I3CLSimStepToPhotonConverterOpenCL
// (public/clsim/I3CLSimStepToPhotonConverterOpenCL.h:78-258): the same setter / getter names, the same
// exceptions in the same situations, the same threading contract (EnqueueSteps blocks at 5 queued bunches,
// GetConversionResult blocks, results may come back in any order, all queries callable concurrently).
//
// Differences, all forced by the platform:
//   * SetDevice takes a CUDA ordinal (the OpenCL class takes an I3CLSimOpenCLDevice);
//   * the constructor takes a seed instead of an I3RandomServicePtr: the only use the reference makes of the
//     service is to draw the MWC x[] seeds (private/opencl/mwcrng_init.h:107-113);
//   * Compile() validates and flattens the description objects, nothing is JIT-compiled;
//   * SetDoublePrecision(true) throws; there is no CPU device and no fallback of any kind.
#ifndef I3CLSIMSTEPTOPHOTONCONVERTERCUDA_H_INCLUDED
#define I3CLSIMSTEPTOPHOTONCONVERTERCUDA_H_INCLUDED

#ifdef CLSIM_CUDA_IN_ICETRAY
#include "clsim/I3CLSimStepToPhotonConverter.h"
#include "clsim/function/I3CLSimFunction.h"
#include "clsim/random_value/I3CLSimRandomValue.h"
#include "clsim/I3CLSimMediumProperties.h"
#include "clsim/I3CLSimSimpleGeometry.h"
#else
#include "clsim_compat.h"
#endif

#include <cstdint>
#include <map>
#include <string>
#include <vector>

struct clsimcu_engine;
struct clsimcu_mcpe;

class I3CLSimStepToPhotonConverterCUDA : public I3CLSimStepToPhotonConverter {
public:
    static const bool default_useNativeMath;

    // useNativeMath=true -> the fast persistent kernel (approximate MUFU math, like the reference's
    // native_* OpenCL built-ins); false -> the reference-order kernel with precise math.
    explicit I3CLSimStepToPhotonConverterCUDA(uint64_t randomSeed = 0, bool useNativeMath = default_useNativeMath);
    virtual ~I3CLSimStepToPhotonConverterCUDA();

    // ---- knobs the factory calls (I3CLSimModuleHelper.cxx:319-369); all throw once initialized
    void SetDevice(int cudaOrdinal);
    void SetEnableDoubleBuffering(bool value);
    bool GetEnableDoubleBuffering() const;
    void SetDoublePrecision(bool value);
    bool GetDoublePrecision() const;
    void SetStopDetectedPhotons(bool value);
    bool GetStopDetectedPhotons() const;
    void SetSaveAllPhotons(bool value);
    bool GetSaveAllPhotons() const;
    void SetSaveAllPhotonsPrescale(double value);
    double GetSaveAllPhotonsPrescale() const;
    void SetFixedNumberOfAbsorptionLengths(double value);
    double GetFixedNumberOfAbsorptionLengths() const;
    void SetDOMPancakeFactor(double value);
    double GetDOMPancakeFactor() const;
    void SetPhotonHistoryEntries(uint32_t value);
    uint32_t GetPhotonHistoryEntries() const;
    void Compile();
    std::size_t GetMaxWorkgroupSize() const;
    void SetWorkgroupSize(std::size_t val);
    void SetMaxNumWorkitems(std::size_t val);
    // CUDA-only: each GPU of a box takes its own slice of the safe-prime multiplier table
    void SetFirstRNGMultiplierRow(uint64_t row);

    // ---- I3CLSimStepToPhotonConverter
    virtual void SetWlenGenerators(const std::vector<I3CLSimRandomValueConstPtr> &wlenGenerators);
    virtual void SetWlenBias(I3CLSimFunctionConstPtr wlenBias);
    virtual void SetMediumProperties(I3CLSimMediumPropertiesConstPtr mediumProperties);
    virtual void SetGeometry(I3CLSimSimpleGeometryConstPtr geometry);
    virtual void Initialize();
    virtual bool IsInitialized() const;
    virtual void EnqueueSteps(I3CLSimStepSeriesConstPtr steps, uint32_t identifier);
    virtual std::size_t GetWorkgroupSize() const;
    virtual std::size_t GetMaxNumWorkitems() const;
    virtual std::size_t QueueSize() const;
    virtual bool MorePhotonsAvailable() const;
    virtual ConversionResult_t GetConversionResult();
    virtual std::map<std::string, double> GetStatistics() const;

    // The flattened description (what Compile() produced), for tests: JSON text of the device tables.
    std::string DescribeTables() const;

    // the neighbours on the device (I3CLSimNeighboursCUDA.h) attach to the engine behind this converter
    clsimcu_engine *GetEngine() { return engine_; }
    // GetConversionResult plus the photo-electrons of an attached I3CLSimPhotonToMCPEConverterCUDA
    ConversionResult_t GetConversionResultWithMCPEs(std::vector<clsimcu_mcpe> &mcpes);

private:
    struct Flat;  // POD arrays behind the clsimcu_config pointers
    void ThrowIfInitialized() const;
    void ThrowIfNotInitialized() const;
    void Flatten();

    uint64_t randomSeed_;
    bool useNativeMath_;
    bool initialized_, compiled_, deviceIsSelected_;
    int device_;
    bool enableDoubleBuffering_, stopDetectedPhotons_, saveAllPhotons_;
    double saveAllPhotonsPrescale_, fixedNumberOfAbsorptionLengths_, pancakeFactor_;
    uint32_t photonHistoryEntries_;
    std::size_t workgroupSize_, maxNumWorkitems_;
    uint64_t firstRNGMultiplierRow_;
    std::vector<I3CLSimRandomValueConstPtr> wlenGenerators_;
    I3CLSimFunctionConstPtr wlenBias_;
    I3CLSimMediumPropertiesConstPtr mediumProperties_;
    I3CLSimSimpleGeometryConstPtr geometry_;
    Flat *flat_;
    clsimcu_engine *engine_;
};

#ifndef CLSIM_CUDA_IN_ICETRAY
typedef std::shared_ptr<I3CLSimStepToPhotonConverterCUDA> I3CLSimStepToPhotonConverterCUDAPtr;
#endif

// Factory with the call sequence of I3CLSimModuleHelper::initializeOpenCL (private/clsim/I3CLSimModuleHelper.cxx:303-372).
struct I3CLSimCUDADevice {
    int ordinal;
    std::size_t approximateNumberOfWorkItems; // I3CLSimOpenCLDevice::GetApproximateNumberOfWorkItems
    bool useNativeMath;
};
namespace I3CLSimModuleHelper {
std::shared_ptr<I3CLSimStepToPhotonConverterCUDA>
initializeCUDA(const I3CLSimCUDADevice &device, uint64_t randomSeed, I3CLSimSimpleGeometryConstPtr geometry, I3CLSimMediumPropertiesConstPtr medium,
               I3CLSimFunctionConstPtr wavelengthGenerationBias, const std::vector<I3CLSimRandomValueConstPtr> &wavelengthGenerators,
               bool enableDoubleBuffering, bool doublePrecision, bool stopDetectedPhotons, bool saveAllPhotons, double saveAllPhotonsPrescale,
               double fixedNumberOfAbsorptionLengths, double pancakeFactor, uint32_t photonHistoryEntries, uint32_t limitWorkgroupSize,
               uint64_t firstRNGMultiplierRow = 0);
}

#endif // I3CLSIMSTEPTOPHOTONCONVERTERCUDA_H_INCLUDED
