// clsim_compat.h -- stand-ins for the reference's public types, so that
// I3CLSimStepToPhotonConverterCUDA compiles and is testable OUTSIDE an IceTray build (this image has
// no icetray / dataclasses / boost).
//
// Inside the reference tree, compile I3CLSimStepToPhotonConverterCUDA.cxx with
// -DCLSIM_CUDA_IN_ICETRAY: this file is then skipped and the real headers are used
// (public/clsim/I3CLSimStep.h, I3CLSimPhoton.h, I3CLSimStepToPhotonConverter.h, function/*.h,
// random_value/*.h ...).  Every class below has the NAME, constructor arguments and accessors of
// its reference twin (file cited at each).  Accessors marked [getter to add upstream] do not
// exist in the reference, whose classes only expose GetOpenCLFunction() source text for these
// members; INTEGRATION.md lists the one-line getters a maintainer adds.
//
// Only description (data-holding) behaviour is restated; nothing here computes on the hot path.
#ifndef CLSIM_COMPAT_H_INCLUDED
#define CLSIM_COMPAT_H_INCLUDED

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CLSIM_POINTER_TYPEDEFS(T)           \
    typedef std::shared_ptr<T> T##Ptr;      \
    typedef std::shared_ptr<const T> T##ConstPtr

// ---- records (public/clsim/I3CLSimStep.h:68-155, I3CLSimPhoton.h:67-213) ---------------------------
struct I3CLSimStep {
    float posAndTime[4];           // x, y, z, time
    float dirAndLengthAndBeta[4];  // theta, phi, length, beta
    uint32_t numPhotons;
    float weight;
    uint32_t identifier;
    uint8_t sourceType;
    uint8_t dummy1;
    uint16_t dummy2;

    float GetPosX() const { return posAndTime[0]; }
    float GetPosY() const { return posAndTime[1]; }
    float GetPosZ() const { return posAndTime[2]; }
    float GetTime() const { return posAndTime[3]; }
    float GetDirTheta() const { return dirAndLengthAndBeta[0]; }
    float GetDirPhi() const { return dirAndLengthAndBeta[1]; }
    float GetLength() const { return dirAndLengthAndBeta[2]; }
    float GetBeta() const { return dirAndLengthAndBeta[3]; }
    uint32_t GetNumPhotons() const { return numPhotons; }
    float GetWeight() const { return weight; }
    uint32_t GetID() const { return identifier; }
    uint8_t GetSourceType() const { return sourceType; }
    void SetPosX(float v) { posAndTime[0] = v; }
    void SetPosY(float v) { posAndTime[1] = v; }
    void SetPosZ(float v) { posAndTime[2] = v; }
    void SetTime(float v) { posAndTime[3] = v; }
    void SetDirTheta(float v) { dirAndLengthAndBeta[0] = v; }
    void SetDirPhi(float v) { dirAndLengthAndBeta[1] = v; }
    void SetLength(float v) { dirAndLengthAndBeta[2] = v; }
    void SetBeta(float v) { dirAndLengthAndBeta[3] = v; }
    void SetNumPhotons(uint32_t v) { numPhotons = v; }
    void SetWeight(float v) { weight = v; }
    void SetID(uint32_t v) { identifier = v; }
    void SetSourceType(uint8_t v) { sourceType = v; }
    // direction of travel -> (theta, phi), I3CLSimStep.h:128-133 (I3Direction::CalcTheta/CalcPhi)
    void SetDir(double x, double y, double z)
    {
        const double len = std::sqrt(x * x + y * y + z * z);
        double phi = std::atan2(y, x);
        if (phi < 0) phi += 2.0 * M_PI;
        dirAndLengthAndBeta[0] = static_cast<float>(std::acos(z / len));
        dirAndLengthAndBeta[1] = static_cast<float>(phi);
    }
};
static_assert(sizeof(I3CLSimStep) == 48, "I3CLSimStep must be the 48-byte device record (private/clsim/I3CLSimStep.cxx:35-37)");

struct I3CLSimPhoton {
    float posAndTime[4];
    float dir[2];
    float wavelength;
    float cherenkovDist;
    uint32_t numScatters;
    float weight;
    uint32_t identifier;
    int16_t stringID;
    uint16_t omID;
    float startPosAndTime[4];
    float startDir[2];
    float groupVelocity;
    float distInAbsLens;

    float GetPosX() const { return posAndTime[0]; }
    float GetPosY() const { return posAndTime[1]; }
    float GetPosZ() const { return posAndTime[2]; }
    float GetTime() const { return posAndTime[3]; }
    float GetDirTheta() const { return dir[0]; }
    float GetDirPhi() const { return dir[1]; }
    float GetStartPosX() const { return startPosAndTime[0]; }
    float GetStartPosY() const { return startPosAndTime[1]; }
    float GetStartPosZ() const { return startPosAndTime[2]; }
    float GetStartTime() const { return startPosAndTime[3]; }
    float GetStartDirTheta() const { return startDir[0]; }
    float GetStartDirPhi() const { return startDir[1]; }
    float GetWavelength() const { return wavelength; }
    float GetCherenkovDist() const { return cherenkovDist; }
    uint32_t GetNumScatters() const { return numScatters; }
    float GetWeight() const { return weight; }
    void SetWeight(const float &val) { weight = val; }   // public/clsim/I3CLSimPhoton.h:119
    uint32_t GetID() const { return identifier; }
    int16_t GetStringID() const { return stringID; }
    uint16_t GetOMID() const { return omID; }
    float GetGroupVelocity() const { return groupVelocity; }
    float GetDistInAbsLens() const { return distInAbsLens; }
};
static_assert(sizeof(I3CLSimPhoton) == 80, "I3CLSimPhoton must be the 80-byte device record (private/clsim/I3CLSimPhoton.cxx:36)");

typedef std::vector<I3CLSimStep> I3CLSimStepSeries;
typedef std::vector<I3CLSimPhoton> I3CLSimPhotonSeries;
CLSIM_POINTER_TYPEDEFS(I3CLSimStepSeries);
CLSIM_POINTER_TYPEDEFS(I3CLSimPhotonSeries);

// public/clsim/I3CLSimPhotonHistory.h:43-73
class I3CLSimPhotonHistory {
public:
    std::size_t size() const { return posX_.size(); }
    float GetX(std::size_t i) const { return posX_[i]; }
    float GetY(std::size_t i) const { return posY_[i]; }
    float GetZ(std::size_t i) const { return posZ_[i]; }
    float GetDistanceInAbsorptionLengths(std::size_t i) const { return distanceInAbsorptionLengths_[i]; }
    void push_back(float x, float y, float z, float abslens)
    {
        posX_.push_back(x); posY_.push_back(y); posZ_.push_back(z); distanceInAbsorptionLengths_.push_back(abslens);
    }

private:
    std::vector<float> posX_, posY_, posZ_, distanceInAbsorptionLengths_;
};
typedef std::vector<I3CLSimPhotonHistory> I3CLSimPhotonHistorySeries;
CLSIM_POINTER_TYPEDEFS(I3CLSimPhotonHistorySeries);

// ---- wavelength functions (public/clsim/function/I3CLSimFunction*.h) -------------------------------
struct I3CLSimFunction {
    virtual ~I3CLSimFunction() {}
    virtual double GetValue(double wlen) const = 0;
    virtual double GetMinWlen() const { return -std::numeric_limits<double>::infinity(); }
    virtual double GetMaxWlen() const { return std::numeric_limits<double>::infinity(); }
};
CLSIM_POINTER_TYPEDEFS(I3CLSimFunction);

// function/I3CLSimFunctionConstant.h:39-98
struct I3CLSimFunctionConstant : public I3CLSimFunction {
    explicit I3CLSimFunctionConstant(double value) : value_(value) {}
    double GetValue(double) const override { return value_; }

private:
    double value_;
};

// function/I3CLSimFunctionFromTable.h:40-135
struct I3CLSimFunctionFromTable : public I3CLSimFunction {
    I3CLSimFunctionFromTable(const std::vector<double> &wlens, const std::vector<double> &values)
        : startWlen_(NAN), wlenStep_(NAN), wlens_(wlens), values_(values), equalSpacingMode_(false)
    {
        if (wlens_.size() != values_.size() || values_.size() < 2) throw std::runtime_error("The wlens and values vectors must have the same size (>= 2)!");
    }
    I3CLSimFunctionFromTable(double startWlen, double wlenStep, const std::vector<double> &values)
        : startWlen_(startWlen), wlenStep_(wlenStep), values_(values), equalSpacingMode_(true)
    {
        if (values_.size() < 2) throw std::runtime_error("The values vector must contain at least 2 elements!");
        for (std::size_t i = 0; i < values_.size(); ++i) wlens_.push_back(startWlen_ + wlenStep_ * static_cast<double>(i));
    }
    double GetValue(double wlen) const override; // unused by the converter
    double GetMinWlen() const override { return wlens_.front(); }
    double GetMaxWlen() const override { return wlens_.back(); }
    double GetFirstWavelength() const { return startWlen_; }
    double GetWavelengthStepping() const { return wlenStep_; }
    std::size_t GetNumEntries() const { return values_.size(); }
    double GetEntryValue(std::size_t i) const { return values_[i]; }
    double GetEntryWavelength(std::size_t i) const { return wlens_[i]; }
    bool GetInEqualSpacingMode() const { return equalSpacingMode_; }

private:
    double startWlen_, wlenStep_;
    std::vector<double> wlens_, values_;
    bool equalSpacingMode_;
};
inline double I3CLSimFunctionFromTable::GetValue(double wlen) const
{
    // host twin, private/clsim/function/I3CLSimFunctionFromTable.cxx:107-147 (equal spacing branch)
    if (!equalSpacingMode_) throw std::runtime_error("GetValue is only restated for equal spacing here");
    double whole;
    double frac = std::modf((wlen - startWlen_) / wlenStep_, &whole);
    long bin = static_cast<long>(whole);
    if (bin < 0 || (bin == 0 && frac < 0)) { bin = 0; frac = 0; }
    else if (bin >= static_cast<long>(values_.size()) - 1) { bin = static_cast<long>(values_.size()) - 2; frac = 1; }
    return values_[bin] + (values_[bin + 1] - values_[bin]) * frac;
}

// function/I3CLSimFunctionAbsLenIceCube.h:40-111
struct I3CLSimFunctionAbsLenIceCube : public I3CLSimFunction {
    I3CLSimFunctionAbsLenIceCube(double kappa, double A, double B, double D, double E, double aDust400, double deltaTau)
        : kappa_(kappa), A_(A), B_(B), D_(D), E_(E), aDust400_(aDust400), deltaTau_(deltaTau) {}
    double GetValue(double wlen) const override
    {
        const double x = wlen / 1e-9; // private/clsim/function/I3CLSimFunctionAbsLenIceCube.cxx:63-67
        return 1.0 / ((D_ * aDust400_ + E_) * std::pow(x, -kappa_) + A_ * std::exp(-B_ / x) * (1.0 + 0.01 * deltaTau_));
    }
    double GetKappa() const { return kappa_; }
    double GetA() const { return A_; }
    double GetB() const { return B_; }
    double GetD() const { return D_; }
    double GetE() const { return E_; }
    double GetADust400() const { return aDust400_; }
    double GetDeltaTau() const { return deltaTau_; }

private:
    double kappa_, A_, B_, D_, E_, aDust400_, deltaTau_;
};

// function/I3CLSimFunctionScatLenIceCube.h:40-96
struct I3CLSimFunctionScatLenIceCube : public I3CLSimFunction {
    I3CLSimFunctionScatLenIceCube(double alpha, double b400) : alpha_(alpha), b400_(b400) {}
    double GetValue(double wlen) const override { return 1.0 / (b400_ * std::pow(wlen / 400e-9, -alpha_)); }
    double GetAlpha() const { return alpha_; }
    double GetB400() const { return b400_; }

private:
    double alpha_, b400_;
};

// function/I3CLSimFunctionRefIndexIceCube.h:42-133 (defaults private/clsim/function/I3CLSimFunctionRefIndexIceCube.cxx:38-47)
struct I3CLSimFunctionRefIndexIceCube : public I3CLSimFunction {
    explicit I3CLSimFunctionRefIndexIceCube(const std::string &mode = "phase", double n0 = 1.55749, double n1 = -1.57988, double n2 = 3.99993,
                                            double n3 = -4.68271, double n4 = 2.09354, double g0 = 1.227106, double g1 = -0.954648,
                                            double g2 = 1.42568, double g3 = -0.711832, double g4 = 0.0)
        : mode_(mode), n_{n0, n1, n2, n3, n4}, g_{g0, g1, g2, g3, g4}
    {
        if (mode_ != "phase" && mode_ != "group") throw std::runtime_error("Invalid mode: " + mode_);
    }
    double GetValue(double wlen) const override
    {
        const double x = wlen / 1e-6;
        const double np = n_[0] + x * (n_[1] + x * (n_[2] + x * (n_[3] + x * n_[4])));
        if (mode_ == "phase") return np;
        return np * (g_[0] + x * (g_[1] + x * (g_[2] + x * (g_[3] + x * g_[4]))));
    }
    const std::string &GetMode() const { return mode_; }            // [getter to add upstream]
    double GetPhaseCoefficient(int i) const { return n_[i]; }       // [getter to add upstream]
    double GetGroupCoefficient(int i) const { return g_[i]; }       // [getter to add upstream]

private:
    std::string mode_;
    double n_[5], g_[5];
};

// ---- random values (public/clsim/random_value/*.h) ---------------------------------------------------
struct I3CLSimRandomValue {
    virtual ~I3CLSimRandomValue() {}
};
CLSIM_POINTER_TYPEDEFS(I3CLSimRandomValue);

// random_value/I3CLSimRandomValueHenyeyGreenstein.h
struct I3CLSimRandomValueHenyeyGreenstein : public I3CLSimRandomValue {
    explicit I3CLSimRandomValueHenyeyGreenstein(double meanCosine) : meanCosine_(meanCosine) {}
    double GetMeanCosine() const { return meanCosine_; }            // [getter to add upstream]

private:
    double meanCosine_;
};
// random_value/I3CLSimRandomValueSimplifiedLiu.h
struct I3CLSimRandomValueSimplifiedLiu : public I3CLSimRandomValue {
    explicit I3CLSimRandomValueSimplifiedLiu(double meanCosine) : meanCosine_(meanCosine) {}
    double GetMeanCosine() const { return meanCosine_; }            // [getter to add upstream]

private:
    double meanCosine_;
};
// random_value/I3CLSimRandomValueMixed.h:36-80
struct I3CLSimRandomValueMixed : public I3CLSimRandomValue {
    I3CLSimRandomValueMixed(double fractionOfFirstDistribution, I3CLSimRandomValueConstPtr firstDistribution,
                            I3CLSimRandomValueConstPtr secondDistribution)
        : fractionOfFirstDistribution_(fractionOfFirstDistribution), firstDistribution_(firstDistribution), secondDistribution_(secondDistribution)
    {
        if (fractionOfFirstDistribution_ < 0 || fractionOfFirstDistribution_ > 1) throw std::runtime_error("fractionOfFirstDistribution must be in [0,1]");
    }
    double GetFractionOfFirstDistribution() const { return fractionOfFirstDistribution_; }   // [getter to add upstream]
    I3CLSimRandomValueConstPtr GetFirstDistribution() const { return firstDistribution_; }   // [getter to add upstream]
    I3CLSimRandomValueConstPtr GetSecondDistribution() const { return secondDistribution_; } // [getter to add upstream]

private:
    double fractionOfFirstDistribution_;
    I3CLSimRandomValueConstPtr firstDistribution_, secondDistribution_;
};
// random_value/I3CLSimRandomValueInterpolatedDistribution.h:40-100
struct I3CLSimRandomValueInterpolatedDistribution : public I3CLSimRandomValue {
    I3CLSimRandomValueInterpolatedDistribution(const std::vector<double> &x, const std::vector<double> &y)
        : x_(x), y_(y), constantXSpacing_(false), firstX_(NAN), xSpacing_(NAN)
    {
        if (x_.size() != y_.size() || y_.size() < 2) throw std::runtime_error("The x and y vectors must have the same size (>= 2)!");
    }
    I3CLSimRandomValueInterpolatedDistribution(double xFirst, double xSpacing, const std::vector<double> &y)
        : y_(y), constantXSpacing_(true), firstX_(xFirst), xSpacing_(xSpacing)
    {
        if (y_.size() < 2) throw std::runtime_error("The y vector must have at least 2 entries!");
    }
    bool GetConstantXSpacing() const { return constantXSpacing_; }  // [getter to add upstream]
    double GetFirstX() const { return firstX_; }                    // [getter to add upstream]
    double GetXSpacing() const { return xSpacing_; }                // [getter to add upstream]
    const std::vector<double> &GetX() const { return x_; }          // [getter to add upstream]
    const std::vector<double> &GetY() const { return y_; }          // [getter to add upstream]

private:
    std::vector<double> x_, y_;
    bool constantXSpacing_;
    double firstX_, xSpacing_;
};
// random_value/I3CLSimRandomValueWlenCherenkovNoDispersion.h
struct I3CLSimRandomValueWlenCherenkovNoDispersion : public I3CLSimRandomValue {
    I3CLSimRandomValueWlenCherenkovNoDispersion(double fromWlen, double toWlen) : fromWlen_(fromWlen), toWlen_(toWlen) {}
    double GetFromWlen() const { return fromWlen_; }                // [getter to add upstream]
    double GetToWlen() const { return toWlen_; }                    // [getter to add upstream]

private:
    double fromWlen_, toWlen_;
};
// random_value/I3CLSimRandomValueConstant.h
struct I3CLSimRandomValueConstant : public I3CLSimRandomValue {
    explicit I3CLSimRandomValueConstant(double value) : value_(value) {}
    double GetValue() const { return value_; }                      // [getter to add upstream]

private:
    double value_;
};

// ---- scalar fields and vector transforms (public/clsim/function/I3CLSimScalarField*.h, I3CLSimVectorTransform*.h)
struct I3CLSimScalarField {
    virtual ~I3CLSimScalarField() {}
};
CLSIM_POINTER_TYPEDEFS(I3CLSimScalarField);
struct I3CLSimScalarFieldConstant : public I3CLSimScalarField {
    explicit I3CLSimScalarFieldConstant(double value) : value_(value) {}
    double GetValue(double, double, double) const { return value_; }

private:
    double value_;
};
// function/I3CLSimScalarFieldIceTiltZShift.h:43-92; zCorrections[iDistance][iZ]
struct I3CLSimScalarFieldIceTiltZShift : public I3CLSimScalarField {
    I3CLSimScalarFieldIceTiltZShift(const std::vector<double> &distancesFromOriginAlongTilt, const std::vector<double> &zCoordinates,
                                    const std::vector<std::vector<double> > &zCorrections, double directionOfTiltAzimuth = 225.0 * M_PI / 180.0)
        : distancesFromOriginAlongTilt_(distancesFromOriginAlongTilt), zCoordinates_(zCoordinates), zCorrections_(zCorrections),
          directionOfTiltAzimuth_(directionOfTiltAzimuth)
    {
        // private/clsim/function/I3CLSimScalarFieldIceTiltZShift.cxx:53-110
        if (zCorrections_.size() != distancesFromOriginAlongTilt_.size()) throw std::runtime_error("zCorrections: dimension 1 must match distancesFromOriginAlongTilt");
        for (std::size_t i = 0; i < zCorrections_.size(); ++i)
            if (zCorrections_[i].size() != zCoordinates_.size()) throw std::runtime_error("zCorrections: dimension 2 must match zCoordinates");
        if (zCoordinates_.size() < 2 || distancesFromOriginAlongTilt_.size() < 2) throw std::runtime_error("need at least 2 nodes in each dimension");
        double sum = 0;
        for (std::size_t i = 1; i < zCoordinates_.size(); ++i) {
            if (!(zCoordinates_[i] > zCoordinates_[i - 1])) throw std::runtime_error("zCoordinates (dimension 2) is not in ascending order.");
            sum += zCoordinates_[i] - zCoordinates_[i - 1];
        }
        zCoordinateSpacing_ = sum / static_cast<double>(zCoordinates_.size() - 1);
        for (std::size_t i = 1; i < zCoordinates_.size(); ++i)
            if (std::fabs(zCoordinates_[i] - zCoordinates_[i - 1] - zCoordinateSpacing_) > 1e-5) throw std::runtime_error("zCoordinates (dimension 2) are not equally spaced");
        for (std::size_t i = 1; i < distancesFromOriginAlongTilt_.size(); ++i)
            if (!(distancesFromOriginAlongTilt_[i] > distancesFromOriginAlongTilt_[i - 1]))
                throw std::runtime_error("distancesFromOriginAlongTilt (dimension 1) is not in ascending order.");
        firstZCoordinate_ = zCoordinates_[0];
    }
    const std::vector<double> &GetDistancesFromOriginAlongTilt() const { return distancesFromOriginAlongTilt_; } // [getter to add upstream]
    const std::vector<double> &GetZCoordinates() const { return zCoordinates_; }                                 // [getter to add upstream]
    const std::vector<std::vector<double> > &GetZCorrections() const { return zCorrections_; }                  // [getter to add upstream] (I3Matrix there)
    double GetDirectionOfTiltAzimuth() const { return directionOfTiltAzimuth_; }                                 // [getter to add upstream]
    double GetFirstZCoordinate() const { return firstZCoordinate_; }                                             // [getter to add upstream]
    double GetZCoordinateSpacing() const { return zCoordinateSpacing_; }                                         // [getter to add upstream]

private:
    std::vector<double> distancesFromOriginAlongTilt_, zCoordinates_;
    std::vector<std::vector<double> > zCorrections_;
    double directionOfTiltAzimuth_, firstZCoordinate_, zCoordinateSpacing_;
};
// function/I3CLSimScalarFieldAnisotropyAbsLenScaling.h:47-91
struct I3CLSimScalarFieldAnisotropyAbsLenScaling : public I3CLSimScalarField {
    explicit I3CLSimScalarFieldAnisotropyAbsLenScaling(double anisotropyDirAzimuth = 216.0 * M_PI / 180.0, double magnitudeAlongDir = 0.04,
                                                       double magnitudePerpToDir = -0.08)
        : anisotropyDirAzimuth_(anisotropyDirAzimuth), magnitudeAlongDir_(magnitudeAlongDir), magnitudePerpToDir_(magnitudePerpToDir) {}
    double GetAnisotropyDirAzimuth() const { return anisotropyDirAzimuth_; } // [getter to add upstream]
    double GetMagnitudeAlongDir() const { return magnitudeAlongDir_; }       // [getter to add upstream]
    double GetMagnitudePerpToDir() const { return magnitudePerpToDir_; }     // [getter to add upstream]

private:
    double anisotropyDirAzimuth_, magnitudeAlongDir_, magnitudePerpToDir_;
};

struct I3CLSimVectorTransform {
    virtual ~I3CLSimVectorTransform() {}
};
CLSIM_POINTER_TYPEDEFS(I3CLSimVectorTransform);
struct I3CLSimVectorTransformConstant : public I3CLSimVectorTransform {}; // identity
// function/I3CLSimVectorTransformMatrix.h:40-75 (row-major 3x3 here, I3Matrix there)
struct I3CLSimVectorTransformMatrix : public I3CLSimVectorTransform {
    I3CLSimVectorTransformMatrix(const double (&matrix)[9], bool renormalize = false) : renormalize_(renormalize)
    {
        for (int i = 0; i < 9; ++i) matrix_[i] = matrix[i];
    }
    double GetMatrixElement(int row, int col) const { return matrix_[3 * row + col]; } // [getter to add upstream]
    bool GetRenormalize() const { return renormalize_; }                                // [getter to add upstream]

private:
    double matrix_[9];
    bool renormalize_;
};

// ---- medium (public/clsim/I3CLSimMediumProperties.h:48-150) ----------------------------------------
class I3CLSimMediumProperties {
public:
    // class defaults private/clsim/I3CLSimMediumProperties.cxx:41-47
    explicit I3CLSimMediumProperties(double mediumDensity = 0.9216, uint32_t layersNum = 1, double layersZStart = -5000.0,
                                     double layersHeight = 10000.0, double rockZCoordinate = -870.0, double airZCoordinate = 1940.0)
        : mediumDensity_(mediumDensity), layersNum_(layersNum), layersZStart_(layersZStart), layersHeight_(layersHeight),
          rockZCoordinate_(rockZCoordinate), airZCoordinate_(airZCoordinate), forcedMinWlen_(-std::numeric_limits<double>::infinity()),
          forcedMaxWlen_(std::numeric_limits<double>::infinity()), efficiency_(1.0), absorptionLength_(layersNum), scatteringLength_(layersNum),
          phaseRefractiveIndex_(layersNum), groupRefractiveIndexOverride_(layersNum),
          directionalAbsorptionLengthCorrection_(new I3CLSimScalarFieldConstant(1.0)), preScatterDirectionTransform_(new I3CLSimVectorTransformConstant()),
          postScatterDirectionTransform_(new I3CLSimVectorTransformConstant()), iceTiltZShift_(new I3CLSimScalarFieldConstant(0.0))
    {
        if (layersNum_ == 0) throw std::runtime_error("layersNum must be > 0");
    }
    bool IsReady() const
    {
        for (uint32_t i = 0; i < layersNum_; ++i)
            if (!absorptionLength_[i] || !scatteringLength_[i] || !phaseRefractiveIndex_[i]) return false;
        return static_cast<bool>(scatteringCosAngleDistribution_);
    }
    void SetAbsorptionLength(uint32_t layer, I3CLSimFunctionConstPtr p) { absorptionLength_.at(layer) = p; }
    void SetScatteringLength(uint32_t layer, I3CLSimFunctionConstPtr p) { scatteringLength_.at(layer) = p; }
    void SetPhaseRefractiveIndex(uint32_t layer, I3CLSimFunctionConstPtr p) { phaseRefractiveIndex_.at(layer) = p; }
    void SetGroupRefractiveIndexOverride(uint32_t layer, I3CLSimFunctionConstPtr p) { groupRefractiveIndexOverride_.at(layer) = p; }
    void SetScatteringCosAngleDistribution(I3CLSimRandomValueConstPtr p) { scatteringCosAngleDistribution_ = p; }
    void SetDirectionalAbsorptionLengthCorrection(I3CLSimScalarFieldConstPtr p) { directionalAbsorptionLengthCorrection_ = p; }
    void SetPreScatterDirectionTransform(I3CLSimVectorTransformConstPtr p) { preScatterDirectionTransform_ = p; }
    void SetPostScatterDirectionTransform(I3CLSimVectorTransformConstPtr p) { postScatterDirectionTransform_ = p; }
    void SetIceTiltZShift(I3CLSimScalarFieldConstPtr p) { iceTiltZShift_ = p; }
    I3CLSimFunctionConstPtr GetAbsorptionLength(uint32_t layer) const { return absorptionLength_.at(layer); }
    I3CLSimFunctionConstPtr GetScatteringLength(uint32_t layer) const { return scatteringLength_.at(layer); }
    I3CLSimFunctionConstPtr GetPhaseRefractiveIndex(uint32_t layer) const { return phaseRefractiveIndex_.at(layer); }
    I3CLSimFunctionConstPtr GetGroupRefractiveIndexOverride(uint32_t layer) const { return groupRefractiveIndexOverride_.at(layer); }
    I3CLSimRandomValueConstPtr GetScatteringCosAngleDistribution() const { return scatteringCosAngleDistribution_; }
    I3CLSimScalarFieldConstPtr GetDirectionalAbsorptionLengthCorrection() const { return directionalAbsorptionLengthCorrection_; }
    I3CLSimVectorTransformConstPtr GetPreScatterDirectionTransform() const { return preScatterDirectionTransform_; }
    I3CLSimVectorTransformConstPtr GetPostScatterDirectionTransform() const { return postScatterDirectionTransform_; }
    I3CLSimScalarFieldConstPtr GetIceTiltZShift() const { return iceTiltZShift_; }
    double GetMediumDensity() const { return mediumDensity_; }
    uint32_t GetLayersNum() const { return layersNum_; }
    double GetLayersZStart() const { return layersZStart_; }
    double GetLayersHeight() const { return layersHeight_; }
    double GetRockZCoord() const { return rockZCoordinate_; }
    double GetAirZCoord() const { return airZCoordinate_; }
    double GetForcedMinWlen() const { return forcedMinWlen_; }
    double GetForcedMaxWlen() const { return forcedMaxWlen_; }
    void SetForcedMinWlen(double v) { forcedMinWlen_ = v; }
    void SetForcedMaxWlen(double v) { forcedMaxWlen_ = v; }
    double GetEfficiency() const { return efficiency_; }
    void SetEfficiency(double v) { efficiency_ = v; }

private:
    double mediumDensity_;
    uint32_t layersNum_;
    double layersZStart_, layersHeight_, rockZCoordinate_, airZCoordinate_, forcedMinWlen_, forcedMaxWlen_, efficiency_;
    std::vector<I3CLSimFunctionConstPtr> absorptionLength_, scatteringLength_, phaseRefractiveIndex_, groupRefractiveIndexOverride_;
    I3CLSimRandomValueConstPtr scatteringCosAngleDistribution_;
    I3CLSimScalarFieldConstPtr directionalAbsorptionLengthCorrection_;
    I3CLSimVectorTransformConstPtr preScatterDirectionTransform_, postScatterDirectionTransform_;
    I3CLSimScalarFieldConstPtr iceTiltZShift_;
};
CLSIM_POINTER_TYPEDEFS(I3CLSimMediumProperties);

// ---- geometry (public/clsim/I3CLSimSimpleGeometry.h:40-66, I3CLSimSimpleGeometryUserConfigurable.h) -----
class I3CLSimSimpleGeometry {
public:
    virtual ~I3CLSimSimpleGeometry() {}
    virtual std::size_t size() const = 0;
    virtual double GetOMRadius() const = 0;
    virtual const std::vector<int32_t> &GetStringIDVector() const = 0;
    virtual const std::vector<uint32_t> &GetDomIDVector() const = 0;
    virtual const std::vector<double> &GetPosXVector() const = 0;
    virtual const std::vector<double> &GetPosYVector() const = 0;
    virtual const std::vector<double> &GetPosZVector() const = 0;
    virtual const std::vector<std::string> &GetSubdetectorVector() const = 0;
};
CLSIM_POINTER_TYPEDEFS(I3CLSimSimpleGeometry);

class I3CLSimSimpleGeometryUserConfigurable : public I3CLSimSimpleGeometry {
public:
    I3CLSimSimpleGeometryUserConfigurable(double OMRadius, std::size_t numOMs)
        : OMRadius_(OMRadius), stringIDs_(numOMs, 0), domIDs_(numOMs, 0), posX_(numOMs, NAN), posY_(numOMs, NAN), posZ_(numOMs, NAN),
          subdetectors_(numOMs, "") {}
    std::size_t size() const override { return stringIDs_.size(); }
    double GetOMRadius() const override { return OMRadius_; }
    const std::vector<int32_t> &GetStringIDVector() const override { return stringIDs_; }
    const std::vector<uint32_t> &GetDomIDVector() const override { return domIDs_; }
    const std::vector<double> &GetPosXVector() const override { return posX_; }
    const std::vector<double> &GetPosYVector() const override { return posY_; }
    const std::vector<double> &GetPosZVector() const override { return posZ_; }
    const std::vector<std::string> &GetSubdetectorVector() const override { return subdetectors_; }
    void SetStringID(std::size_t pos, int32_t v) { stringIDs_.at(pos) = v; }
    void SetDomID(std::size_t pos, uint32_t v) { domIDs_.at(pos) = v; }
    void SetPosX(std::size_t pos, double v) { posX_.at(pos) = v; }
    void SetPosY(std::size_t pos, double v) { posY_.at(pos) = v; }
    void SetPosZ(std::size_t pos, double v) { posZ_.at(pos) = v; }
    void SetSubdetector(std::size_t pos, const std::string &v) { subdetectors_.at(pos) = v; }

private:
    double OMRadius_;
    std::vector<int32_t> stringIDs_;
    std::vector<uint32_t> domIDs_;
    std::vector<double> posX_, posY_, posZ_;
    std::vector<std::string> subdetectors_;
};

// ---- the interface being implemented (public/clsim/I3CLSimStepToPhotonConverter.h:57-192) ------------
class I3CLSimStepToPhotonConverter_exception : public std::runtime_error {
public:
    explicit I3CLSimStepToPhotonConverter_exception(const std::string &msg) : std::runtime_error(msg) {}
};

struct I3CLSimStepToPhotonConverter {
    struct ConversionResult_t {
        ConversionResult_t() : identifier(0) {}
        explicit ConversionResult_t(uint32_t identifier_, I3CLSimPhotonSeriesPtr photons_ = I3CLSimPhotonSeriesPtr(),
                                    I3CLSimPhotonHistorySeriesPtr photonHistories_ = I3CLSimPhotonHistorySeriesPtr())
            : identifier(identifier_), photons(photons_), photonHistories(photonHistories_) {}
        uint32_t identifier;
        I3CLSimPhotonSeriesPtr photons;
        I3CLSimPhotonHistorySeriesPtr photonHistories;
    };
    I3CLSimStepToPhotonConverter() {}
    I3CLSimStepToPhotonConverter(const I3CLSimStepToPhotonConverter &) = delete;            // boost::noncopyable there
    I3CLSimStepToPhotonConverter &operator=(const I3CLSimStepToPhotonConverter &) = delete;
    virtual ~I3CLSimStepToPhotonConverter() {}
    virtual void SetWlenGenerators(const std::vector<I3CLSimRandomValueConstPtr> &wlenGenerators) = 0;
    virtual void SetWlenBias(I3CLSimFunctionConstPtr wlenBias) = 0;
    virtual void SetMediumProperties(I3CLSimMediumPropertiesConstPtr mediumProperties) = 0;
    virtual void SetGeometry(I3CLSimSimpleGeometryConstPtr geometry) = 0;
    virtual void Initialize() = 0;
    virtual bool IsInitialized() const = 0;
    virtual void EnqueueSteps(I3CLSimStepSeriesConstPtr steps, uint32_t identifier) = 0;
    virtual std::size_t GetWorkgroupSize() const = 0;
    virtual std::size_t GetMaxNumWorkitems() const = 0;
    virtual std::size_t QueueSize() const = 0;
    virtual bool MorePhotonsAvailable() const = 0;
    virtual ConversionResult_t GetConversionResult() = 0;
    virtual std::map<std::string, double> GetStatistics() const { return std::map<std::string, double>(); }
};
CLSIM_POINTER_TYPEDEFS(I3CLSimStepToPhotonConverter);

#endif // CLSIM_COMPAT_H_INCLUDED
