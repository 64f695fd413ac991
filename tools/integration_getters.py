#!/usr/bin/env python
"""INTEGRATION.md section 2 in executable form: copies the reference's public headers (public/clsim) to a scratch directory and
adds the one-line getters the CUDA converter class needs -- "the only edits to existing reference classes".  Nothing is written
into /root/reference, nothing of the reference is stored in this repository (the patched copy is an intermediate of the build
of oracle/_ref/libclsim_icetray_mode.so and is removed again).

usage: python tools/integration_getters.py <reference root> <scratch dir>      -> <scratch dir>/public/clsim/...
Each entry: header, the line after which the getters go (the class's destructor declaration, in its public section), the getters."""
import os
import shutil
import sys

GETTERS = [
    ("function/I3CLSimFunctionRefIndexIceCube.h", "virtual ~I3CLSimFunctionRefIndexIceCube();", [
        "inline const std::string &GetMode() const {return mode_;}",
        "inline double GetPhaseCoefficient(int i) const {const double n[5] = {n0_, n1_, n2_, n3_, n4_}; return n[i];}",
        "inline double GetGroupCoefficient(int i) const {const double g[5] = {g0_, g1_, g2_, g3_, g4_}; return g[i];}"]),
    ("random_value/I3CLSimRandomValueHenyeyGreenstein.h", "virtual ~I3CLSimRandomValueHenyeyGreenstein();", [
        "inline double GetMeanCosine() const {return meanCosine_;}"]),
    ("random_value/I3CLSimRandomValueSimplifiedLiu.h", "virtual ~I3CLSimRandomValueSimplifiedLiu();", [
        "inline double GetMeanCosine() const {return meanCosine_;}"]),
    ("random_value/I3CLSimRandomValueMixed.h", "virtual ~I3CLSimRandomValueMixed();", [
        "inline double GetFractionOfFirstDistribution() const {return fractionOfFirstDistribution_;}",
        "inline I3CLSimRandomValueConstPtr GetFirstDistribution() const {return firstDistribution_;}",
        "inline I3CLSimRandomValueConstPtr GetSecondDistribution() const {return secondDistribution_;}"]),
    ("random_value/I3CLSimRandomValueInterpolatedDistribution.h", "virtual ~I3CLSimRandomValueInterpolatedDistribution();", [
        "inline bool GetConstantXSpacing() const {return constantXSpacing_ == constantXSpacing_;}   // (NaN: tabulated x values)",
        "inline double GetFirstX() const {return firstX_;}",
        "inline double GetXSpacing() const {return constantXSpacing_;}",
        "inline const std::vector<double> &GetX() const {return x_;}",
        "inline const std::vector<double> &GetY() const {return y_;}"]),
    ("random_value/I3CLSimRandomValueWlenCherenkovNoDispersion.h", "virtual ~I3CLSimRandomValueWlenCherenkovNoDispersion();", [
        "inline double GetFromWlen() const {return fromWlen_;}",
        "inline double GetToWlen() const {return toWlen_;}"]),
    ("random_value/I3CLSimRandomValueConstant.h", "virtual ~I3CLSimRandomValueConstant();", [
        "inline double GetValue() const {return value_;}"]),
    ("function/I3CLSimScalarFieldIceTiltZShift.h", "virtual ~I3CLSimScalarFieldIceTiltZShift();", [
        "inline const std::vector<double> &GetDistancesFromOriginAlongTilt() const {return distancesFromOriginAlongTilt_;}",
        "inline const std::vector<double> &GetZCoordinates() const {return zCoordinates_;}",
        "inline const I3Matrix &GetZCorrections() const {return zCorrections_;}",
        "inline double GetDirectionOfTiltAzimuth() const {return directionOfTiltAzimuth_;}",
        "inline double GetFirstZCoordinate() const {return firstZCoordinate_;}",
        "inline double GetZCoordinateSpacing() const {return zCoordinateSpacing_;}"]),
    ("function/I3CLSimScalarFieldAnisotropyAbsLenScaling.h", "virtual ~I3CLSimScalarFieldAnisotropyAbsLenScaling();", [
        "inline double GetAnisotropyDirAzimuth() const {return anisotropyDirAzimuth_;}",
        "inline double GetMagnitudeAlongDir() const {return magnitudeAlongDir_;}",
        "inline double GetMagnitudePerpToDir() const {return magnitudePerpToDir_;}"]),
    ("function/I3CLSimVectorTransformMatrix.h", "virtual ~I3CLSimVectorTransformMatrix();", [
        "inline double GetMatrixElement(std::size_t r, std::size_t c) const {return matrix_(r, c);}",
        "inline bool GetRenormalize() const {return renormalize_;}"]),
]


def main():
    reference, scratch = sys.argv[1], sys.argv[2]
    dst = os.path.join(scratch, "public", "clsim")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(os.path.join(reference, "public", "clsim"), dst)
    added = 0
    for header, anchor, getters in GETTERS:
        path = os.path.join(dst, header)
        with open(path) as f:
            text = f.read()
        if text.count(anchor) != 1:
            raise SystemExit("%s: anchor %r found %d times" % (header, anchor, text.count(anchor)))
        text = text.replace(anchor, anchor + "\n    // -- added for the CUDA converter (INTEGRATION.md section 2)\n" + "".join("    %s\n" % g for g in getters))
        with open(path, "w") as f:
            f.write(text)
        added += len(getters)
    sys.stdout.write("integration_getters.py: %d getters added to %d headers under %s\n" % (added, len(GETTERS), dst))


if __name__ == "__main__":
    main()
