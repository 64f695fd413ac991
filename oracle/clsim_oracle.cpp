/*
 * clsim_oracle.cpp -- CPU oracle for clsim's step->photon path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  See clsim_oracle.h for scope and parity status
 * ("whole-kernel parity unpinned"; component tables pinned by reference data).
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference tree).  Arithmetic is scalar fp32 in the reference's operation order;
 * build with -ffp-contract=off so that no multiply-add is fused (the reference
 * passes -cl-mad-enable, which *permits* but does not require fusing; unfused is
 * the portable choice and is stated in DESIGN.md).  Transcendentals are libm's
 * (precise path, useNativeMath=false).
 */
#include "clsim_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_last_error;

// ---------------------------------------------------------------------------------
// Literal rounding.  The generators print doubles with 10 digits after the point in
// scientific notation and append 'f' (private/clsim/I3CLSimHelperToFloatString.h:37-60;
// the geometry and distribution writers use the same stream settings,
// I3CLSimHelperGenerateGeometrySource.cxx:616-617, 1144-1145,
// I3CLSimRandomValueInterpolatedDistribution.cxx:198-199).  The OpenCL compiler then
// parses that decimal string as a float.  Reproduce both roundings.
// ---------------------------------------------------------------------------------
float lit(double v)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.10e", v);
    return std::strtof(buf, nullptr);
}

// ---------------------------------------------------------------------------------
// R1: MWC RNG (resources/kernels/mwcrng_kernel.cl:12-28)
// ---------------------------------------------------------------------------------
struct Rng {
    uint64_t x;
    uint32_t a;
    uint64_t draws;
};

// convert_float_rtz(uint): round toward zero to 24 significant bits
inline float uint_to_float_rtz(uint32_t u)
{
    if (u == 0) return 0.0f;
    int top = 31 - __builtin_clz(u);
    if (top > 23) u &= ~((1u << (top - 23)) - 1u);
    return static_cast<float>(u); // exact now
}

inline float rand_co(Rng &r)
{
    r.x = (r.x & 0xffffffffull) * r.a + (r.x >> 32);
    ++r.draws;
    return uint_to_float_rtz(static_cast<uint32_t>(r.x & 0xffffffffull)) / 4294967296.0f;
}

inline float rand_oc(Rng &r) { return 1.0f - rand_co(r); }

// ---------------------------------------------------------------------------------
// R2: safe primes (private/make_safeprimes/main.cxx:32-104).  The reference uses GMP's
// probabilistic test with 100 rounds; for n < 2^64 Miller-Rabin with the first twelve
// prime bases is deterministic.
// ---------------------------------------------------------------------------------
inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m)
{
    return static_cast<uint64_t>((static_cast<unsigned __int128>(a) * b) % m);
}
inline uint64_t powmod(uint64_t b, uint64_t e, uint64_t m)
{
    uint64_t r = 1;
    b %= m;
    while (e) {
        if (e & 1) r = mulmod(r, b, m);
        b = mulmod(b, b, m);
        e >>= 1;
    }
    return r;
}
bool is_prime_u64(uint64_t n)
{
    if (n < 2) return false;
    static const uint64_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t p : small) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    uint64_t d = n - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; ++s; }
    for (uint64_t base : small) {
        uint64_t x = powmod(base, d, n);
        if (x == 1 || x == n - 1) continue;
        bool composite = true;
        for (int i = 1; i < s; ++i) {
            x = mulmod(x, x, n);
            if (x == n - 1) { composite = false; break; }
        }
        if (composite) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------
// Wavelength generators (R3a)
// ---------------------------------------------------------------------------------
struct WlenGen {
    int kind = 0;
    int n = 0;
    float x0 = 0, dx = 0;          // equal spacing
    std::vector<float> xs;         // unequal spacing: distXValues
    std::vector<float> beta;       // distYValues
    std::vector<float> acu;        // distYCumulativeValues
    float min_val = 0, range = 0;  // no dispersion
    float value = 0;               // constant
};

// I3CLSimRandomValueInterpolatedDistribution::InitTables (…InterpolatedDistribution.cxx:137-175)
// and WriteTableCode (:177-230)
void init_interp_tables(const oracle_wlen_generator &g, WlenGen &out)
{
    const int n = g.n;
    if (n <= 1) throw std::runtime_error("wavelength generator needs at least two entries");
    std::vector<double> acu(n), beta(n);
    acu[0] = 0.;
    if (g.kind == 1) {
        for (int j = 1; j < n; ++j) acu[j] = acu[j - 1] + (g.x[j] - g.x[j - 1]) * (g.y[j] + g.y[j - 1]) / 2.;
    } else {
        for (int j = 1; j < n; ++j) acu[j] = acu[j - 1] + (g.dx) * (g.y[j] + g.y[j - 1]) / 2.;
    }
    const double norm = acu[n - 1];
    for (int j = 0; j < n; ++j) {
        beta[j] = g.y[j] / norm;
        acu[j] = acu[j] / norm;
    }
    out.n = n;
    out.beta.resize(n);
    out.acu.resize(n);
    for (int j = 0; j < n; ++j) {
        out.beta[j] = lit(beta[j]);
        out.acu[j] = lit(acu[j]);
    }
    if (g.kind == 1) {
        out.xs.resize(n);
        for (int j = 0; j < n; ++j) out.xs[j] = lit(g.x[j]);
    } else {
        out.x0 = lit(g.x0);
        out.dx = lit(g.dx);
    }
}

// device body of I3CLSimRandomValueInterpolatedDistribution (…InterpolatedDistribution.cxx:236-337)
float sample_interp(const WlenGen &g, Rng &rng)
{
    const float randomNumber = rand_oc(rng);
    unsigned int k = 0;
    float this_acu = 0.f;
    for (;;) {
        float next_acu = g.acu[k + 1];
        if (next_acu >= randomNumber) break;
        this_acu = next_acu;
        ++k;
        if (static_cast<int>(k) + 1 >= g.n) { --k; break; } // cannot happen: acu[n-1]==1 >= r (kept as a guard against reading past the table)
    }
    const float b = g.beta[k];
    float x0, slope;
    if (g.kind == 1) {
        x0 = g.xs[k];
        slope = (g.beta[k + 1] - b) / (g.xs[k + 1] - x0);
    } else {
        x0 = static_cast<float>(k) * (g.dx) + (g.x0); // convert_float_rtz(k): exact for small k
        slope = (g.beta[k + 1] - b) / (g.dx);
    }
    const float dy = randomNumber - this_acu;
    if ((b == 0.f) && (slope == 0.f)) {
        return x0;
    } else if (b == 0.f) {
        return x0 + std::sqrt(2.f * dy / slope);
    } else if (slope == 0.f) {
        return x0 + dy / b;
    } else {
        return x0 + (std::sqrt(dy * (2.f * slope) / (b * b) + 1.f) - 1.f) * b / slope;
    }
}

float sample_wlen(const WlenGen &g, Rng &rng)
{
    switch (g.kind) {
    case 0:
    case 1:
        return sample_interp(g, rng);
    case 2: {
        // I3CLSimRandomValueWlenCherenkovNoDispersion.cxx:76-97
        const float r = rand_oc(rng);
        return 1.f / (g.min_val + r * g.range);
    }
    default:
        // I3CLSimRandomValueConstant.cxx:75-100: no random number consumed
        return g.value;
    }
}

// ---------------------------------------------------------------------------------
// Geometry tables (R7), restating write_geometry_code_and_fill_buffer
// (private/opencl/I3CLSimHelperGenerateGeometrySource.cxx:712-1275) and
// generate_get_dom_position_code (:499-709).
// ---------------------------------------------------------------------------------
struct DomRec {
    unsigned int domID;
    double x, y, z;
};
struct StringRec {
    int stringID = 0;
    double meanX = 0, meanY = 0, maxZ = NAN, minZ = NAN, meandZ = 0, maxR = NAN;
    std::vector<DomRec> doms;
    unsigned short subdet = 0;
};

struct CellGrid {
    int numX = 1, numY = 1;
    float startX = 0, startY = 0, widthX = 0, widthY = 0;
    std::vector<unsigned short> index; // [y*numX + x]
};

struct GeoTables {
    int numStrings = 0;
    float omRadius = 0;
    float stringMaxRadius = 0;
    std::vector<float> stringPosX, stringPosY, stringMinZ, stringMaxZ, stringRadius;
    std::vector<unsigned char> stringInSet;
    int numSets = 0, maxLayers = 0;
    std::vector<unsigned short> layerNum;
    std::vector<float> layerStartZ, layerHeight;
    std::vector<unsigned short> layerToOM; // [set*maxLayers + layer], padded like the device buffer
    std::vector<CellGrid> cells;           // per subdetector
    // DOM position templates
    int maxDomIndex = 0;
    float mulX = 0, mulY = 0;
    std::vector<short> tmplX, tmplY;
    std::vector<float> tmplZ;
    std::vector<unsigned int> stringTmplStart;
    std::vector<float> stringMeanX, stringMeanY;
    // host-side inverse maps (…GeometrySource.cxx:1114-1132)
    std::vector<int> stringIndexToID;
    std::vector<std::vector<unsigned int>> domIndexToID;
};

// divideIntoCells (…GeometrySource.cxx:135-271)
bool divide_into_cells(const std::vector<StringRec> &strings, int subdet, double &startX, double &startY,
                       double &widthX, double &widthY, unsigned int numX, unsigned int numY,
                       std::vector<unsigned short> &cellToString)
{
    cellToString.assign(numX * numY, 0xFFFF);
    size_t count = 0;
    double minX = NAN, minY = NAN, maxX = NAN, maxY = NAN;
    for (const StringRec &s : strings) {
        if (s.subdet != static_cast<unsigned short>(subdet)) continue;
        ++count;
        if ((s.meanX - s.maxR < minX) || std::isnan(minX)) minX = s.meanX - s.maxR;
        if ((s.meanY - s.maxR < minY) || std::isnan(minY)) minY = s.meanY - s.maxR;
        if ((s.meanX + s.maxR > maxX) || std::isnan(maxX)) maxX = s.meanX + s.maxR;
        if ((s.meanY + s.maxR > maxY) || std::isnan(maxY)) maxY = s.meanY + s.maxR;
    }
    if (count == 0) throw std::runtime_error("no strings found");
    startX = minX;
    startY = minY;
    widthX = (maxX - minX) / static_cast<double>(numX);
    widthY = (maxY - minY) / static_cast<double>(numY);
    for (unsigned int i = 0; i < numX; ++i) {
        for (unsigned int j = 0; j < numY; ++j) {
            const double cxmin = startX + static_cast<double>(i) * widthX;
            const double cxmax = startX + static_cast<double>(i + 1) * widthX;
            const double cymin = startY + static_cast<double>(j) * widthY;
            const double cymax = startY + static_cast<double>(j + 1) * widthY;
            bool found = false;
            unsigned long foundNum = 0xFFFF;
            for (unsigned long t = 0; t < strings.size(); ++t) {
                const StringRec &s = strings[t];
                if (s.subdet != static_cast<unsigned short>(subdet)) continue;
                bool inX = false, inY = false;
                if ((s.meanX - s.maxR <= cxmin) && (s.meanX + s.maxR >= cxmin)) inX = true;
                if ((s.meanX - s.maxR <= cxmax) && (s.meanX + s.maxR >= cxmax)) inX = true;
                if ((s.meanX - s.maxR >= cxmin) && (s.meanX + s.maxR <= cxmax)) inX = true;
                if ((s.meanY - s.maxR <= cymin) && (s.meanY + s.maxR >= cymin)) inY = true;
                if ((s.meanY - s.maxR <= cymax) && (s.meanY + s.maxR >= cymax)) inY = true;
                if ((s.meanY - s.maxR >= cymin) && (s.meanY + s.maxR <= cymax)) inY = true;
                if (inX && inY) {
                    if (found) return false;
                    found = true;
                    foundNum = t;
                }
            }
            cellToString[j * numX + i] = found ? static_cast<unsigned short>(foundNum) : 0xFFFF;
        }
    }
    return true;
}

inline bool layer_contains_dom(double domZ, double r, double zmin, double zmax)
{
    bool c = false;
    if ((domZ - r <= zmin) && (domZ + r >= zmin)) c = true;
    if ((domZ - r <= zmax) && (domZ + r >= zmax)) c = true;
    if ((domZ - r >= zmin) && (domZ + r <= zmax)) c = true;
    return c;
}

// doesMatchLayering (…GeometrySource.cxx:273-342)
bool does_match_layering(const StringRec &s, double startZ, double height, unsigned int num, double r,
                         const std::vector<unsigned short> &layerToOM)
{
    if (num == 0) return false;
    if (r < 0.) return false;
    size_t assigned = 0;
    for (unsigned int i = 0; i < num; ++i) {
        const double zmin = startZ + static_cast<double>(i) * height;
        const double zmax = startZ + static_cast<double>(i + 1) * height;
        unsigned short should = 0xFFFF;
        for (unsigned long d = 0; d < s.doms.size(); ++d) {
            if (layer_contains_dom(s.doms[d].z, r, zmin, zmax)) {
                if (should != 0xFFFF) return false;
                should = static_cast<unsigned short>(d);
                ++assigned;
            }
        }
        if (layerToOM[i] != should) return false;
    }
    if (assigned != s.doms.size()) return false;
    return true;
}

// divideIntoLayers (…GeometrySource.cxx:375-446)
bool divide_into_layers(const StringRec &s, double &startZ, double &height, unsigned int num, double r,
                        double minZHint, double maxZHint, std::vector<unsigned short> &layerToOM)
{
    if (num == 0) return false;
    if (r < 0.) return false;
    layerToOM.assign(num, 0xFFFF);
    double minZ = minZHint, maxZ = maxZHint;
    if ((s.minZ - r < minZ) || std::isnan(minZ)) minZ = s.minZ - r;
    if ((s.maxZ + r > maxZ) || std::isnan(maxZ)) maxZ = s.maxZ + r;
    startZ = minZ;
    height = (maxZ - minZ) / static_cast<double>(num);
    for (unsigned int i = 0; i < num; ++i) {
        const double zmin = startZ + static_cast<double>(i) * height;
        const double zmax = startZ + static_cast<double>(i + 1) * height;
        for (unsigned long d = 0; d < s.doms.size(); ++d) {
            if (layer_contains_dom(s.doms[d].z, r, zmin, zmax)) {
                if (layerToOM[i] != 0xFFFF) return false;
                layerToOM[i] = static_cast<unsigned short>(d);
            }
        }
    }
    return true;
}

void build_geometry(const oracle_geometry &g, GeoTables &t)
{
    const size_t n = static_cast<size_t>(g.num_doms);
    if (n == 0) throw std::runtime_error("Empty geometry provided.");
    const double omRadius = g.om_radius;
    if (omRadius < 0.) throw std::runtime_error("Zero or negative OM radius.");

    // (stringID, subdetector) pairs in set order (:737-744)
    std::set<std::pair<int, int>> stringSet;
    std::set<int> subdetSet;
    for (size_t i = 0; i < n; ++i) {
        stringSet.insert(std::make_pair(g.string_id[i], g.subdetector[i]));
        subdetSet.insert(g.subdetector[i]);
    }
    std::map<int, unsigned short> subdetIndex;
    {
        unsigned short k = 0;
        for (int s : subdetSet) subdetIndex[s] = k++;
    }
    const unsigned short numSubdet = static_cast<unsigned short>(subdetSet.size());
    if (numSubdet > 9) throw std::runtime_error("more than 9 subdetectors are currently not supported.");

    std::vector<StringRec> strings(stringSet.size());
    double stringMaxR = NAN;
    unsigned int si = 0;
    for (const auto &pr : stringSet) { // :779-882
        StringRec &cs = strings[si];
        cs.stringID = pr.first;
        cs.subdet = subdetIndex[pr.second];
        unsigned long numDoms = 0;
        double lastZ = NAN, lastdZ = NAN, meandZ = 0.;
        unsigned int numdZ = 0;
        for (size_t i = 0; i < n; ++i) {
            if (g.string_id[i] != pr.first) continue;
            if (g.subdetector[i] != pr.second) continue;
            cs.meanX += g.x[i];
            cs.meanY += g.y[i];
            if ((g.z[i] > cs.maxZ) || std::isnan(cs.maxZ)) cs.maxZ = g.z[i];
            if ((g.z[i] < cs.minZ) || std::isnan(cs.minZ)) cs.minZ = g.z[i];
            if (std::isnan(lastZ)) {
                lastZ = g.z[i];
            } else {
                double dZ = std::abs(lastZ - g.z[i]);
                lastZ = g.z[i];
                if (!std::isnan(lastdZ)) {
                    if (dZ < 1.75 * meandZ / static_cast<double>(numdZ)) {
                        meandZ += dZ;
                        numdZ++;
                        lastdZ = dZ;
                    }
                } else {
                    lastdZ = dZ;
                    meandZ += dZ;
                    numdZ++;
                }
            }
            cs.doms.push_back(DomRec{g.dom_id[i], g.x[i], g.y[i], g.z[i]});
            ++numDoms;
        }
        cs.meanX /= static_cast<double>(numDoms);
        cs.meanY /= static_cast<double>(numDoms);
        meandZ /= static_cast<double>(numdZ);
        cs.meandZ = meandZ;
        for (size_t i = 0; i < n; ++i) {
            if (g.string_id[i] != pr.first) continue;
            if (g.subdetector[i] != pr.second) continue;
            const double dX = cs.meanX - g.x[i];
            const double dY = cs.meanY - g.y[i];
            const double thisR = std::sqrt(dX * dX + dY * dY) + omRadius;
            if ((thisR > cs.maxR) || std::isnan(cs.maxR)) cs.maxR = thisR;
            if ((thisR > stringMaxR) || std::isnan(stringMaxR)) stringMaxR = thisR;
        }
        ++si;
    }

    // xy cells per subdetector (:905-949)
    t.cells.resize(numSubdet);
    for (unsigned short sd = 0; sd < numSubdet; ++sd) {
        unsigned int nx = 1, ny = 1;
        double sx, sy, wx, wy;
        std::vector<unsigned short> c2s;
        for (;;) {
            if (divide_into_cells(strings, sd, sx, sy, wx, wy, nx, ny, c2s)) break;
            ++nx;
            ++ny;
            if (nx >= 1000) throw std::runtime_error("Could not generate a x-y cell division");
        }
        CellGrid &cg = t.cells[sd];
        cg.numX = static_cast<int>(nx);
        cg.numY = static_cast<int>(ny);
        cg.startX = lit(sx);
        cg.startY = lit(sy);
        cg.widthX = lit(wx);
        cg.widthY = lit(wy);
        cg.index = c2s;
    }

    // z layers / string sets (:956-1082)
    unsigned int numSets = 0;
    std::vector<unsigned int> geoLayerNum;
    std::vector<double> layerStartZ, layerHeight;
    std::vector<std::vector<unsigned short>> perSet;
    std::vector<unsigned char> inSet(strings.size());
    unsigned int maxLayerNum = 0;
    for (unsigned int s = 0; s < strings.size(); ++s) {
        bool match = false;
        unsigned int existing = 0;
        for (unsigned int k = 0; k < numSets; ++k) {
            if (does_match_layering(strings[s], layerStartZ[k], layerHeight[k], geoLayerNum[k], omRadius, perSet[k])) {
                existing = k;
                match = true;
                break;
            }
        }
        if (match) {
            inSet[s] = static_cast<unsigned char>(existing);
            continue;
        }
        inSet[s] = static_cast<unsigned char>(numSets);
        ++numSets;
        if (numSets >= 0xFF) throw std::runtime_error("Not more than 255 string sets are supported!");
        layerStartZ.push_back(NAN);
        layerHeight.push_back(NAN);
        geoLayerNum.push_back(1);
        perSet.push_back(std::vector<unsigned short>());
        const StringRec &cs = strings[s];
        const double lo = cs.minZ - cs.meandZ / 2., hi = cs.maxZ + cs.meandZ / 2.;
        geoLayerNum.back() = static_cast<unsigned int>((cs.maxZ - cs.minZ + cs.meandZ) / cs.meandZ);
        bool ok = divide_into_layers(cs, layerStartZ.back(), layerHeight.back(), geoLayerNum.back(), omRadius, lo, hi, perSet.back());
        if (!ok) {
            geoLayerNum.back() = static_cast<unsigned int>((cs.maxZ - cs.minZ + cs.meandZ) / cs.meandZ) + 1;
            ok = divide_into_layers(cs, layerStartZ.back(), layerHeight.back(), geoLayerNum.back(), omRadius, lo, hi, perSet.back());
        }
        if (!ok) {
            geoLayerNum.back() = 1;
            for (;;) {
                ok = divide_into_layers(cs, layerStartZ.back(), layerHeight.back(), geoLayerNum.back(), omRadius, lo, hi, perSet.back());
                if (ok) break;
                ++geoLayerNum.back();
                if (geoLayerNum.back() >= 1000) throw std::runtime_error("no possible layer division for a string");
            }
        }
        if (geoLayerNum.back() > maxLayerNum) maxLayerNum = geoLayerNum.back();
    }

    // flat layer->OM table padded to a multiple of 64 (:1084-1112)
    unsigned int bufSize = ((numSets * maxLayerNum) / 64) + 1;
    bufSize *= 64;
    t.layerToOM.assign(bufSize, 0xFFFF);
    for (unsigned int j = 0; j < numSets; ++j)
        for (unsigned int i = 0; i < geoLayerNum[j]; ++i) t.layerToOM[j * maxLayerNum + i] = perSet[j][i];

    t.numStrings = static_cast<int>(strings.size());
    t.omRadius = lit(omRadius);
    t.stringMaxRadius = lit(stringMaxR);
    t.numSets = static_cast<int>(numSets);
    t.maxLayers = static_cast<int>(maxLayerNum);
    for (const StringRec &s : strings) {
        t.stringPosX.push_back(lit(s.meanX));
        t.stringPosY.push_back(lit(s.meanY));
        t.stringRadius.push_back(lit(s.maxR));
        t.stringMinZ.push_back(lit(s.minZ));
        t.stringMaxZ.push_back(lit(s.maxZ));
        t.stringIndexToID.push_back(s.stringID);
        std::vector<unsigned int> ids;
        for (const DomRec &d : s.doms) ids.push_back(d.domID);
        t.domIndexToID.push_back(ids);
    }
    t.stringInSet = inSet;
    for (unsigned int k = 0; k < numSets; ++k) {
        t.layerNum.push_back(static_cast<unsigned short>(geoLayerNum[k]));
        t.layerStartZ.push_back(lit(layerStartZ[k]));
        t.layerHeight.push_back(lit(layerHeight[k]));
    }

    // DOM position templates (:499-709)
    const size_t ns = strings.size();
    std::vector<double> meanX(ns, 0.), meanY(ns, 0.);
    size_t maxNumDoms = 0;
    for (size_t i = 0; i < ns; ++i) {
        if (strings[i].doms.size() > maxNumDoms) maxNumDoms = strings[i].doms.size();
        for (const DomRec &d : strings[i].doms) {
            meanX[i] += d.x;
            meanY[i] += d.y;
        }
        meanX[i] /= static_cast<double>(strings[i].doms.size());
        meanY[i] /= static_cast<double>(strings[i].doms.size());
    }
    std::vector<size_t> inTemplate(ns);
    std::vector<std::vector<double>> tx, ty, tz;
    const double eps = 1e-1 * 1e-3; // 1e-1*I3Units::mm
    for (size_t i = 0; i < ns; ++i) {
        bool found = false;
        for (size_t k = 0; k < tx.size() && !found; ++k) { // isStringInTemplate (:449-495)
            if (strings[i].doms.size() != tx[k].size()) continue;
            bool m = true;
            for (size_t j = 0; j < strings[i].doms.size(); ++j) {
                const DomRec &d = strings[i].doms[j];
                if (std::abs(tx[k][j] - (d.x - meanX[i])) > eps) { m = false; break; }
                if (std::abs(ty[k][j] - (d.y - meanY[i])) > eps) { m = false; break; }
                if (std::abs(tz[k][j] - (d.z)) > eps) { m = false; break; }
            }
            if (m) {
                inTemplate[i] = k;
                found = true;
            }
        }
        if (found) continue;
        tx.emplace_back();
        ty.emplace_back();
        tz.emplace_back();
        for (const DomRec &d : strings[i].doms) {
            tx.back().push_back(d.x - meanX[i]);
            ty.back().push_back(d.y - meanY[i]);
            tz.back().push_back(d.z);
        }
        inTemplate[i] = tx.size() - 1;
    }
    double maxAbsX = NAN, maxAbsY = NAN;
    std::vector<double> fx, fy, fz;
    std::vector<size_t> tmplStart(tx.size());
    for (size_t k = 0; k < tx.size(); ++k) {
        tmplStart[k] = fx.size();
        for (size_t j = 0; j < tx[k].size(); ++j) {
            fx.push_back(tx[k][j]);
            fy.push_back(ty[k][j]);
            fz.push_back(tz[k][j]);
            const double ax = std::abs(tx[k][j]), ay = std::abs(ty[k][j]);
            if ((ax > maxAbsX) || std::isnan(maxAbsX)) maxAbsX = ax;
            if ((ay > maxAbsY) || std::isnan(maxAbsY)) maxAbsY = ay;
        }
    }
    t.maxDomIndex = static_cast<int>(maxNumDoms);
    t.mulX = lit(maxAbsX / 32767.);
    t.mulY = lit(maxAbsY / 32767.);
    for (size_t i = 0; i < fx.size(); ++i) {
        // static_cast<short>(value/(maxAbs/32767.)) (:641, :651).  For perfectly straight
        // strings maxAbs==0 and the quotient is NaN; the cast is then undefined in C++.  The
        // multiplier is 0 in that case so any finite short yields the string mean; use 0.
        const double qx = fx[i] / (maxAbsX / 32767.);
        const double qy = fy[i] / (maxAbsY / 32767.);
        t.tmplX.push_back(std::isfinite(qx) ? static_cast<short>(qx) : static_cast<short>(0));
        t.tmplY.push_back(std::isfinite(qy) ? static_cast<short>(qy) : static_cast<short>(0));
        t.tmplZ.push_back(lit(fz[i]));
    }
    for (size_t i = 0; i < ns; ++i) {
        t.stringTmplStart.push_back(static_cast<unsigned int>(tmplStart[inTemplate[i]]));
        t.stringMeanX.push_back(lit(meanX[i]));
        t.stringMeanY.push_back(lit(meanY[i]));
    }
}

// ---------------------------------------------------------------------------------
// Medium (R4, R4a, R4b, R9)
// ---------------------------------------------------------------------------------
struct Medium {
    int L = 0;
    float z0 = 0, h = 0;
    float kappa = 0, A = 0, B = 0, D = 0, E = 0;
    std::vector<float> aDust, dTau;
    float alpha = 0, refWlenRecip = 0;
    std::vector<float> b400;
    float n[5], g[5];
    float cLight = 0;
    int scatKind = 0;
    float fsl = 0, oneMinusFsl = 0, gg = 0, gg2 = 0, slBeta = 0;
    // tilt
    int tiltND = 0, tiltNZ = 0;
    std::vector<float> tiltDist, tiltCorr;
    float tiltZ0 = 0, tiltDZ = 0, tiltLnx = 0, tiltLny = 0;
    // anisotropy
    bool aniso = false;
    float l[3], rl[3], azx = 0, azy = 0, mazy = 0, B2 = 0;
    float pre[9], post[9];
    bool preRenorm = false, postRenorm = false;
};

struct Bias {
    int kind = 0;
    int n = 0;
    float x0 = 0, dx = 0, value = 1;
    std::vector<float> v;
};

} // namespace

struct oracle_scene {
    Medium med;
    std::vector<WlenGen> gens;
    Bias bias;
    GeoTables geo;
    bool haveGeo = false;
    bool stop = true, saveAll = false, fixedAbs = false, pancake = false;
    float prescale = 0, fixedAbsLens = 0, pancakeFactor = 1;
    int history = 0;
};

namespace {

void build_medium(const oracle_medium &m, Medium &o)
{
    o.L = m.num_layers;
    if (o.L < 1) throw std::runtime_error("medium needs at least one layer");
    o.z0 = lit(m.layers_zstart);
    o.h = lit(m.layers_height);
    o.kappa = lit(m.kappa);
    o.A = lit(m.A);
    o.B = lit(m.B);
    o.D = lit(m.D);
    o.E = lit(m.E);
    o.alpha = lit(m.alpha);
    o.refWlenRecip = lit(1. / (400. * 1e-9));
    for (int i = 0; i < o.L; ++i) {
        o.aDust.push_back(lit(m.a_dust400[i]));
        o.dTau.push_back(lit(m.delta_tau[i]));
        o.b400.push_back(lit(m.b400[i]));
    }
    for (int i = 0; i < 5; ++i) {
        o.n[i] = lit(m.n_phase[i]);
        o.g[i] = lit(m.n_group[i]);
    }
    o.cLight = lit(0.299792458); // I3Constants::c in m/ns
    o.scatKind = m.scat_kind;
    o.fsl = lit(m.f_sl);
    o.oneMinusFsl = lit(1. - m.f_sl);
    o.gg = lit(m.mean_cos);
    o.gg2 = lit(m.mean_cos * m.mean_cos);
    o.slBeta = lit((1. - m.mean_cos) / (1. + m.mean_cos));
    o.tiltND = m.tilt_num_dist;
    o.tiltNZ = m.tilt_num_z;
    if (o.tiltND > 0) {
        if (o.tiltND < 2 || o.tiltNZ < 2) throw std::runtime_error("tilt table needs at least 2x2 entries");
        for (int i = 0; i < o.tiltND; ++i) o.tiltDist.push_back(lit(m.tilt_dist[i]));
        for (int i = 0; i < o.tiltND * o.tiltNZ; ++i) o.tiltCorr.push_back(lit(m.tilt_corr[i]));
        o.tiltZ0 = lit(m.tilt_z0);
        o.tiltDZ = lit(m.tilt_dz);
        o.tiltLnx = lit(std::cos(m.tilt_azimuth));
        o.tiltLny = lit(std::sin(m.tilt_azimuth));
    }
    o.aniso = m.has_anisotropy != 0;
    if (o.aniso) {
        // I3CLSimScalarFieldAnisotropyAbsLenScaling.cxx:92-134
        const double azx = std::cos(m.aniso_azimuth), azy = std::sin(m.aniso_azimuth);
        const double k1 = std::exp(m.aniso_along), k2 = std::exp(m.aniso_perp), kz = 1. / (k1 * k2);
        const double l1 = k1 * k1, l2 = k2 * k2, l3 = kz * kz;
        const double B2 = 1. / l1 + 1. / l2 + 1. / l3;
        o.l[0] = lit(l1); o.l[1] = lit(l2); o.l[2] = lit(l3);
        o.rl[0] = lit(1. / l1); o.rl[1] = lit(1. / l2); o.rl[2] = lit(1. / l3);
        o.azx = lit(azx); o.azy = lit(azy); o.mazy = lit(-azy);
        o.B2 = lit(B2);
        for (int i = 0; i < 9; ++i) {
            o.pre[i] = lit(m.pre_matrix[i]);
            o.post[i] = lit(m.post_matrix[i]);
        }
        o.preRenorm = m.pre_renormalize != 0;
        o.postRenorm = m.post_renormalize != 0;
    }
}

// I3CLSimFunctionRefIndexIceCube.cxx:128-180 (device text)
inline float phase_ref_index(const Medium &m, float wlen)
{
    const float x = wlen / 1e-6f;
    return m.n[0] + x * (m.n[1] + x * (m.n[2] + x * (m.n[3] + x * m.n[4])));
}
inline float group_ref_index(const Medium &m, float wlen)
{
    const float x = wlen / 1e-6f;
    const float np = m.n[0] + x * (m.n[1] + x * (m.n[2] + x * (m.n[3] + x * m.n[4])));
    const float np_corr = m.g[0] + x * (m.g[1] + x * (m.g[2] + x * (m.g[3] + x * m.g[4])));
    return np * np_corr;
}
// I3CLSimHelperGenerateMediumPropertiesSource.cxx:256-272
inline float group_velocity(const Medium &m, float wlen) { return m.cLight / group_ref_index(m, wlen); }

// …_Optimizers.cxx:195-250 (layered) == I3CLSimFunctionScatLenIceCube.cxx:62-85 (single layer)
inline float scattering_length(const Medium &m, int layer, float wlen)
{
    return 1.f / (m.b400[layer] * std::pow(wlen * m.refWlenRecip, -m.alpha));
}
// …_Optimizers.cxx:123-190 (layered) == I3CLSimFunctionAbsLenIceCube.cxx:70-97 (single layer)
inline float absorption_length(const Medium &m, int layer, float wlen)
{
    const float x = wlen / 1e-9f;
    return 1.f / ((m.D * m.aDust[layer] + m.E) * std::pow(x, -m.kappa) + m.A * std::exp(-m.B / x) * (1.f + 0.01f * m.dTau[layer]));
}

// I3CLSimFunctionFromTable.cxx:241-305 / I3CLSimFunctionConstant.cxx:80-99
inline float wavelength_bias(const Bias &b, float wavelength)
{
    if (b.kind == 0) return b.value;
    float fbin;
    float fraction = std::modf((wavelength - b.x0) / b.dx, &fbin);
    int ibin = static_cast<int>(fbin);
    if ((ibin < 0) || ((ibin == 0) && (fraction < 0))) {
        ibin = 0;
        fraction = 0.f;
    } else if (ibin >= b.n - 1) {
        ibin = b.n - 2;
        fraction = 1.f;
    }
    return b.v[ibin] + (b.v[ibin + 1] - b.v[ibin]) * fraction; // mix()
}

// I3CLSimScalarFieldIceTiltZShift.cxx:145-216
inline float tilt_z_shift(const Medium &m, float x, float y, float z)
{
    if (m.tiltND == 0) return 0.f;
    const float z_rescaled = (z - m.tiltZ0) / m.tiltDZ;
    const int k = std::min(std::max(static_cast<int>(std::floor(z_rescaled)), 0), m.tiltNZ - 2);
    const float fraction_z_above = z_rescaled - static_cast<float>(k);
    const float fraction_z_below = 1.f - fraction_z_above;
    const float nr = m.tiltLnx * x + m.tiltLny * y;
    for (int j = 1; j < m.tiltND; j++) {
        const float thisDist = m.tiltDist[j];
        if ((nr < thisDist) || (j == m.tiltND - 1)) {
            const float previousDist = m.tiltDist[j - 1];
            const float thisDistanceBinWidth = thisDist - previousDist;
            const float frac_at_lower = (thisDist - nr) / thisDistanceBinWidth;
            const float frac_at_upper = 1.f - frac_at_lower;
            const float val_at_lower = (m.tiltCorr[(j - 1) * m.tiltNZ + k + 1] * fraction_z_above + m.tiltCorr[(j - 1) * m.tiltNZ + k] * fraction_z_below);
            const float val_at_upper = (m.tiltCorr[j * m.tiltNZ + k + 1] * fraction_z_above + m.tiltCorr[j * m.tiltNZ + k] * fraction_z_below);
            return (val_at_upper * frac_at_upper + val_at_lower * frac_at_lower);
        }
    }
    return 0.f;
}

// I3CLSimScalarFieldAnisotropyAbsLenScaling.cxx:92-134 / I3CLSimScalarFieldConstant.cxx:62-78
inline float abs_len_corr_factor(const Medium &m, float dx, float dy, float dz)
{
    if (!m.aniso) return 1.f;
    const float n0 = (m.azx * dx) + (m.azy * dy);
    const float n1 = (m.mazy * dx) + (m.azx * dy);
    const float n2 = dz;
    const float s0 = n0 * n0, s1 = n1 * n1, s2 = n2 * n2;
    const float nB = ((s0 * m.rl[0] + s1 * m.rl[1]) + s2 * m.rl[2]) + 0.f;
    const float An = ((s0 * m.l[0] + s1 * m.l[1]) + s2 * m.l[2]) + 0.f;
    return 2.f / ((m.B2 - nB) * An);
}

// I3CLSimVectorTransformMatrix.cxx:101-133 (precise branch) / …Constant.cxx:60-71
inline void transform_dir(const float *M, bool renorm, float &x, float &y, float &z)
{
    const float nx = (M[0] * x) + (M[1] * y) + (M[2] * z);
    const float ny = (M[3] * x) + (M[4] * y) + (M[5] * z);
    const float nz = (M[6] * x) + (M[7] * y) + (M[8] * z);
    x = nx; y = ny; z = nz;
    if (renorm) {
        const float norm = 1.f / std::sqrt(x * x + y * y + z * z); // rsqrt()
        x = x * norm; y = y * norm; z = z * norm;
    }
}

inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }

// I3CLSimRandomValueHenyeyGreenstein.cxx:69-91
inline float hg_cos(const Medium &m, float u)
{
    const float s = 2.f * (u) - 1.f;
    const float ii = ((1.f - m.gg2) / (1.f + m.gg * s));
    return clampf((1.f + m.gg2 - ii * ii) / (2.f * m.gg), -1.f, 1.f);
}
// I3CLSimRandomValueSimplifiedLiu.cxx:64-87
inline float sl_cos(const Medium &m, float u) { return clampf(2.f * std::pow((u), m.slBeta) - 1.f, -1.f, 1.f); }
// I3CLSimRandomValueMixed.cxx:117-146 (single-random-number form)
inline float scattering_cos_angle(const Medium &m, Rng &rng)
{
    if (m.scatKind == 1) return hg_cos(m, rand_co(rng));
    if (m.scatKind == 2) return sl_cos(m, rand_co(rng));
    const float rr = rand_co(rng);
    if (rr < m.fsl) return sl_cos(m, rr / m.fsl);
    return hg_cos(m, (1.f - rr) / m.oneMinusFsl);
}

// ---------------------------------------------------------------------------------
// propagation_kernel.c.cl
// ---------------------------------------------------------------------------------
const float kSpeedOfLight = 0.299792458f;
const float kPI = 3.14159265359f;
const float kEpsilon = 0.00001f;

struct Vec4 {
    float x, y, z, w;
};

inline float sqr(float a) { return a * a; }

// propagation_kernel.c.cl:73-81
inline int find_layer(const Medium &m, float z) { return static_cast<int>((z - m.z0) / m.h); }
inline float layer_boundary(const Medium &m, int layer) { return (static_cast<float>(layer) * m.h) + m.z0; }

// OpenCL sign(): +-1, +-0 passed through, 0 for NaN
inline float signf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : ((v != v) ? 0.f : v)); }

// propagation_kernel.c.cl:83-129
inline void scatter_direction_by_angle(float cosa, float sina, Vec4 &d, float randomNumber)
{
    const float b = 2.0f * kPI * randomNumber;
    const float cosb = std::cos(b);
    const float sinb = std::sin(b);
    const float sinth = std::sqrt(std::max(0.f, 1.f - d.z * d.z));
    if (sinth > 0.f) {
        const Vec4 o = d;
        d.x = o.x * cosa - ((o.y * cosb + o.z * o.x * sinb) * sina) / sinth;
        d.y = o.y * cosa + ((o.x * cosb - o.z * o.y * sinb) * sina) / sinth;
        d.z = o.z * cosa + sina * sinb * sinth;
    } else {
        d.x = sina * cosb;
        d.y = sina * sinb;
        d.z = cosa * signf(d.z);
    }
    {
        const float recip_length = 1.f / std::sqrt(sqr(d.x) + sqr(d.y) + sqr(d.z));
        d.x *= recip_length;
        d.y *= recip_length;
        d.z *= recip_length;
    }
}

// propagation_kernel.c.cl:132-184
inline void create_photon_from_track(const oracle_scene &sc, const oracle_step &step, const Vec4 &stepDir, Rng &rng,
                                     Vec4 &pos, Vec4 &dir)
{
    const float shiftMultiplied = step.dir_and_length_and_beta[2] * rand_co(rng);
    const float inverseParticleSpeed = 1.f / (kSpeedOfLight * step.dir_and_length_and_beta[3]);
    pos.x = step.pos_and_time[0] + stepDir.x * shiftMultiplied;
    pos.y = step.pos_and_time[1] + stepDir.y * shiftMultiplied;
    pos.z = step.pos_and_time[2] + stepDir.z * shiftMultiplied;
    pos.w = step.pos_and_time[3] + inverseParticleSpeed * shiftMultiplied;
    const unsigned int layer = static_cast<unsigned int>(std::min(std::max(find_layer(sc.med, pos.z), 0), sc.med.L - 1));
    (void)layer; // the refractive index is layer independent for the supported medium
    const bool noFlasher = sc.gens.size() <= 1; // -DNO_FLASHER (…OpenCL.cxx:646-650)
    if (noFlasher || step.source_type == 0) {
        const float wavelength = sample_wlen(sc.gens[0], rng);
        const float cosCherenkov = std::min(1.f, 1.f / (step.dir_and_length_and_beta[3] * phase_ref_index(sc.med, wavelength)));
        const float sinCherenkov = std::sqrt(1.f - cosCherenkov * cosCherenkov);
        dir.x = stepDir.x; dir.y = stepDir.y; dir.z = stepDir.z;
        dir.w = wavelength;
        scatter_direction_by_angle(cosCherenkov, sinCherenkov, dir, rand_co(rng));
    } else {
        // generateWavelength(number,…) master function (…MediumPropertiesSource.cxx:394-436):
        // unknown generator numbers return 0
        const unsigned int k = step.source_type;
        const float wavelength = (k < sc.gens.size()) ? sample_wlen(sc.gens[k], rng) : 0.f;
        dir.x = stepDir.x; dir.y = stepDir.y; dir.z = stepDir.z;
        dir.w = wavelength;
    }
}

// propagation_kernel.c.cl:206-223
inline void sph_dir_from_car(const Vec4 &c, float &theta, float &phi)
{
    const float r_inv = 1.f / std::sqrt(c.x * c.x + c.y * c.y + c.z * c.z);
    theta = 0.f;
    if (std::fabs(c.z * r_inv) <= 1.f) {
        theta = std::acos(c.z * r_inv);
    } else {
        if (c.z < 0.f) theta = kPI;
    }
    if (theta < 0.f) theta += 2.f * kPI;
    phi = std::atan2(c.y, c.x);
    if (phi < 0.f) phi += 2.f * kPI;
}

// generated geometryGetDomPosition (…GeometrySource.cxx:685-700)
inline void dom_position(const GeoTables &g, unsigned short stringNum, unsigned short domNum, float &x, float &y, float &z)
{
    const unsigned int index = g.stringTmplStart[stringNum] + static_cast<unsigned int>(domNum);
    x = static_cast<float>(g.tmplX[index]) * g.mulX + g.stringMeanX[stringNum];
    y = static_cast<float>(g.tmplY[index]) * g.mulY + g.stringMeanY[stringNum];
    z = g.tmplZ[index];
}

struct HitSink {
    oracle_photon *out;
    size_t cap;
    float *history; // raw ring layout, [slot][entry][4]
    uint64_t count;
    std::vector<oracle_photon> local;   // used by the OpenMP driver
    std::vector<float> localHistory;
    bool useLocal;
};

struct PhotonState {
    Vec4 pos, dir, startPos, startDir;
    uint32_t numScatters;
    float totalPath;
    float invGroupVel;
    float absLensInitial;
};

// propagation_kernel.c.cl:307-404
void save_hit(const oracle_scene &sc, HitSink &sink, const Vec4 &pos, const Vec4 &dir, float thisStepLength, float invGroupVel,
              float totalPath, uint32_t numScatters, float distInAbsLens, const Vec4 &startPos, const Vec4 &startDir,
              const oracle_step &step, unsigned short hitOnString, unsigned short hitOnDom, const float *curHistory)
{
    oracle_photon p;
    std::memset(&p, 0, sizeof(p));
    float domX = 0.f, domY = 0.f, domZ = 0.f;
    if (!sc.saveAll) {
        dom_position(sc.geo, hitOnString, hitOnDom, domX, domY, domZ);
        if (sc.pancake) {
            const float px = pos.x - domX, py = pos.y - domY, pz = pos.z - domZ;
            const float parallel = px * dir.x + py * dir.y + pz * dir.z;
            const float nx = px - parallel * dir.x;
            const float ny = py - parallel * dir.y;
            const float nz = pz - parallel * dir.z;
            const float f = ((sc.pancakeFactor - 1.f) / sc.pancakeFactor);
            domX += f * nx;
            domY += f * ny;
            domZ += f * nz;
        }
    }
    // Deviation: in SAVE_ALL mode the reference revision references an undefined
    // geometryGetDomPosition; absolute coordinates (DOM at the origin, no pancake shift) are used.
    p.pos_and_time[0] = pos.x + thisStepLength * dir.x - domX;
    p.pos_and_time[1] = pos.y + thisStepLength * dir.y - domY;
    p.pos_and_time[2] = pos.z + thisStepLength * dir.z - domZ;
    p.pos_and_time[3] = pos.w + thisStepLength * invGroupVel;
    sph_dir_from_car(dir, p.dir[0], p.dir[1]);
    p.wavelength = dir.w;
    p.cherenkov_dist = totalPath + thisStepLength;
    p.num_scatters = numScatters;
    p.weight = step.weight / wavelength_bias(sc.bias, dir.w);
    p.identifier = step.identifier;
    if (sc.saveAll) {
        p.string_id = 0;
        p.om_id = 0;
    } else {
        // host-side index -> ID rewrite (…OpenCL.cxx:1565-1602)
        p.string_id = static_cast<int16_t>(sc.geo.stringIndexToID.at(hitOnString));
        p.om_id = static_cast<uint16_t>(sc.geo.domIndexToID.at(hitOnString).at(hitOnDom));
    }
    p.start_pos_and_time[0] = startPos.x;
    p.start_pos_and_time[1] = startPos.y;
    p.start_pos_and_time[2] = startPos.z;
    p.start_pos_and_time[3] = startPos.w;
    sph_dir_from_car(startDir, p.start_dir[0], p.start_dir[1]);
    p.group_velocity = 1.f / invGroupVel;
    p.dist_in_abs_lens = distInAbsLens;

    const uint64_t myIndex = sink.count++;
    if (sink.useLocal) {
        sink.local.push_back(p);
        if (sc.history > 0) sink.localHistory.insert(sink.localHistory.end(), curHistory, curHistory + 4 * sc.history);
    } else if (myIndex < sink.cap) {
        sink.out[myIndex] = p;
        if (sc.history > 0 && sink.history) std::memcpy(sink.history + myIndex * 4 * sc.history, curHistory, sizeof(float) * 4 * sc.history);
    }
}

struct SegmentCtx {
    const oracle_scene &sc;
    HitSink &sink;
    const oracle_step &step;
    const PhotonState &ph;
    float distInAbsLens;
    const float *curHistory;
};

// sparse_collision_kernel.c.cl:27-192
void check_on_string(SegmentCtx &c, unsigned short stringNum, float dirLenXYSqr, float *thisStepLength, bool *hitRecorded,
                     unsigned short *hitOnString, unsigned short *hitOnDom)
{
    const GeoTables &g = c.sc.geo;
    const Vec4 &pos = c.ph.pos;
    const Vec4 &dir = c.ph.dir;
    const unsigned char stringSet = g.stringInSet[stringNum];
    {
        const float smin = sqr(((pos.x - g.stringPosX[stringNum]) * dir.y - (pos.y - g.stringPosY[stringNum]) * dir.x)) / dirLenXYSqr;
        if (smin > sqr(g.stringMaxRadius)) return;
    }
    {
        if ((dir.z > 0.f) && (pos.z > g.stringMaxZ[stringNum] + g.omRadius)) return;
        if ((dir.z < 0.f) && (pos.z < g.stringMinZ[stringNum] - g.omRadius)) return;
    }
    int lowLayerZ = static_cast<int>((pos.z - g.layerStartZ[stringSet]) / g.layerHeight[stringSet]);
    int highLayerZ = static_cast<int>((pos.z + dir.z * (*thisStepLength) - g.layerStartZ[stringSet]) / g.layerHeight[stringSet]);
    if (highLayerZ < lowLayerZ) std::swap(lowLayerZ, highLayerZ);
    lowLayerZ = std::min(std::max(lowLayerZ, 0), static_cast<int>(g.layerNum[stringSet]) - 1);
    highLayerZ = std::min(std::max(highLayerZ, 0), static_cast<int>(g.layerNum[stringSet]) - 1);

    std::vector<bool> domChecked; // non-STOP de-duplication, see DESIGN.md (quirk 8: intent restated)
    if (!c.sc.stop) domChecked.assign(g.maxDomIndex + 1, false);

    const unsigned short *layerToOM = &g.layerToOM[static_cast<unsigned int>(stringSet) * g.maxLayers + lowLayerZ];
    for (int layer_z = lowLayerZ; layer_z <= highLayerZ; ++layer_z, ++layerToOM) {
        const unsigned short domNum = *layerToOM;
        if (domNum == 0xFFFF) continue;
        if (!c.sc.stop) {
            if (domChecked[domNum]) continue;
            domChecked[domNum] = true;
        }
        float domX, domY, domZ;
        dom_position(g, stringNum, domNum, domX, domY, domZ);
        float urdot, discr;
        {
            const float dx = domX - pos.x, dy = domY - pos.y, dz = domZ - pos.z;
            const float dr2 = ((dx * dx + dy * dy) + dz * dz) + 0.f;
            urdot = ((dx * dir.x + dy * dir.y) + dz * dir.z) + 0.f; // drvec.w == 0
            discr = sqr(urdot) - dr2 + g.omRadius * g.omRadius;
        }
        if (discr < 0.f) continue;
        if (c.sc.pancake) discr = std::sqrt(discr) / c.sc.pancakeFactor;
        else discr = std::sqrt(discr);
        {
            const float smin2 = urdot + discr;
            if (smin2 < 0.f) continue;
        }
        const float smin1 = urdot - discr;
        if (smin1 < 0.f) continue;
        if (smin1 < *thisStepLength) {
            if (c.sc.stop) {
                *thisStepLength = smin1;
                *hitOnString = stringNum;
                *hitOnDom = domNum;
                *hitRecorded = true;
            } else {
                save_hit(c.sc, c.sink, pos, dir, smin1, c.ph.invGroupVel, c.ph.totalPath, c.ph.numScatters, c.distInAbsLens,
                         c.ph.startPos, c.ph.startDir, c.step, stringNum, domNum, c.curHistory);
            }
        }
    }
}

// sparse_collision_kernel.c.cl:194-303
void check_in_cell(SegmentCtx &c, const CellGrid &cg, float dirLenXYSqr, float *thisStepLength, bool *hitRecorded,
                   unsigned short *hitOnString, unsigned short *hitOnDom)
{
    const Vec4 &pos = c.ph.pos;
    const Vec4 &dir = c.ph.dir;
    int lowCellX = static_cast<int>((pos.x - cg.startX) / cg.widthX);
    int lowCellY = static_cast<int>((pos.y - cg.startY) / cg.widthY);
    int highCellX = static_cast<int>((pos.x + dir.x * (*thisStepLength) - cg.startX) / cg.widthX);
    int highCellY = static_cast<int>((pos.y + dir.y * (*thisStepLength) - cg.startY) / cg.widthY);
    if (highCellX < lowCellX) std::swap(lowCellX, highCellX);
    if (highCellY < lowCellY) std::swap(lowCellY, highCellY);
    lowCellX = std::min(std::max(lowCellX, 0), cg.numX - 1);
    lowCellY = std::min(std::max(lowCellY, 0), cg.numY - 1);
    highCellX = std::min(std::max(highCellX, 0), cg.numX - 1);
    highCellY = std::min(std::max(highCellY, 0), cg.numY - 1);

    std::vector<bool> stringChecked;
    if (!c.sc.stop) stringChecked.assign(c.sc.geo.numStrings, false);

    for (int cell_y = lowCellY; cell_y <= highCellY; ++cell_y) {
        for (int cell_x = lowCellX; cell_x <= highCellX; ++cell_x) {
            const unsigned short stringNum = cg.index[cell_y * cg.numX + cell_x];
            if (stringNum == 0xFFFF) continue;
            if (!c.sc.stop) {
                if (stringChecked[stringNum]) continue;
                stringChecked[stringNum] = true;
            }
            check_on_string(c, stringNum, dirLenXYSqr, thisStepLength, hitRecorded, hitOnString, hitOnDom);
        }
    }
}

// sparse_collision_kernel.c.cl:462-587 (+ :305-460)
bool check_for_collision(SegmentCtx &c, float *thisStepLength)
{
    const Vec4 &dir = c.ph.dir;
    const float dirLenXYSqr = sqr(dir.x) + sqr(dir.y);
    if (dirLenXYSqr <= 0.f) return false;
    bool hitRecorded = false;
    unsigned short hitOnString = 0, hitOnDom = 0;
    for (const CellGrid &cg : c.sc.geo.cells) check_in_cell(c, cg, dirLenXYSqr, thisStepLength, &hitRecorded, &hitOnString, &hitOnDom);
    if (c.sc.stop) {
        if (hitRecorded) {
            save_hit(c.sc, c.sink, c.ph.pos, c.ph.dir, *thisStepLength, c.ph.invGroupVel, c.ph.totalPath, c.ph.numScatters,
                     c.distInAbsLens, c.ph.startPos, c.ph.startDir, c.step, hitOnString, hitOnDom, c.curHistory);
        }
        return hitRecorded;
    }
    return false;
}

struct TrajSink {
    float *buf;
    int maxPoints;
    int n;
    void record(const PhotonState &ph, float absLeft)
    {
        if (!buf || n >= maxPoints) { if (buf) ++n; return; }
        float *r = buf + 8 * n;
        r[0] = ph.pos.x; r[1] = ph.pos.y; r[2] = ph.pos.z; r[3] = ph.pos.w;
        r[4] = ph.dir.x; r[5] = ph.dir.y; r[6] = ph.dir.z; r[7] = absLeft;
        ++n;
    }
};

struct WorkItemStats {
    uint64_t photons = 0, segments = 0, crossings = 0;
};

// ---- table-maker variant (-DTABULATE) --------------------------------------------------------------------
// The binning code Axes::GenerateBinningCode prints (private/clsim/tabulator/Axes.cxx:71-123, Axis.cxx:44-60),
// as data, and the table the host loop of I3CLSimStepToTableConverter sums the entries into (…cxx:484-497).
struct TableAxis {
    float scale, offset;
    int nBins;
    int kind;      // 0 linear, 1 power
    unsigned power;
    float invPower;
    uint64_t stride;
};
struct TableSink {
    int geometry;  // 0 spherical_coordinates.c.cl, 1 cylindrical_coordinates.c.cl
    int ndim;
    bool fullAzimuth;
    TableAxis axes[5];
    float max0, max3, stepLength, minInvGroupVel, tanThetaC;
    std::vector<float> angular;   // getAngularAcceptance literals
    Vec4 refPos, refDir, refPerp; // I3CLSimReferenceParticle
    double *bins;                 // summed in double: the checker's "exact" sum
    double *squared;
    uint64_t entries;
};

inline float dot4(const Vec4 &a, const Vec4 &b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }   // OpenCL dot(float4, float4)
inline float magnitude3(const Vec4 &v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

// getAngularAcceptance as I3CLSimFunctionPolynomial::GetOpenCLFunction nests it (…Polynomial.cxx:139-153)
float table_angular(const TableSink &t, float x)
{
    if (t.angular.empty()) return 0.f;
    float v = t.angular.back();
    for (int i = static_cast<int>(t.angular.size()) - 2; i >= 0; --i) v = t.angular[i] + x * v;
    return v;
}

void table_coords(const TableSink &t, const Vec4 &absPos, Vec4 dirAndWlen, Rng &rng, float c[5])
{
    Vec4 pos{absPos.x - t.refPos.x, absPos.y - t.refPos.y, absPos.z - t.refPos.z, absPos.w - t.refPos.w};
    const float l = dot4(pos, t.refDir);
    Vec4 rho{pos.x - l * t.refDir.x, pos.y - l * t.refDir.y, pos.z - l * t.refDir.z, pos.w - l * t.refDir.w};
    if (t.geometry == 0) {
        const float n_rho = magnitude3(rho);
        c[0] = magnitude3(pos);
        const float azimuth = (n_rho > 0) ? std::acos(dot4(rho, t.refPerp) / n_rho) / (kPI / 180) : 0;
        if (t.fullAzimuth) {
            const float cx = rho.y * t.refPerp.z - rho.z * t.refPerp.y, cy = rho.z * t.refPerp.x - rho.x * t.refPerp.z,
                        cz = rho.x * t.refPerp.y - rho.y * t.refPerp.x;
            const float azisign = (cx * t.refDir.x + cy * t.refDir.y) + cz * t.refDir.z;
            c[1] = (azisign > 0) ? 360.f - azimuth : azimuth;
        } else {
            c[1] = azimuth;
        }
        c[2] = (c[0] > 0) ? (l / c[0]) : 0;
        c[3] = pos.w - c[0] * t.minInvGroupVel;
        if (t.ndim > 4) {
            const float sina = std::sqrt(rand_co(rng));
            scatter_direction_by_angle(std::sqrt(1 - sina * sina), sina, dirAndWlen, rand_co(rng));
            c[4] = (c[0] > 0) ? (dot4(dirAndWlen, pos) / c[0]) : 1;
        }
    } else {
        c[0] = magnitude3(rho);
        c[1] = (c[0] > 0) ? std::acos(dot4(rho, t.refPerp) / c[0]) : 0;
        c[2] = t.refPos.z + l * t.refDir.z;
        c[3] = pos.w - (l + c[0] * t.tanThetaC) * 3.33564095f;
        if (t.ndim > 4) {
            const float sina = std::sqrt(rand_co(rng));
            scatter_direction_by_angle(std::sqrt(1 - sina * sina), sina, dirAndWlen, rand_co(rng));
            const float k = 1.f / t.tanThetaC;
            Vec4 cpos{absPos.x - (t.refPos.x + (l - rho.x * k) * t.refDir.x), absPos.y - (t.refPos.y + (l - rho.y * k) * t.refDir.y),
                      absPos.z - (t.refPos.z + (l - rho.z * k) * t.refDir.z), absPos.w - (t.refPos.w + (l - rho.w * k) * t.refDir.w)};
            const float cdist = magnitude3(cpos);
            c[4] = (cdist > 0) ? (dot4(dirAndWlen, cpos) / cdist) : 1;
        }
    }
}

uint64_t table_index(const TableSink &t, const float c[5])
{
    uint64_t index = 0;
    for (int i = 0; i < t.ndim; ++i) {
        const TableAxis &ax = t.axes[i];
        float v = c[i];
        if (ax.kind == 1) {
            if (ax.power == 0) v = 1.f;
            else if (ax.power == 2) v = std::sqrt(v);
            else if (ax.power == 3) v = std::cbrt(v);
            else if (ax.power != 1) v = std::pow(v, ax.invPower);
        }
        const float f = std::floor(ax.scale * v - ax.offset);   // convert_int_sat_rtn
        long long k = (f != f) ? 0 : ((f >= 2147483648.f) ? 2147483647ll : ((f <= -2147483648.f) ? -2147483648ll : static_cast<long long>(f)));
        k = std::min<long long>(std::max<long long>(k, -1), ax.nBins) + 1;
        index += ax.stride * static_cast<uint64_t>(k);
    }
    return index;
}

// savePath (propagation_kernel.c.cl:226-304) with an entry buffer that never runs out
void save_path(TableSink &t, const oracle_step &step, const PhotonState &ph, float thisStepLength, float *prevStepLength, float depth,
               float thisStepDepth, bool *stop, Rng &rng)
{
    const float impactWeight = (t.ndim > 4) ? step.weight : step.weight * table_angular(t, ph.dir.z);
    float d = *prevStepLength;
    for (; d < thisStepLength; d += t.stepLength) {
        Vec4 pos = ph.pos;
        pos.x = ph.pos.x + d * ph.dir.x;
        pos.y = ph.pos.y + d * ph.dir.y;
        pos.z = ph.pos.z + d * ph.dir.z;
        pos.w = ph.pos.w + d * ph.invGroupVel;
        float c[5];
        table_coords(t, pos, ph.dir, rng, c);
        const bool out = (t.geometry == 0) ? ((c[3] > t.max3) || (c[0] > t.max0)) : (c[3] > t.max3);
        if (out) {
            *stop = true;
            break;
        }
        const uint64_t index = table_index(t, c);
        const float weight = impactWeight * std::exp(-(depth + (d / thisStepLength) * thisStepDepth));
        t.bins[index] += weight;
        if (t.squared) t.squared[index] += static_cast<double>(weight) * weight;
        ++t.entries;
    }
    *prevStepLength = d - thisStepLength;
}

// One work-item of propKernel (propagation_kernel.c.cl:406-913).  max_photons limits the
// photons taken from the step (single-photon replay uses 1).
void run_work_item(const oracle_scene &sc, const oracle_step &step, Rng &rng, HitSink &sink, WorkItemStats &st,
                   uint32_t maxPhotons, TrajSink *traj, const uint64_t *xAfterCreation = nullptr, uint32_t aAfterCreation = 0,
                   TableSink *table = nullptr)
{
    float depthPropagated = 0.f, prevStepRemainder = 0.f;   // TABULATE
    const Medium &m = sc.med;
    Vec4 stepDir;
    {
        const float rho = std::sin(step.dir_and_length_and_beta[0]);
        stepDir.x = rho * std::cos(step.dir_and_length_and_beta[1]);
        stepDir.y = rho * std::sin(step.dir_and_length_and_beta[1]);
        stepDir.z = std::cos(step.dir_and_length_and_beta[0]);
        stepDir.w = 0.f;
    }
    uint32_t photonsLeft = std::min(step.num_photons, maxPhotons);
    float abs_lens_left = 0.f;
    PhotonState ph;
    std::memset(&ph, 0, sizeof(ph));
    const bool tiltConstant = (m.tiltND == 0); // getTiltZShift_IS_CONSTANT (I3CLSimScalarFieldConstant.cxx:67)
    int currentPhotonLayer = 0;
    std::vector<float> curHistory(sc.history > 0 ? 4 * sc.history : 4, 0.f);

    while (photonsLeft > 0) {
        if (abs_lens_left < kEpsilon) {
            create_photon_from_track(sc, step, stepDir, rng, ph.pos, ph.dir);
            ph.startPos = ph.pos;
            ph.startDir = ph.dir;
            ph.numScatters = 0;
            ph.totalPath = 0.f;
            if (table) prevStepRemainder = table->stepLength * rand_oc(rng);   // propagation_kernel.c.cl:566-569
            if (tiltConstant) currentPhotonLayer = std::min(std::max(find_layer(m, ph.pos.z), 0), m.L - 1);
            ph.invGroupVel = 1.f / group_velocity(m, ph.dir.w);
            if (sc.fixedAbs) ph.absLensInitial = sc.fixedAbsLens;
            else ph.absLensInitial = -std::log(rand_oc(rng));
            abs_lens_left = ph.absLensInitial;
            if (table) depthPropagated = 0.f;   // :590-592
            // replay of a photon that was created from one stream and propagated from another
            // (the B200 fast kernel keeps a creation stream and a propagation stream per lane)
            if (xAfterCreation) {
                rng.x = *xAfterCreation;
                rng.a = aAfterCreation;
            }
            ++st.photons;
            if (traj) traj->record(ph, abs_lens_left);
        }

        float distancePropagated;
        {
            float effective_z;
            if (tiltConstant) {
                effective_z = ph.pos.z - 0.f; // (photonPosAndTime.z-getTiltZShift_IS_CONSTANT), constant 0.f
            } else {
                effective_z = ph.pos.z - tilt_z_shift(m, ph.pos.x, ph.pos.y, ph.pos.z);
                currentPhotonLayer = std::min(std::max(find_layer(m, effective_z), 0), m.L - 1);
            }
            const float photon_dz = ph.dir.z;
            const float abs_len_correction_factor = abs_len_corr_factor(m, ph.dir.x, ph.dir.y, ph.dir.z);
            abs_lens_left *= abs_len_correction_factor;

            float mediumBoundary = (photon_dz < 0.f) ? (layer_boundary(m, currentPhotonLayer)) : (layer_boundary(m, currentPhotonLayer) + m.h);
            const float sca_step_left = -std::log(rand_oc(rng));
            float currentScaLen = scattering_length(m, currentPhotonLayer, ph.dir.w);
            float currentAbsLen = absorption_length(m, currentPhotonLayer, ph.dir.w);
            float ais = (photon_dz * sca_step_left - ((mediumBoundary - effective_z) / currentScaLen)) * (1.f / m.h);
            float aia = (photon_dz * abs_lens_left - ((mediumBoundary - effective_z) / currentAbsLen)) * (1.f / m.h);

            int j = currentPhotonLayer;
            if (photon_dz < 0) {
                while ((j > 0) && (ais < 0.f) && (aia < 0.f)) {
                    --j;
                    mediumBoundary -= m.h;
                    currentScaLen = scattering_length(m, j, ph.dir.w);
                    currentAbsLen = absorption_length(m, j, ph.dir.w);
                    ais += 1.f / currentScaLen;
                    aia += 1.f / currentAbsLen;
                    ++st.crossings;
                }
            } else {
                while ((j < m.L - 1) && (ais > 0.f) && (aia > 0.f)) {
                    ++j;
                    mediumBoundary += m.h;
                    currentScaLen = scattering_length(m, j, ph.dir.w);
                    currentAbsLen = absorption_length(m, j, ph.dir.w);
                    ais -= 1.f / currentScaLen;
                    aia -= 1.f / currentAbsLen;
                    ++st.crossings;
                }
            }

            float distanceToAbsorption;
            if ((currentPhotonLayer == j) || (std::fabs(photon_dz) < kEpsilon)) {
                distancePropagated = sca_step_left * currentScaLen;
                distanceToAbsorption = abs_lens_left * currentAbsLen;
            } else {
                const float recip_photon_dz = 1.f / photon_dz;
                distancePropagated = (ais * m.h * currentScaLen + mediumBoundary - effective_z) * recip_photon_dz;
                distanceToAbsorption = (aia * m.h * currentAbsLen + mediumBoundary - effective_z) * recip_photon_dz;
            }
            if (tiltConstant) currentPhotonLayer = j;

            if (distanceToAbsorption < distancePropagated) {
                distancePropagated = distanceToAbsorption;
                abs_lens_left = 0.f;
            } else {
                abs_lens_left = (distanceToAbsorption - distancePropagated) / currentAbsLen;
            }
            abs_lens_left = abs_lens_left / abs_len_correction_factor;
        }
        ++st.segments;

        if (!sc.saveAll) {
            SegmentCtx ctx{sc, sink, step, ph, ph.absLensInitial - abs_lens_left, curHistory.data()};
            const bool collided = check_for_collision(ctx, &distancePropagated);
            if (sc.stop && collided) abs_lens_left = 0.f;
        }

        if (table) {
            // :755-785
            bool stop = false;
            save_path(*table, step, ph, distancePropagated, &prevStepRemainder, depthPropagated,
                      ph.absLensInitial - abs_lens_left - depthPropagated, &stop, rng);
            if (stop) abs_lens_left = 0.f;
            depthPropagated = ph.absLensInitial - abs_lens_left;
        }

        ph.pos.x += ph.dir.x * distancePropagated;
        ph.pos.y += ph.dir.y * distancePropagated;
        ph.pos.z += ph.dir.z * distancePropagated;
        ph.pos.w += ph.invGroupVel * distancePropagated;
        ph.totalPath += distancePropagated;

        if (abs_lens_left < kEpsilon) {
            --photonsLeft;
            if (sc.saveAll && !table) {   // #if defined(SAVE_ALL_PHOTONS) && !defined(TABULATE)
                if (rand_co(rng) < sc.prescale) {
                    save_hit(sc, sink, ph.pos, ph.dir, 0.f, ph.invGroupVel, ph.totalPath, ph.numScatters, ph.absLensInitial,
                             ph.startPos, ph.startDir, step, 0, 0, curHistory.data());
                }
            }
            if (traj) traj->record(ph, abs_lens_left);
        } else {
            if (sc.history > 0) {
                float *hrow = &curHistory[4 * (ph.numScatters % sc.history)];
                hrow[0] = ph.pos.x; hrow[1] = ph.pos.y; hrow[2] = ph.pos.z;
                hrow[3] = ph.absLensInitial - abs_lens_left;
            }
            if (m.aniso) transform_dir(m.pre, m.preRenorm, ph.dir.x, ph.dir.y, ph.dir.z);
            const float cosScatAngle = scattering_cos_angle(m, rng);
            const float sinScatAngle = std::sqrt(1.f - sqr(cosScatAngle));
            scatter_direction_by_angle(cosScatAngle, sinScatAngle, ph.dir, rand_co(rng));
            if (m.aniso) transform_dir(m.post, m.postRenorm, ph.dir.x, ph.dir.y, ph.dir.z);
            ++ph.numScatters;
            if (traj) traj->record(ph, abs_lens_left);
        }
    }
}

std::string json_array_f(const std::vector<float> &v)
{
    std::ostringstream o;
    o << "[";
    for (size_t i = 0; i < v.size(); ++i) {
        char b[40];
        std::snprintf(b, sizeof(b), "%.9g", v[i]);
        o << (i ? "," : "") << b;
    }
    o << "]";
    return o.str();
}
template <class T> std::string json_array_i(const std::vector<T> &v)
{
    std::ostringstream o;
    o << "[";
    for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << static_cast<long long>(v[i]);
    o << "]";
    return o.str();
}

std::string describe(const oracle_scene &sc)
{
    const GeoTables &g = sc.geo;
    std::ostringstream o;
    o << "{";
    o << "\"num_strings\":" << g.numStrings;
    char b[64];
    std::snprintf(b, sizeof(b), "%.9g", g.omRadius);
    o << ",\"om_radius\":" << b;
    std::snprintf(b, sizeof(b), "%.9g", g.stringMaxRadius);
    o << ",\"string_max_radius\":" << b;
    o << ",\"string_pos_x\":" << json_array_f(g.stringPosX);
    o << ",\"string_pos_y\":" << json_array_f(g.stringPosY);
    o << ",\"string_min_z\":" << json_array_f(g.stringMinZ);
    o << ",\"string_max_z\":" << json_array_f(g.stringMaxZ);
    o << ",\"string_in_set\":" << json_array_i(g.stringInSet);
    o << ",\"num_sets\":" << g.numSets << ",\"max_layers\":" << g.maxLayers;
    o << ",\"layer_num\":" << json_array_i(g.layerNum);
    o << ",\"layer_start_z\":" << json_array_f(g.layerStartZ);
    o << ",\"layer_height\":" << json_array_f(g.layerHeight);
    o << ",\"layer_to_om\":" << json_array_i(g.layerToOM);
    o << ",\"cells\":[";
    for (size_t i = 0; i < g.cells.size(); ++i) {
        const CellGrid &c = g.cells[i];
        o << (i ? "," : "") << "{\"num_x\":" << c.numX << ",\"num_y\":" << c.numY;
        o << ",\"start_width\":" << json_array_f({c.startX, c.startY, c.widthX, c.widthY});
        o << ",\"index\":" << json_array_i(c.index) << "}";
    }
    o << "]";
    o << ",\"max_dom_index\":" << g.maxDomIndex;
    o << ",\"tmpl_mul\":" << json_array_f({g.mulX, g.mulY});
    o << ",\"tmpl_x\":" << json_array_i(g.tmplX) << ",\"tmpl_y\":" << json_array_i(g.tmplY);
    o << ",\"tmpl_z\":" << json_array_f(g.tmplZ);
    o << ",\"string_tmpl_start\":" << json_array_i(g.stringTmplStart);
    o << ",\"string_mean_x\":" << json_array_f(g.stringMeanX) << ",\"string_mean_y\":" << json_array_f(g.stringMeanY);
    o << ",\"string_index_to_id\":" << json_array_i(g.stringIndexToID);
    o << ",\"dom_index_to_id\":[";
    for (size_t i = 0; i < g.domIndexToID.size(); ++i) o << (i ? "," : "") << json_array_i(g.domIndexToID[i]);
    o << "]";
    // medium + generator tables
    o << ",\"medium\":{\"b400\":" << json_array_f(sc.med.b400) << ",\"a_dust400\":" << json_array_f(sc.med.aDust)
      << ",\"delta_tau\":" << json_array_f(sc.med.dTau) << "}";
    o << ",\"wlen_generators\":[";
    for (size_t i = 0; i < sc.gens.size(); ++i) {
        o << (i ? "," : "") << "{\"kind\":" << sc.gens[i].kind << ",\"beta\":" << json_array_f(sc.gens[i].beta)
          << ",\"acu\":" << json_array_f(sc.gens[i].acu) << ",\"xs\":" << json_array_f(sc.gens[i].xs) << "}";
    }
    o << "]";
    o << "}";
    return o.str();
}

inline uint64_t splitmix64(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

} // namespace

extern "C" {

size_t oracle_sizeof_config(void) { return sizeof(oracle_config); }
const char *oracle_last_error(void) { return g_last_error.c_str(); }

oracle_scene *oracle_scene_create(const oracle_config *cfg)
{
    try {
        if (!cfg) throw std::runtime_error("config is NULL");
        if (cfg->struct_size != static_cast<int32_t>(sizeof(oracle_config))) throw std::runtime_error("config struct_size mismatch");
        // …OpenCL.cxx:507-508
        if (cfg->save_all_photons && cfg->stop_detected_photons)
            throw std::runtime_error("both the saveAllPhotons and stopDetectedPhotons options are set at the same time.");
        if (cfg->num_wlen_generators < 1 || !cfg->wlen_generators) throw std::runtime_error("WlenGenerators not set!");
        std::unique_ptr<oracle_scene> sc(new oracle_scene());
        build_medium(cfg->medium, sc->med);
        for (int i = 0; i < cfg->num_wlen_generators; ++i) {
            const oracle_wlen_generator &g = cfg->wlen_generators[i];
            WlenGen w;
            w.kind = g.kind;
            if (g.kind == 0 || g.kind == 1) {
                init_interp_tables(g, w);
            } else if (g.kind == 2) {
                const double minVal = 1. / g.to_wlen;
                const double range = (1. / g.from_wlen) - minVal;
                w.min_val = lit(minVal);
                w.range = lit(range);
            } else if (g.kind == 3) {
                w.value = lit(g.value);
            } else {
                throw std::runtime_error("unknown wavelength generator kind");
            }
            sc->gens.push_back(w);
        }
        sc->bias.kind = cfg->wlen_bias.kind;
        if (cfg->wlen_bias.kind == 1) {
            sc->bias.n = cfg->wlen_bias.n;
            if (sc->bias.n < 2) throw std::runtime_error("bias table needs at least 2 entries");
            sc->bias.x0 = lit(cfg->wlen_bias.x0);
            sc->bias.dx = lit(cfg->wlen_bias.dx);
            for (int i = 0; i < sc->bias.n; ++i) sc->bias.v.push_back(lit(cfg->wlen_bias.v[i]));
        } else {
            sc->bias.value = lit(cfg->wlen_bias.value);
        }
        sc->stop = cfg->stop_detected_photons != 0;
        sc->saveAll = cfg->save_all_photons != 0;
        sc->prescale = lit(cfg->save_all_photons_prescale);
        sc->fixedAbs = !std::isnan(cfg->fixed_number_of_absorption_lengths);
        if (sc->fixedAbs) sc->fixedAbsLens = lit(cfg->fixed_number_of_absorption_lengths);
        sc->pancake = (cfg->pancake_factor != 1.); // …OpenCL.cxx:434-440
        sc->pancakeFactor = lit(cfg->pancake_factor);
        sc->history = cfg->photon_history_entries;
        if (!sc->saveAll) {
            build_geometry(cfg->geometry, sc->geo);
            sc->haveGeo = true;
        }
        return sc.release();
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return nullptr;
    }
}

void oracle_scene_destroy(oracle_scene *scene) { delete scene; }

uint64_t oracle_propagate(const oracle_scene *scene, const oracle_step *steps, size_t n, uint64_t *rng_x, const uint32_t *rng_a,
                          oracle_photon *out, size_t cap, float *history, int num_threads, uint64_t stats[4])
{
    const oracle_scene &sc = *scene;
    uint64_t totPhot = 0, totSeg = 0, totCross = 0, totDraws = 0;
    if (num_threads <= 1) {
        HitSink sink{out, cap, history, 0, {}, {}, false};
        for (size_t i = 0; i < n; ++i) {
            Rng rng{rng_x[i], rng_a[i], 0};
            WorkItemStats st;
            run_work_item(sc, steps[i], rng, sink, st, 0xffffffffu, nullptr);
            rng_x[i] = rng.x;
            totPhot += st.photons; totSeg += st.segments; totCross += st.crossings; totDraws += rng.draws;
        }
        if (stats) { stats[0] = totPhot; stats[1] = totSeg; stats[2] = totCross; stats[3] = totDraws; }
        return sink.count;
    }
    // chunked so that the output order stays (work-item, emission)
    const size_t chunk = 256;
    const size_t numChunks = (n + chunk - 1) / chunk;
    std::vector<std::vector<oracle_photon>> chunkHits(numChunks);
    std::vector<std::vector<float>> chunkHist(numChunks);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads) reduction(+ : totPhot, totSeg, totCross, totDraws)
#endif
    for (long long c = 0; c < static_cast<long long>(numChunks); ++c) {
        HitSink sink{nullptr, 0, nullptr, 0, {}, {}, true};
        const size_t lo = c * chunk, hi = std::min(n, lo + chunk);
        for (size_t i = lo; i < hi; ++i) {
            Rng rng{rng_x[i], rng_a[i], 0};
            WorkItemStats st;
            run_work_item(sc, steps[i], rng, sink, st, 0xffffffffu, nullptr);
            rng_x[i] = rng.x;
            totPhot += st.photons; totSeg += st.segments; totCross += st.crossings; totDraws += rng.draws;
        }
        chunkHits[c].swap(sink.local);
        chunkHist[c].swap(sink.localHistory);
    }
    uint64_t count = 0;
    for (size_t c = 0; c < numChunks; ++c) {
        for (size_t k = 0; k < chunkHits[c].size(); ++k) {
            if (count < cap && out) {
                out[count] = chunkHits[c][k];
                if (sc.history > 0 && history) std::memcpy(history + count * 4 * sc.history, &chunkHist[c][k * 4 * sc.history], sizeof(float) * 4 * sc.history);
            }
            ++count;
        }
    }
    if (stats) { stats[0] = totPhot; stats[1] = totSeg; stats[2] = totCross; stats[3] = totDraws; }
    return count;
}

uint64_t oracle_tabulate(const oracle_scene *scene, const oracle_table_config *tc, const oracle_step *steps, size_t n, uint64_t *rng_x,
                         const uint32_t *rng_a, const double reference[7], double *bins, double *squared)
{
    if (!scene->saveAll || !scene->fixedAbs) {
        g_last_error = "the table-maker variant needs save_all_photons and a fixed number of absorption lengths";
        return 0;
    }
    TableSink t;
    t.geometry = tc->geometry;
    t.ndim = tc->num_axes;
    t.fullAzimuth = (tc->geometry == 0 && tc->axis_max[1] > 180.);
    uint64_t stride = 1;
    for (int i = tc->num_axes - 1; i >= 0; --i) {
        TableAxis &ax = t.axes[i];
        ax.kind = tc->axis_kind[i];
        ax.power = tc->axis_power[i];
        ax.nBins = static_cast<int>(tc->axis_bins[i]);
        auto inverse = [&](double v) { return ax.kind == 1 ? std::pow(v, 1. / ax.power) : v; };
        // Axis::GetIndexCode (Axis.cxx:44-60)
        const double scale = ax.nBins / (inverse(tc->axis_max[i]) - inverse(tc->axis_min[i]));
        ax.scale = lit(scale);
        ax.offset = lit(scale * inverse(tc->axis_min[i]));
        ax.invPower = (ax.kind == 1 && ax.power > 0) ? lit(1. / ax.power) : 0.f;
        ax.stride = stride;
        stride *= static_cast<uint64_t>(ax.nBins) + 2;
    }
    t.max0 = lit(tc->axis_max[0]);
    t.max3 = lit(tc->axis_max[3]);
    t.stepLength = lit(tc->step_length);
    t.minInvGroupVel = lit(tc->n_group / 0.299792458);
    t.tanThetaC = lit(std::sqrt(tc->n_phase * tc->n_phase - 1.));
    for (int i = 0; i < tc->num_angular_coefficients; ++i) t.angular.push_back(lit(tc->angular_coefficients[i]));
    // I3CLSimReferenceParticle (…StepToTableConverter.cxx:65-93)
    t.refPos = Vec4{static_cast<float>(reference[0]), static_cast<float>(reference[1]), static_cast<float>(reference[2]), static_cast<float>(reference[3])};
    t.refDir = Vec4{static_cast<float>(reference[4]), static_cast<float>(reference[5]), static_cast<float>(reference[6]), 0.f};
    {
        const double perpz = std::hypot(reference[4], reference[5]);
        double px = 1., py = 0., pz = 0.;
        if (perpz > 0.) {
            px = -reference[4] * reference[6] / perpz;
            py = -reference[5] * reference[6] / perpz;
            pz = perpz;
            const double norm = std::sqrt(px * px + py * py + pz * pz);
            px /= norm; py /= norm; pz /= norm;
        }
        t.refPerp = Vec4{static_cast<float>(px), static_cast<float>(py), static_cast<float>(pz), 0.f};
    }
    t.bins = bins;
    t.squared = squared;
    t.entries = 0;
    HitSink sink{nullptr, 0, nullptr, 0, {}, {}, false};
    for (size_t i = 0; i < n; ++i) {
        Rng rng{rng_x[i], rng_a[i], 0};
        WorkItemStats st;
        run_work_item(*scene, steps[i], rng, sink, st, 0xffffffffu, nullptr, nullptr, 0, &t);
        rng_x[i] = rng.x;
    }
    return t.entries;
}

int oracle_propagate_single_photon(const oracle_scene *scene, const oracle_step *step, uint64_t *x, uint32_t a, oracle_photon *out,
                                   float *traj, int max_points, int *num_points)
{
    Rng rng{*x, a, 0};
    oracle_photon tmp[4];
    HitSink sink{tmp, 4, nullptr, 0, {}, {}, false};
    if (scene->history > 0) { sink.useLocal = true; }
    WorkItemStats st;
    TrajSink ts{traj, max_points, 0};
    run_work_item(*scene, *step, rng, sink, st, 1, &ts);
    *x = rng.x;
    if (num_points) *num_points = ts.n;
    if (sink.count > 0 && out) *out = sink.useLocal ? sink.local[0] : tmp[0];
    return sink.count > 0 ? 1 : 0;
}

int oracle_propagate_single_photon_split(const oracle_scene *scene, const oracle_step *step, uint64_t x_create, uint32_t a_create,
                                         uint64_t x_propagate, uint32_t a_propagate, oracle_photon *out, float *traj, int max_points,
                                         int *num_points)
{
    Rng rng{x_create, a_create, 0};
    oracle_photon tmp[4];
    HitSink sink{tmp, 4, nullptr, 0, {}, {}, false};
    if (scene->history > 0) { sink.useLocal = true; }
    WorkItemStats st;
    TrajSink ts{traj, max_points, 0};
    run_work_item(*scene, *step, rng, sink, st, 1, &ts, &x_propagate, a_propagate);
    if (num_points) *num_points = ts.n;
    if (sink.count > 0 && out) *out = sink.useLocal ? sink.local[0] : tmp[0];
    return sink.count > 0 ? 1 : 0;
}

void oracle_rng_uniform_co(uint64_t *x, uint32_t a, float *out, size_t n)
{
    Rng r{*x, a, 0};
    for (size_t i = 0; i < n; ++i) out[i] = rand_co(r);
    *x = r.x;
}

int oracle_safeprimes(uint64_t first, uint64_t n, uint32_t *a_out, uint64_t *n2_out, uint64_t *n1_out)
{
    // make_safeprimes/main.cxx:59-104 descends one candidate at a time; the test of a candidate does not depend on the
    // others, so blocks of candidates are tested in parallel and their survivors appended in descending order: the
    // same rows, found on all host cores.
    uint64_t a = 4294967118ull;
    uint64_t row = 0, written = 0;
    const uint64_t block = 1u << 18;
    std::vector<uint8_t> keep(block);
    while (written < n) {
        if (a == 0) { g_last_error = "ran out of multiplier candidates"; return -1; }
        const uint64_t count = std::min<uint64_t>(block, a);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
        for (long long k = 0; k < static_cast<long long>(count); ++k) {
            const uint64_t cand = a - static_cast<uint64_t>(k);
            const uint64_t n2 = (cand << 32) - 1;
            keep[k] = (is_prime_u64(n2) && is_prime_u64((n2 - 1) >> 1)) ? 1 : 0;
        }
        for (uint64_t k = 0; k < count && written < n; ++k) {
            if (!keep[k]) continue;
            if (row >= first) {
                const uint64_t cand = a - k, n2 = (cand << 32) - 1;
                a_out[written] = static_cast<uint32_t>(cand);
                if (n2_out) n2_out[written] = n2;
                if (n1_out) n1_out[written] = (n2 - 1) >> 1;
                ++written;
            }
            ++row;
        }
        a -= count;
    }
    return 0;
}

void oracle_rng_seed_states(uint64_t seed, const uint32_t *a, uint64_t *x, size_t n)
{
    uint64_t s = seed;
    for (size_t i = 0; i < n; ++i) {
        // mwcrng_init.h:107-113 with splitmix64 standing in for I3RandomService::Integer
        x[i] = 0;
        while ((x[i] == 0) | ((static_cast<uint32_t>(x[i] >> 32)) >= (a[i] - 1)) | ((static_cast<uint32_t>(x[i])) >= 0xfffffffful)) {
            const uint64_t r = splitmix64(s);
            x[i] = static_cast<uint32_t>(r >> 32);
            x[i] = x[i] << 32;
            x[i] += static_cast<uint32_t>(r);
        }
    }
}

int oracle_describe_tables(const oracle_scene *scene, char *buf, size_t cap, size_t *needed)
{
    const std::string s = describe(*scene);
    if (needed) *needed = s.size() + 1;
    if (buf && cap > s.size()) {
        std::memcpy(buf, s.c_str(), s.size() + 1);
        return 0;
    }
    return buf ? -1 : 0;
}

void oracle_eval_wlen_function(const oracle_scene *scene, int which, const float *in, float *out, size_t n)
{
    const Medium &m = scene->med;
    for (size_t i = 0; i < n; ++i) {
        const int layer = static_cast<int>(in[2 * i]);
        const float w = in[2 * i + 1];
        switch (which) {
        case 0: out[i] = phase_ref_index(m, w); break;
        case 1: out[i] = group_velocity(m, w); break;
        case 2: out[i] = scattering_length(m, layer, w); break;
        case 3: out[i] = absorption_length(m, layer, w); break;
        default: out[i] = wavelength_bias(scene->bias, w); break;
        }
    }
}

void oracle_eval_scalar_field(const oracle_scene *scene, int which, const float *xyz, float *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        if (which == 0) out[i] = tilt_z_shift(scene->med, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        else out[i] = abs_len_corr_factor(scene->med, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
}

void oracle_eval_vector_transform(const oracle_scene *scene, int which, const float *xyz, float *out, size_t n)
{
    const Medium &m = scene->med;
    for (size_t i = 0; i < n; ++i) {
        float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (m.aniso) {
            if (which == 0) transform_dir(m.pre, m.preRenorm, x, y, z);
            else transform_dir(m.post, m.postRenorm, x, y, z);
        }
        out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
}

void oracle_sample(const oracle_scene *scene, int which, uint64_t *x, uint32_t a, float *out, size_t n)
{
    Rng r{*x, a, 0};
    for (size_t i = 0; i < n; ++i) {
        if (which == 0) out[i] = scattering_cos_angle(scene->med, r);
        else out[i] = sample_wlen(scene->gens.at(which - 1), r);
    }
    *x = r.x;
}

void oracle_scatter_direction(const float *in6, float *out3, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        Vec4 d{in6[6 * i + 2], in6[6 * i + 3], in6[6 * i + 4], 0.f};
        scatter_direction_by_angle(in6[6 * i], in6[6 * i + 1], d, in6[6 * i + 5]);
        out3[3 * i] = d.x; out3[3 * i + 1] = d.y; out3[3 * i + 2] = d.z;
    }
}

} // extern "C"
