// tables.cpp -- see tables.h.  Host only (no CUDA), so table parity is testable without a GPU.
#include "tables.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace clsimcu {

float float_literal(double value)
{
    // ToFloatString prints with std::scientific and precision digits10+4 = 10, the OpenCL
    // front end then reads the decimal text as a float: two roundings, reproduced here.
    char text[48];
    std::snprintf(text, sizeof text, "%.10e", value);
    return std::strtof(text, nullptr);
}

namespace {

const double kNaN = std::numeric_limits<double>::quiet_NaN();

// ------------------------------------------------------------------ wavelength generators
// Cumulative/normalised tables of I3CLSimRandomValueInterpolatedDistribution
// (private/clsim/random_value/I3CLSimRandomValueInterpolatedDistribution.cxx:137-230).
WlenGeneratorTable make_generator(const clsimcu_wlen_generator &g)
{
    WlenGeneratorTable t;
    t.kind = g.kind;
    switch (g.kind) {
    case CLSIMCU_WLEN_INTERP_EQUAL:
    case CLSIMCU_WLEN_INTERP_UNEQUAL: {
        if (g.n <= 1 || !g.y) throw std::runtime_error("At least two entries have to be specified for an interpolated distribution.");
        const bool unequal = (g.kind == CLSIMCU_WLEN_INTERP_UNEQUAL);
        if (unequal && !g.x) throw std::runtime_error("The \"x\" and \"y\" vectors must have the same size!");
        if (!unequal && !(g.dx > 0.)) throw std::runtime_error("\"xSpacing\" must not be <= 0!");
        std::vector<double> integral(g.n, 0.);
        for (int j = 1; j < g.n; ++j) {
            const double width = unequal ? (g.x[j] - g.x[j - 1]) : (g.dx);
            integral[j] = integral[j - 1] + width * (g.y[j] + g.y[j - 1]) / 2.;
        }
        const double total = integral[g.n - 1];
        t.n = g.n;
        t.density.resize(g.n);
        t.cumulative.resize(g.n);
        for (int j = 0; j < g.n; ++j) {
            t.density[j] = float_literal(g.y[j] / total);
            t.cumulative[j] = float_literal(integral[j] / total);
        }
        if (unequal) {
            t.xs.resize(g.n);
            for (int j = 0; j < g.n; ++j) t.xs[j] = float_literal(g.x[j]);
        } else {
            t.x0 = float_literal(g.x0);
            t.dx = float_literal(g.dx);
        }
        break;
    }
    case CLSIMCU_WLEN_NO_DISPERSION: {
        // I3CLSimRandomValueWlenCherenkovNoDispersion.cxx:83-97
        if (g.from_wlen > g.to_wlen) throw std::runtime_error("The \"fromWlen\" argument must not be greater than \"toWlen\".");
        const double lowest = 1. / g.to_wlen;
        t.min_val = float_literal(lowest);
        t.range = float_literal((1. / g.from_wlen) - lowest);
        break;
    }
    case CLSIMCU_WLEN_CONSTANT:
        t.value = float_literal(g.value);
        break;
    default:
        throw std::runtime_error("unsupported wavelength generator kind " + std::to_string(g.kind));
    }
    return t;
}

// ------------------------------------------------------------------ medium
void make_medium(const clsimcu_medium &m, MediumTables &t)
{
    if (m.num_layers < 1) throw std::runtime_error("MediumProperties not set!");
    if (!m.a_dust400 || !m.delta_tau || !m.b400) throw std::runtime_error("medium layer tables are NULL");
    if (m.scat_kind < 0 || m.scat_kind > 2) throw std::runtime_error("unsupported scattering angle distribution");
    t.num_layers = m.num_layers;
    t.z0 = float_literal(m.layers_zstart);
    t.h = float_literal(m.layers_height);
    t.kappa = float_literal(m.kappa);
    t.A = float_literal(m.A);
    t.B = float_literal(m.B);
    t.D = float_literal(m.D);
    t.E = float_literal(m.E);
    t.alpha = float_literal(m.alpha);
    t.inv_ref_wlen = float_literal(1. / (400. * 1e-9)); // …_Optimizers.cxx:229
    t.a_dust400.resize(m.num_layers);
    t.delta_tau.resize(m.num_layers);
    t.b400.resize(m.num_layers);
    for (int i = 0; i < m.num_layers; ++i) {
        t.a_dust400[i] = float_literal(m.a_dust400[i]);
        t.delta_tau[i] = float_literal(m.delta_tau[i]);
        t.b400[i] = float_literal(m.b400[i]);
    }
    for (int i = 0; i < 5; ++i) {
        t.n_phase[i] = float_literal(m.n_phase[i]);
        t.n_group[i] = float_literal(m.n_group[i]);
    }
    t.c_light = float_literal(0.299792458); // I3Constants::c [m/ns], …MediumPropertiesSource.cxx:267
    t.scat_kind = m.scat_kind;
    t.f_sl = float_literal(m.f_sl);                 // I3CLSimRandomValueMixed.cxx:135-141
    t.one_minus_f_sl = float_literal(1. - m.f_sl);
    t.g = float_literal(m.mean_cos);                // I3CLSimRandomValueHenyeyGreenstein.cxx:82-83
    t.g2 = float_literal(m.mean_cos * m.mean_cos);
    t.sl_beta = float_literal((1. - m.mean_cos) / (1. + m.mean_cos)); // …SimplifiedLiu.cxx:78

    t.tilt_nd = m.tilt_num_dist;
    t.tilt_nz = m.tilt_num_z;
    if (t.tilt_nd != 0) {
        // I3CLSimScalarFieldIceTiltZShift.cxx:57-60
        if (t.tilt_nd < 2) throw std::runtime_error("distancesFromOriginAlongTilt (dimension 1) needs at least 2 entries.");
        if (t.tilt_nz < 2) throw std::runtime_error("zCoordinates (dimension 2) needs at least 2 entries.");
        if (!m.tilt_dist || !m.tilt_corr) throw std::runtime_error("tilt tables are NULL");
        t.tilt_dist.resize(t.tilt_nd);
        for (int i = 0; i < t.tilt_nd; ++i) {
            if (i > 0 && !(m.tilt_dist[i] - m.tilt_dist[i - 1] > 0.))
                throw std::runtime_error("distancesFromOriginAlongTilt (dimension 1) is not in ascending order.");
            t.tilt_dist[i] = float_literal(m.tilt_dist[i]);
        }
        t.tilt_corr.resize(static_cast<size_t>(t.tilt_nd) * t.tilt_nz);
        for (size_t i = 0; i < t.tilt_corr.size(); ++i) t.tilt_corr[i] = float_literal(m.tilt_corr[i]);
        t.tilt_z0 = float_literal(m.tilt_z0);
        t.tilt_dz = float_literal(m.tilt_dz);
        t.tilt_lnx = float_literal(std::cos(m.tilt_azimuth));
        t.tilt_lny = float_literal(std::sin(m.tilt_azimuth));
    }

    t.anisotropy = (m.has_anisotropy != 0);
    if (t.anisotropy) {
        // constants of I3CLSimScalarFieldAnisotropyAbsLenScaling::GetOpenCLFunction (:92-134)
        const double ca = std::cos(m.aniso_azimuth), sa = std::sin(m.aniso_azimuth);
        const double k1 = std::exp(m.aniso_along), k2 = std::exp(m.aniso_perp);
        const double kz = 1. / (k1 * k2);
        const double lam[3] = {k1 * k1, k2 * k2, kz * kz};
        for (int i = 0; i < 3; ++i) {
            t.l[i] = float_literal(lam[i]);
            t.rl[i] = float_literal(1. / lam[i]);
        }
        t.B2 = float_literal(1. / lam[0] + 1. / lam[1] + 1. / lam[2]);
        t.azx = float_literal(ca);
        t.azy = float_literal(sa);
        t.neg_azy = float_literal(-sa);
        for (int i = 0; i < 9; ++i) {
            t.pre[i] = float_literal(m.pre_matrix[i]);
            t.post[i] = float_literal(m.post_matrix[i]);
        }
        t.pre_renorm = m.pre_renormalize != 0;
        t.post_renorm = m.post_renormalize != 0;
    }
}

// ------------------------------------------------------------------ geometry
struct Dom {
    uint32_t id;
    double x, y, z;
};
struct DetString {
    int id = 0;
    int subdet = 0;
    std::vector<Dom> doms;
    double cx = 0., cy = 0.;        // mean x/y
    double zlo = kNaN, zhi = kNaN;  // DOM-centre z range
    double dz = 0.;                 // typical DOM spacing
    double reach = kNaN;            // largest lateral DOM offset + OM radius
};

inline void grow_min(double &m, double v) { if (std::isnan(m) || v < m) m = v; }
inline void grow_max(double &m, double v) { if (std::isnan(m) || v > m) m = v; }

// interval [lo,hi] touches the slab [a,b] (the reference's three-clause test,
// …GeometrySource.cxx:212-229, 305-313)
inline bool touches(double lo, double hi, double a, double b)
{
    return ((lo <= a) && (hi >= a)) || ((lo <= b) && (hi >= b)) || ((lo >= a) && (hi <= b));
}

// …GeometrySource.cxx:135-271: n x n grid over the bounding box of one subdetector, fails if
// any cell is touched by two strings.
bool try_grid(const std::vector<DetString> &strings, int subdet, unsigned n, CellGridTable &grid, double &sx, double &sy,
              double &wx, double &wy)
{
    double xlo = kNaN, xhi = kNaN, ylo = kNaN, yhi = kNaN;
    size_t members = 0;
    for (const DetString &s : strings) {
        if (s.subdet != subdet) continue;
        ++members;
        grow_min(xlo, s.cx - s.reach);
        grow_min(ylo, s.cy - s.reach);
        grow_max(xhi, s.cx + s.reach);
        grow_max(yhi, s.cy + s.reach);
    }
    if (members == 0) throw std::runtime_error("no strings found");
    sx = xlo;
    sy = ylo;
    wx = (xhi - xlo) / static_cast<double>(n);
    wy = (yhi - ylo) / static_cast<double>(n);
    grid.cell_to_string.assign(static_cast<size_t>(n) * n, 0xFFFF);
    for (unsigned ix = 0; ix < n; ++ix) {
        const double ax = sx + static_cast<double>(ix) * wx, bx = sx + static_cast<double>(ix + 1) * wx;
        for (unsigned iy = 0; iy < n; ++iy) {
            const double ay = sy + static_cast<double>(iy) * wy, by = sy + static_cast<double>(iy + 1) * wy;
            int owner = -1;
            for (size_t k = 0; k < strings.size(); ++k) {
                const DetString &s = strings[k];
                if (s.subdet != subdet) continue;
                if (touches(s.cx - s.reach, s.cx + s.reach, ax, bx) && touches(s.cy - s.reach, s.cy + s.reach, ay, by)) {
                    if (owner >= 0) return false;
                    owner = static_cast<int>(k);
                }
            }
            if (owner >= 0) grid.cell_to_string[iy * n + ix] = static_cast<uint16_t>(owner);
        }
    }
    grid.num_x = grid.num_y = static_cast<int>(n);
    return true;
}

struct Layering {
    double start = kNaN, height = kNaN;
    unsigned count = 0;
    std::vector<uint16_t> dom_of_layer;
};

// …GeometrySource.cxx:375-446
bool try_layering(const DetString &s, unsigned count, double radius, Layering &out)
{
    if (count == 0 || radius < 0.) return false;
    double lo = s.zlo - s.dz / 2., hi = s.zhi + s.dz / 2.; // hints
    grow_min(lo, s.zlo - radius);
    grow_max(hi, s.zhi + radius);
    out.start = lo;
    out.height = (hi - lo) / static_cast<double>(count);
    out.count = count;
    out.dom_of_layer.assign(count, 0xFFFF);
    for (unsigned i = 0; i < count; ++i) {
        const double a = out.start + static_cast<double>(i) * out.height;
        const double b = out.start + static_cast<double>(i + 1) * out.height;
        for (size_t d = 0; d < s.doms.size(); ++d) {
            if (!touches(s.doms[d].z - radius, s.doms[d].z + radius, a, b)) continue;
            if (out.dom_of_layer[i] != 0xFFFF) return false;
            out.dom_of_layer[i] = static_cast<uint16_t>(d);
        }
    }
    return true;
}

// …GeometrySource.cxx:273-342
bool fits_layering(const DetString &s, const Layering &lay, double radius)
{
    if (lay.count == 0 || radius < 0.) return false;
    size_t placed = 0;
    for (unsigned i = 0; i < lay.count; ++i) {
        const double a = lay.start + static_cast<double>(i) * lay.height;
        const double b = lay.start + static_cast<double>(i + 1) * lay.height;
        uint16_t expect = 0xFFFF;
        for (size_t d = 0; d < s.doms.size(); ++d) {
            if (!touches(s.doms[d].z - radius, s.doms[d].z + radius, a, b)) continue;
            if (expect != 0xFFFF) return false;
            expect = static_cast<uint16_t>(d);
            ++placed;
        }
        if (lay.dom_of_layer[i] != expect) return false;
    }
    return placed == s.doms.size();
}

void make_geometry(const clsimcu_geometry &g, GeometryTables &t)
{
    const size_t n = g.num_doms > 0 ? static_cast<size_t>(g.num_doms) : 0;
    if (n == 0) throw std::runtime_error("Empty geometry provided.");
    if (!g.string_id || !g.dom_id || !g.x || !g.y || !g.z || !g.subdetector) throw std::runtime_error("Geometry not set!");
    const double radius = g.om_radius;
    if (radius < 0.) throw std::runtime_error("Zero or negative OM radius.");

    // strings in (stringID, subdetector) order, DOMs in input order (:737-882)
    std::map<std::pair<int, int>, std::vector<size_t>> members;
    std::map<int, int> subdet_rank;
    for (size_t i = 0; i < n; ++i) {
        members[std::make_pair(g.string_id[i], g.subdetector[i])].push_back(i);
        subdet_rank[g.subdetector[i]] = 0;
    }
    if (members.size() >= 0xFFFF - 1) throw std::runtime_error("More than 65534 strings are not supported.");
    {
        int r = 0;
        for (auto &kv : subdet_rank) kv.second = r++;
    }
    if (subdet_rank.size() > 9) throw std::runtime_error("more than 9 subdetectors are currently not supported.");

    std::vector<DetString> strings;
    strings.reserve(members.size());
    double widest = kNaN;
    for (const auto &kv : members) {
        DetString s;
        s.id = kv.first.first;
        s.subdet = subdet_rank[kv.first.second];
        double prev_z = kNaN, gap_sum = 0.;
        bool have_gap = false;
        unsigned gaps = 0;
        for (size_t i : kv.second) {
            s.cx += g.x[i];
            s.cy += g.y[i];
            grow_max(s.zhi, g.z[i]);
            grow_min(s.zlo, g.z[i]);
            if (!std::isnan(prev_z)) {
                const double gap = std::abs(prev_z - g.z[i]);
                // spacings of >= 175 % of the running mean are a missing DOM: not averaged (:821-832)
                if (!have_gap || gap < 1.75 * gap_sum / static_cast<double>(gaps)) {
                    gap_sum += gap;
                    ++gaps;
                    have_gap = true;
                }
            }
            prev_z = g.z[i];
            if (s.doms.size() >= 0xFFFF - 1) throw std::runtime_error("Dom numbers >= 65535 are not supported!");
            s.doms.push_back(Dom{g.dom_id[i], g.x[i], g.y[i], g.z[i]});
        }
        s.cx /= static_cast<double>(s.doms.size());
        s.cy /= static_cast<double>(s.doms.size());
        s.dz = gap_sum / static_cast<double>(gaps);
        for (const Dom &d : s.doms) {
            const double ox = s.cx - d.x, oy = s.cy - d.y;
            const double r = std::sqrt(ox * ox + oy * oy) + radius;
            grow_max(s.reach, r);
            grow_max(widest, r);
        }
        strings.push_back(std::move(s));
    }

    // xy cell grids: smallest n x n with at most one string per cell (:913-949)
    t.grids.resize(subdet_rank.size());
    for (size_t sd = 0; sd < t.grids.size(); ++sd) {
        double sx, sy, wx, wy;
        unsigned cells = 1;
        while (!try_grid(strings, static_cast<int>(sd), cells, t.grids[sd], sx, sy, wx, wy)) {
            if (++cells >= 1000) throw std::runtime_error("Could not generate a x-y cell division for your subdetector.");
        }
        t.grids[sd].start_x = float_literal(sx);
        t.grids[sd].start_y = float_literal(sy);
        t.grids[sd].width_x = float_literal(wx);
        t.grids[sd].width_y = float_literal(wy);
    }

    // z layerings shared between strings ("string sets", :956-1082)
    std::vector<Layering> sets;
    t.string_set.resize(strings.size());
    for (size_t k = 0; k < strings.size(); ++k) {
        const DetString &s = strings[k];
        size_t found = sets.size();
        for (size_t q = 0; q < sets.size(); ++q) {
            if (fits_layering(s, sets[q], radius)) { found = q; break; }
        }
        if (found == sets.size()) {
            if (sets.size() + 1 >= 0xFF) throw std::runtime_error("Not more than 255 different string layer divisions (\"string sets\") are supported!");
            Layering lay;
            const unsigned guess = static_cast<unsigned>((s.zhi - s.zlo + s.dz) / s.dz);
            bool ok = try_layering(s, guess, radius, lay);
            if (!ok) ok = try_layering(s, guess + 1, radius, lay);
            for (unsigned c = 1; !ok; ++c) {
                if (c >= 1000) throw std::runtime_error("There does not seem to be a possible layer division for your string.");
                ok = try_layering(s, c, radius, lay);
            }
            sets.push_back(lay);
        }
        t.string_set[k] = static_cast<uint8_t>(found);
    }
    unsigned deepest = 0;
    for (const Layering &lay : sets) deepest = std::max(deepest, lay.count);
    t.num_sets = static_cast<int>(sets.size());
    t.max_layers = static_cast<int>(deepest);
    // device buffer: padded up to the next multiple of 64 entries (:1104-1112)
    const size_t padded = ((sets.size() * deepest) / 64 + 1) * 64;
    t.layer_to_dom.assign(padded, 0xFFFF);
    for (size_t q = 0; q < sets.size(); ++q) {
        std::copy(sets[q].dom_of_layer.begin(), sets[q].dom_of_layer.end(), t.layer_to_dom.begin() + q * deepest);
        t.set_layer_count.push_back(static_cast<uint16_t>(sets[q].count));
        t.set_start_z.push_back(float_literal(sets[q].start));
        t.set_layer_height.push_back(float_literal(sets[q].height));
    }

    t.num_strings = static_cast<int>(strings.size());
    t.om_radius = float_literal(radius);
    t.string_max_radius = float_literal(widest);
    for (const DetString &s : strings) {
        t.string_x.push_back(float_literal(s.cx));
        t.string_y.push_back(float_literal(s.cy));
        t.string_radius.push_back(float_literal(s.reach));
        t.string_min_z.push_back(float_literal(s.zlo));
        t.string_max_z.push_back(float_literal(s.zhi));
        t.string_index_to_id.push_back(s.id);
        std::vector<uint32_t> ids;
        for (const Dom &d : s.doms) ids.push_back(d.id);
        t.dom_index_to_id.push_back(std::move(ids));
        t.max_doms_per_string = std::max(t.max_doms_per_string, static_cast<int>(s.doms.size()));
    }

    // DOM positions as shared per-string templates, lateral offsets as int16 (:499-709)
    struct Template {
        std::vector<double> dx, dy, z;
        size_t flat_start = 0;
    };
    std::vector<Template> templates;
    std::vector<size_t> template_of(strings.size());
    const double tolerance = 1e-1 * 1e-3; // 0.1 mm
    for (size_t k = 0; k < strings.size(); ++k) {
        const DetString &s = strings[k];
        size_t match = templates.size();
        for (size_t q = 0; q < templates.size() && match == templates.size(); ++q) {
            const Template &tp = templates[q];
            if (tp.z.size() != s.doms.size()) continue;
            bool same = true;
            for (size_t j = 0; j < s.doms.size() && same; ++j) {
                same = !(std::abs(tp.dx[j] - (s.doms[j].x - s.cx)) > tolerance) &&
                       !(std::abs(tp.dy[j] - (s.doms[j].y - s.cy)) > tolerance) &&
                       !(std::abs(tp.z[j] - s.doms[j].z) > tolerance);
            }
            if (same) match = q;
        }
        if (match == templates.size()) {
            Template tp;
            for (const Dom &d : s.doms) {
                tp.dx.push_back(d.x - s.cx);
                tp.dy.push_back(d.y - s.cy);
                tp.z.push_back(d.z);
            }
            templates.push_back(std::move(tp));
        }
        template_of[k] = match;
    }
    double span_x = kNaN, span_y = kNaN;
    size_t flat = 0;
    for (Template &tp : templates) {
        tp.flat_start = flat;
        flat += tp.z.size();
        for (size_t j = 0; j < tp.z.size(); ++j) {
            grow_max(span_x, std::abs(tp.dx[j]));
            grow_max(span_y, std::abs(tp.dy[j]));
        }
    }
    const double unit_x = span_x / 32767., unit_y = span_y / 32767.;
    t.tmpl_scale_x = float_literal(unit_x);
    t.tmpl_scale_y = float_literal(unit_y);
    auto quantise = [](double offset, double unit) -> int16_t {
        const double q = offset / unit;
        // perfectly straight strings give 0/0; the scale is 0 then, any value decodes to the mean
        return std::isfinite(q) ? static_cast<int16_t>(q) : static_cast<int16_t>(0);
    };
    for (const Template &tp : templates) {
        for (size_t j = 0; j < tp.z.size(); ++j) {
            t.tmpl_dx.push_back(quantise(tp.dx[j], unit_x));
            t.tmpl_dy.push_back(quantise(tp.dy[j], unit_y));
            t.tmpl_z.push_back(float_literal(tp.z[j]));
        }
    }
    for (size_t k = 0; k < strings.size(); ++k) {
        t.string_tmpl_start.push_back(static_cast<uint32_t>(templates[template_of[k]].flat_start));
        t.string_mean_x.push_back(float_literal(strings[k].cx));
        t.string_mean_y.push_back(float_literal(strings[k].cy));
    }
}

// ------------------------------------------------------------------ JSON description
template <class T> void put_floats(std::ostringstream &o, const T &v)
{
    o << "[";
    bool first = true;
    for (float f : v) {
        char b[40];
        std::snprintf(b, sizeof b, "%.9g", f);
        o << (first ? "" : ",") << b;
        first = false;
    }
    o << "]";
}
template <class T> void put_ints(std::ostringstream &o, const T &v)
{
    o << "[";
    bool first = true;
    for (auto e : v) {
        o << (first ? "" : ",") << static_cast<long long>(e);
        first = false;
    }
    o << "]";
}

} // namespace

void build_scene_tables(const clsimcu_config &c, SceneTables &out)
{
    // I3CLSimStepToPhotonConverterOpenCL::Compile preconditions (…OpenCL.cxx:492-508)
    if (c.num_wlen_generators < 1 || !c.wlen_generators) throw std::runtime_error("WlenGenerators not set!");
    if (c.save_all_photons && c.stop_detected_photons)
        throw std::runtime_error("Internal error: both the saveAllPhotons and stopDetectedPhotons options are set at the same time.");
    if (c.photon_history_entries < 0) throw std::runtime_error("photon_history_entries must not be negative");
    if (c.photon_history_entries > 32) throw std::runtime_error("photon_history_entries: at most 32 scatter points per photon are kept on the device");
    make_medium(c.medium, out.medium);
    out.generators.clear();
    for (int i = 0; i < c.num_wlen_generators; ++i) out.generators.push_back(make_generator(c.wlen_generators[i]));
    out.bias.kind = c.wlen_bias.kind;
    if (c.wlen_bias.kind == CLSIMCU_BIAS_TABLE) {
        if (c.wlen_bias.n < 2 || !c.wlen_bias.v) throw std::runtime_error("values must contain at least 2 elements!");
        if (!(c.wlen_bias.dx > 0.)) throw std::runtime_error("wlenStep must not be <= 0!");
        out.bias.n = c.wlen_bias.n;
        out.bias.x0 = float_literal(c.wlen_bias.x0);
        out.bias.dx = float_literal(c.wlen_bias.dx);
        out.bias.v.resize(c.wlen_bias.n);
        for (int i = 0; i < c.wlen_bias.n; ++i) out.bias.v[i] = float_literal(c.wlen_bias.v[i]);
    } else if (c.wlen_bias.kind == CLSIMCU_BIAS_CONSTANT) {
        out.bias.value = float_literal(c.wlen_bias.value);
    } else {
        throw std::runtime_error("WlenBias not set!");
    }
    out.stop_detected = c.stop_detected_photons != 0;
    out.save_all = c.save_all_photons != 0;
    out.prescale = float_literal(c.save_all_photons_prescale);
    out.fixed_abs = !std::isnan(c.fixed_number_of_absorption_lengths);
    out.fixed_abs_lens = out.fixed_abs ? float_literal(c.fixed_number_of_absorption_lengths) : 0.f;
    out.pancake = (c.pancake_factor != 1.); // PANCAKE_FACTOR only defined when != 1 (…OpenCL.cxx:434-440)
    out.pancake_factor = float_literal(c.pancake_factor);
    out.history_entries = c.photon_history_entries;
    out.has_geometry = !out.save_all;
    if (out.has_geometry) make_geometry(c.geometry, out.geometry);
}

std::string describe_scene_tables(const SceneTables &s)
{
    const GeometryTables &g = s.geometry;
    std::ostringstream o;
    char b[64];
    o << "{\"num_strings\":" << g.num_strings;
    std::snprintf(b, sizeof b, "%.9g", g.om_radius);
    o << ",\"om_radius\":" << b;
    std::snprintf(b, sizeof b, "%.9g", g.string_max_radius);
    o << ",\"string_max_radius\":" << b;
    o << ",\"string_pos_x\":"; put_floats(o, g.string_x);
    o << ",\"string_pos_y\":"; put_floats(o, g.string_y);
    o << ",\"string_min_z\":"; put_floats(o, g.string_min_z);
    o << ",\"string_max_z\":"; put_floats(o, g.string_max_z);
    o << ",\"string_in_set\":"; put_ints(o, g.string_set);
    o << ",\"num_sets\":" << g.num_sets << ",\"max_layers\":" << g.max_layers;
    o << ",\"layer_num\":"; put_ints(o, g.set_layer_count);
    o << ",\"layer_start_z\":"; put_floats(o, g.set_start_z);
    o << ",\"layer_height\":"; put_floats(o, g.set_layer_height);
    o << ",\"layer_to_om\":"; put_ints(o, g.layer_to_dom);
    o << ",\"cells\":[";
    for (size_t i = 0; i < g.grids.size(); ++i) {
        const CellGridTable &c = g.grids[i];
        o << (i ? "," : "") << "{\"num_x\":" << c.num_x << ",\"num_y\":" << c.num_y << ",\"start_width\":";
        put_floats(o, std::vector<float>{c.start_x, c.start_y, c.width_x, c.width_y});
        o << ",\"index\":"; put_ints(o, c.cell_to_string);
        o << "}";
    }
    o << "],\"max_dom_index\":" << g.max_doms_per_string;
    o << ",\"tmpl_mul\":"; put_floats(o, std::vector<float>{g.tmpl_scale_x, g.tmpl_scale_y});
    o << ",\"tmpl_x\":"; put_ints(o, g.tmpl_dx);
    o << ",\"tmpl_y\":"; put_ints(o, g.tmpl_dy);
    o << ",\"tmpl_z\":"; put_floats(o, g.tmpl_z);
    o << ",\"string_tmpl_start\":"; put_ints(o, g.string_tmpl_start);
    o << ",\"string_mean_x\":"; put_floats(o, g.string_mean_x);
    o << ",\"string_mean_y\":"; put_floats(o, g.string_mean_y);
    o << ",\"string_index_to_id\":"; put_ints(o, g.string_index_to_id);
    o << ",\"dom_index_to_id\":[";
    for (size_t i = 0; i < g.dom_index_to_id.size(); ++i) {
        o << (i ? "," : "");
        put_ints(o, g.dom_index_to_id[i]);
    }
    o << "],\"medium\":{\"b400\":"; put_floats(o, s.medium.b400);
    o << ",\"a_dust400\":"; put_floats(o, s.medium.a_dust400);
    o << ",\"delta_tau\":"; put_floats(o, s.medium.delta_tau);
    o << "},\"wlen_generators\":[";
    for (size_t i = 0; i < s.generators.size(); ++i) {
        const WlenGeneratorTable &w = s.generators[i];
        o << (i ? "," : "") << "{\"kind\":" << w.kind << ",\"beta\":"; put_floats(o, w.density);
        o << ",\"acu\":"; put_floats(o, w.cumulative);
        o << ",\"xs\":"; put_floats(o, w.xs);
        o << "}";
    }
    o << "]}";
    return o.str();
}

// ------------------------------------------------------------------ safe primes
namespace {

// Montgomery arithmetic modulo an odd 64-bit n
struct Mont {
    uint64_t n, ninv, r2; // ninv = -n^-1 mod 2^64, r2 = 2^128 mod n
    explicit Mont(uint64_t mod) : n(mod)
    {
        uint64_t inv = mod; // Newton iteration for n^-1 mod 2^64
        for (int i = 0; i < 5; ++i) inv *= 2 - mod * inv;
        ninv = ~inv + 1;
        const unsigned __int128 r = (static_cast<unsigned __int128>(1) << 64) % mod;
        r2 = static_cast<uint64_t>((r * r) % mod);
    }
    uint64_t reduce(unsigned __int128 t) const
    {
        const uint64_t m = static_cast<uint64_t>(t) * ninv;
        const unsigned __int128 u = (t + static_cast<unsigned __int128>(m) * n) >> 64; // may wrap past 2^128
        uint64_t res = static_cast<uint64_t>(u);
        // detect the carry lost in the 128-bit addition
        const unsigned __int128 mn = static_cast<unsigned __int128>(m) * n;
        const bool carry = (t + mn) < t;
        if (carry || res >= n) res -= n;
        return res;
    }
    uint64_t mul(uint64_t a, uint64_t b) const { return reduce(static_cast<unsigned __int128>(a) * b); }
    uint64_t to(uint64_t a) const { return mul(a % n, r2); }
};

const uint32_t kSievePrimes[] = {3,   5,   7,   11,  13,  17,  19,  23,  29,  31,  37,  41,  43,  47,  53,  59,  61,  67,  71,  73,
                                 79,  83,  89,  97,  101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179,
                                 181, 191, 193, 197, 199, 211, 223, 227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283};

bool miller_rabin(uint64_t n)
{
    if (n < 2) return false;
    if (n % 2 == 0) return n == 2;
    for (uint32_t p : kSievePrimes) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    const Mont mt(n);
    uint64_t d = n - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; ++s; }
    const uint64_t one = mt.to(1), minus_one = n - one;
    // deterministic for n < 2^64 with the first twelve primes as bases
    static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t base : bases) {
        uint64_t acc = one, sq = mt.to(base);
        for (uint64_t e = d; e; e >>= 1) {
            if (e & 1) acc = mt.mul(acc, sq);
            sq = mt.mul(sq, sq);
        }
        if (acc == one || acc == minus_one) continue;
        bool witness = true;
        for (int i = 1; i < s && witness; ++i) {
            acc = mt.mul(acc, acc);
            if (acc == minus_one) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// a is a valid multiplier iff n2 = a*2^32-1 and n1 = (n2-1)/2 are both prime
inline bool is_safe_multiplier(uint64_t a)
{
    const uint64_t n2 = (a << 32) - 1;
    const uint64_t n1 = (n2 - 1) >> 1;
    // cheap joint sieve first
    for (uint32_t p : kSievePrimes) {
        if (n2 % p == 0 || n1 % p == 0) return false;
    }
    return miller_rabin(n2) && miller_rabin(n1);
}

const uint64_t kFirstCandidate = 4294967118ull; // make_safeprimes/main.cxx:59
const char kCacheTag[] = "safeprimes_base32"; // 17 bytes, mwcrng_init.h:67-75

std::mutex g_prime_mutex;
std::vector<uint32_t> g_primes;       // rows known so far, from row 0
uint64_t g_next_candidate = kFirstCandidate;

void extend_primes(uint64_t rows_needed)
{
    unsigned workers = std::thread::hardware_concurrency();
    if (workers == 0) workers = 4;
    workers = std::min(workers, 32u);
    while (g_primes.size() < rows_needed) {
        // scan a window of candidates in parallel, keep descending order
        const uint64_t missing = rows_needed - g_primes.size();
        uint64_t window = std::max<uint64_t>(65536, missing * 800);
        window = std::min<uint64_t>(window, g_next_candidate);
        if (window == 0) throw std::runtime_error("ran out of MWC multiplier candidates");
        const uint64_t hi = g_next_candidate, lo = hi - window; // candidates (lo, hi]
        std::vector<std::vector<uint32_t>> found(workers);
        std::vector<std::thread> pool;
        const uint64_t chunk = (window + workers - 1) / workers;
        for (unsigned w = 0; w < workers; ++w) {
            pool.emplace_back([&, w]() {
                const uint64_t top = hi - static_cast<uint64_t>(w) * chunk;
                const uint64_t bottom = (top > lo + chunk) ? top - chunk : lo;
                if (top <= lo) return;
                for (uint64_t a = top; a > bottom; --a) {
                    if (is_safe_multiplier(a)) found[w].push_back(static_cast<uint32_t>(a));
                }
            });
        }
        for (std::thread &th : pool) th.join();
        for (unsigned w = 0; w < workers; ++w) g_primes.insert(g_primes.end(), found[w].begin(), found[w].end());
        g_next_candidate = lo;
    }
}

void load_cache(const std::string &path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) return;
    char tag[17];
    in.read(tag, 17);
    if (!in || std::memcmp(tag, kCacheTag, 17) != 0) return;
    std::vector<uint32_t> rows;
    int64_t v;
    while (in.read(reinterpret_cast<char *>(&v), sizeof v)) {
        if (v <= 0 || v > 0xffffffffll) return; // corrupt
        rows.push_back(static_cast<uint32_t>(v));
    }
    if (rows.size() > g_primes.size() && !rows.empty() && rows[0] == kFirstCandidate) {
        g_primes.swap(rows);
        g_next_candidate = static_cast<uint64_t>(g_primes.back()) - 1;
    }
}

void store_cache(const std::string &path)
{
    const std::string tmp = path + ".tmp";
    {
        std::ofstream out(tmp, std::ios::binary | std::ios::trunc);
        if (!out) return;
        out.write(kCacheTag, 17);
        for (uint32_t a : g_primes) {
            const int64_t v = a;
            out.write(reinterpret_cast<const char *>(&v), sizeof v);
        }
        if (!out) return;
    }
    std::rename(tmp.c_str(), path.c_str());
}

} // namespace

void safeprime_multipliers(uint64_t first, uint64_t n, uint32_t *out, const std::string &cache_path)
{
    std::lock_guard<std::mutex> lock(g_prime_mutex);
    const uint64_t need = first + n;
    if (g_primes.size() < need && !cache_path.empty()) load_cache(cache_path);
    if (g_primes.size() < need) {
        extend_primes(need);
        if (!cache_path.empty()) store_cache(cache_path);
    }
    std::copy(g_primes.begin() + first, g_primes.begin() + need, out);
}

// ------------------------------------------------------------------ collision map of the fast kernel
CollisionMap build_collision_map(const GeometryTables &g, int pixel_budget)
{
    if (g.num_strings < 1) throw std::runtime_error("collision map: no strings");
    CollisionMap m;
    float xlo = g.string_x[0], xhi = g.string_x[0], ylo = g.string_y[0], yhi = g.string_y[0];
    for (int i = 0; i < g.num_strings; ++i) {
        xlo = std::min(xlo, g.string_x[i]); xhi = std::max(xhi, g.string_x[i]);
        ylo = std::min(ylo, g.string_y[i]); yhi = std::max(yhi, g.string_y[i]);
    }
    // one pixel of margin is enough: a point outside the map is farther from every string
    // than its projection onto the map, so the border pixels' bounds hold for it
    float pixel = 4.f;
    for (;;) {
        const double w = (xhi - xlo) + 2.0 * pixel, h = (yhi - ylo) + 2.0 * pixel;
        if (std::ceil(w / pixel) * std::ceil(h / pixel) <= static_cast<double>(pixel_budget)) break;
        pixel *= 1.05f;
    }
    m.pixel = pixel;
    m.x0 = xlo - pixel;
    m.y0 = ylo - pixel;
    m.inv_pixel = 1.f / pixel;
    m.off_x = -m.x0 * m.inv_pixel;
    m.off_y = -m.y0 * m.inv_pixel;
    m.nx = static_cast<int>(std::ceil((xhi - xlo + 2 * pixel) / pixel));
    m.ny = static_cast<int>(std::ceil((yhi - ylo + 2 * pixel) / pixel));
    const double half_diag = 0.5 * std::sqrt(2.0) * pixel + 1e-2; // + slack for fp32 pixel assignment
    const double min_range = 1.0; // below this the map cannot limit flights sensibly: cell walk
    m.info.resize(static_cast<size_t>(m.nx) * m.ny);
    for (int iy = 0; iy < m.ny; ++iy) {
        for (int ix = 0; ix < m.nx; ++ix) {
            const double cx = m.x0 + (ix + 0.5) * pixel, cy = m.y0 + (iy + 0.5) * pixel;
            double best = 1e30, second = 1e30;
            int who = 0;
            for (int k = 0; k < g.num_strings; ++k) {
                const double d = std::hypot(cx - g.string_x[k], cy - g.string_y[k]);
                if (d < best) { second = best; best = d; who = k; }
                else if (d < second) second = d;
            }
            // how far a photon anywhere in this pixel may fly before a string other than
            // `who` can come within the collision radius
            double range = std::min(second, 1e9) - half_diag - g.string_max_radius - 1e-2;
            const float rf = static_cast<float>(range);
            uint32_t bits;
            std::memcpy(&bits, &rf, 4);
            bits &= 0xffff0000u; // truncation of a positive float rounds down: the bound stays a lower bound
            uint32_t low = static_cast<uint32_t>(who) << 4; // byte offset of the string's 16-byte record
            if (range < min_range) {
                // strings too dense for the pixel size: no range (+inf) and no string (the record behind the
                // last one, which the kernel fills with NaN): every leg takes the reference's cell walk
                bits = 0x7f800000u;
                low = static_cast<uint32_t>(g.num_strings) << 4;
            }
            m.info[static_cast<size_t>(iy) * m.nx + ix] = low | bits;
        }
    }
    return m;
}

std::string describe_collision_map(const GeometryTables &g, const CollisionMap &m)
{
    std::ostringstream o;
    o.precision(9);
    o << "{\"nx\":" << m.nx << ",\"ny\":" << m.ny << ",\"x0\":" << m.x0 << ",\"y0\":" << m.y0 << ",\"pixel\":" << m.pixel
      << ",\"inv_pixel\":" << m.inv_pixel << ",\"off_x\":" << m.off_x << ",\"off_y\":" << m.off_y
      << ",\"num_strings\":" << g.num_strings << ",\"string_max_radius\":" << g.string_max_radius;
    o << ",\"string_pos_x\":"; put_floats(o, g.string_x);
    o << ",\"string_pos_y\":"; put_floats(o, g.string_y);
    o << ",\"info\":"; put_ints(o, m.info);
    o << "}";
    return o.str();
}

void seed_rng_states(uint64_t seed, const uint32_t *a, uint64_t *x, size_t n)
{
    // splitmix64 stands in for I3RandomService::Integer(0xffffffff); the acceptance rule is the
    // reference's: x != 0, hi32(x) < a-1, lo32(x) < 0xffffffff.
    uint64_t state = seed;
    auto next = [&state]() {
        uint64_t z = (state += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    for (size_t i = 0; i < n; ++i) {
        for (;;) {
            const uint64_t r = next();
            const uint32_t hi = static_cast<uint32_t>(r >> 32), lo = static_cast<uint32_t>(r);
            const uint64_t cand = (static_cast<uint64_t>(hi) << 32) + lo;
            if (cand != 0 && hi < a[i] - 1 && lo < 0xffffffffu) {
                x[i] = cand;
                break;
            }
        }
    }
}

} // namespace clsimcu
