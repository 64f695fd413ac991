// oracle/_ref/libclsim_ref_geometry.so: the reference's OWN geometry source generator
// (private/opencl/I3CLSimHelperGenerateGeometrySource.cxx, 1279 lines), compiled unmodified from where it lies under
// /root/reference, behind a C entry point.  It returns the OpenCL source text the reference would hand to its kernel
// compiler for a detector, plus the three side buffers; tests/test_ref_geometry.py parses the text and holds the
// product's tables (csrc/tables.cpp) and the oracle's against it, number for number.
//
// Test infrastructure.  Nothing under clsim_b200/ links or loads this.  Built by oracle/Makefile only where the reference
// tree exists; headers the reference expects from IceTray / boost / OpenCL are the stand-ins under ref_shim/host/.
#include <cstring>
#include <string>
#include <vector>

#include "opencl/I3CLSimHelperGenerateGeometrySource.cxx"   // -I $(REFERENCE)/private

namespace {

// the abstract interface of public/clsim/I3CLSimSimpleGeometry.h over plain arrays
class ArrayGeometry : public I3CLSimSimpleGeometry {
public:
    double om_radius = 0.;
    std::vector<int32_t> string_ids;
    std::vector<uint32_t> dom_ids;
    std::vector<double> x, y, z;
    std::vector<std::string> subdetectors;

    std::size_t size() const override { return string_ids.size(); }
    double GetOMRadius() const override { return om_radius; }
    const std::vector<int32_t> &GetStringIDVector() const override { return string_ids; }
    const std::vector<uint32_t> &GetDomIDVector() const override { return dom_ids; }
    const std::vector<double> &GetPosXVector() const override { return x; }
    const std::vector<double> &GetPosYVector() const override { return y; }
    const std::vector<double> &GetPosZVector() const override { return z; }
    const std::vector<std::string> &GetSubdetectorVector() const override { return subdetectors; }
    int32_t GetStringID(std::size_t pos) const override { return string_ids.at(pos); }
    uint32_t GetDomID(std::size_t pos) const override { return dom_ids.at(pos); }
    double GetPosX(std::size_t pos) const override { return x.at(pos); }
    double GetPosY(std::size_t pos) const override { return y.at(pos); }
    double GetPosZ(std::size_t pos) const override { return z.at(pos); }
    std::string GetSubdetector(std::size_t pos) const override { return subdetectors.at(pos); }
};

thread_local std::string t_text, t_error;
thread_local std::vector<unsigned short> t_layer_to_om;
thread_local std::vector<int> t_string_ids;
thread_local std::vector<unsigned int> t_dom_ids_flat, t_dom_ids_start;

} // namespace

extern "C" {

// -> 0, or -1 with ref_geometry_error() set (the reference's log_fatal text).  subdetectors: n C strings, or NULL for "".
int ref_geometry_generate(size_t n, const int32_t *string_ids, const uint32_t *dom_ids, const double *x, const double *y, const double *z,
                          const char *const *subdetectors, double om_radius)
{
    ArrayGeometry g;
    g.om_radius = om_radius;
    g.string_ids.assign(string_ids, string_ids + n);
    g.dom_ids.assign(dom_ids, dom_ids + n);
    g.x.assign(x, x + n);
    g.y.assign(y, y + n);
    g.z.assign(z, z + n);
    g.subdetectors.resize(n);
    for (size_t i = 0; i < n; ++i) g.subdetectors[i] = subdetectors ? subdetectors[i] : "";
    t_error.clear();
    try {
        std::vector<std::vector<unsigned int>> dom_ids_per_string;
        t_text = I3CLSimHelper::GenerateGeometrySource(g, t_layer_to_om, t_string_ids, dom_ids_per_string);
        t_dom_ids_flat.clear();
        t_dom_ids_start.assign(1, 0u);
        for (const auto &v : dom_ids_per_string) {
            t_dom_ids_flat.insert(t_dom_ids_flat.end(), v.begin(), v.end());
            t_dom_ids_start.push_back(static_cast<unsigned int>(t_dom_ids_flat.size()));
        }
    } catch (const std::exception &ex) {
        t_error = ex.what();
        return -1;
    }
    return 0;
}

const char *ref_geometry_error(void) { return t_error.c_str(); }
const char *ref_geometry_text(void) { return t_text.c_str(); }
size_t ref_geometry_layer_to_om(const unsigned short **p) { *p = t_layer_to_om.data(); return t_layer_to_om.size(); }
size_t ref_geometry_string_ids(const int **p) { *p = t_string_ids.data(); return t_string_ids.size(); }
size_t ref_geometry_dom_ids(const unsigned int **flat, const unsigned int **start)
{
    *flat = t_dom_ids_flat.data();
    *start = t_dom_ids_start.data();
    return t_dom_ids_start.size() - 1;
}

} // extern "C"
