// Stand-in for icetray/I3Module.h (+ I3Context, I3Frame, I3ConditionalModule): just enough of the module protocol to
// configure ONE module by hand and hand it frames.  TEST INFRASTRUCTURE (oracle/_ref/libclsim_ref_mcpe.so).
//   parameters: AddParameter registers nothing (the member already holds its default); GetParameter overwrites the member
//               with the value the driver stored under that name, if any.
//   frames:     a map name -> shared_ptr<const I3FrameObject>; Get<shared_ptr<const T>> is a dynamic cast.
#ifndef CLSIM_REF_SHIM_I3MODULE_H
#define CLSIM_REF_SHIM_I3MODULE_H
#include <any>
#include <map>
#include <sstream>
#include <string>
#include "icetray/I3TrayHeaders.h"
#include "icetray/I3FrameObject.h"
#include "icetray/OMKey.h"

#define log_info_stream(x) ((void)0)
#define log_warn_stream(x) ((void)0)
#define log_debug_stream(x) ((void)0)
#define log_trace_stream(x) ((void)0)
#define log_fatal_stream(x)                    \
    do {                                       \
        std::ostringstream ref_shim_os;        \
        ref_shim_os << x;                      \
        ref_shim::fatal("%s", ref_shim_os.str().c_str()); \
    } while (0)
#define SET_LOGGER(name)
#define I3_MODULE(M)

class I3Context {
public:
    template <class T> T Get() const { return T(); }                       // (no services installed)
    template <class T> T Get(const std::string &) const { return T(); }
};

class I3Frame {
public:
    template <class P> P Get(const std::string &name) const
    {
        std::map<std::string, boost::shared_ptr<const I3FrameObject> >::const_iterator it = objects_.find(name);
        if (it == objects_.end()) return P();
        return boost::dynamic_pointer_cast<typename P::element_type>(it->second);
    }
    template <class T> void Put(const std::string &name, boost::shared_ptr<T> obj) { objects_[name] = obj; }
private:
    std::map<std::string, boost::shared_ptr<const I3FrameObject> > objects_;
};
I3_POINTER_TYPEDEFS(I3Frame);

class I3Module {
public:
    explicit I3Module(const I3Context &context) : context_(context) {}
    virtual ~I3Module() {}
    virtual void Configure() {}
    virtual void Finish() {}
    const std::string GetName() const { return "oracle_ref"; }
    template <class T> void Set(const std::string &name, const T &value) { overrides_[name] = value; }   // the driver's side
protected:
    template <class T> void AddParameter(const std::string &, const std::string &, const T &) {}
    void AddParameter(const std::string &, const std::string &) {}
    template <class T> void GetParameter(const std::string &name, T &value) const
    {
        std::map<std::string, std::any>::const_iterator it = overrides_.find(name);
        if (it != overrides_.end()) value = std::any_cast<T>(it->second);
    }
    void AddOutBox(const std::string &) {}
    void PushFrame(I3FramePtr) {}
    const I3Context &context_;
private:
    std::map<std::string, std::any> overrides_;
};
#endif
