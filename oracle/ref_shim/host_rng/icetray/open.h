// Stand-in for icetray/open.h: I3::dataio::open on a boost::iostreams::filtering_istream (here: a plain file stream that can
// be opened again; the gzip filter of the real one is not needed for the text table).
#ifndef CLSIM_REF_SHIM_OPEN_H
#define CLSIM_REF_SHIM_OPEN_H
#include <fstream>
#include <string>
#include "icetray/I3TrayHeaders.h"
namespace boost { namespace iostreams {
class filtering_istream : public std::ifstream {};
}} // namespace boost::iostreams
namespace I3 { namespace dataio {
inline void open(boost::iostreams::filtering_istream &s, const std::string &name)
{
    if (s.is_open()) s.close();
    s.clear();
    s.std::ifstream::open(name.c_str(), std::ios::in | std::ios::binary);
}
}} // namespace I3::dataio
#endif
