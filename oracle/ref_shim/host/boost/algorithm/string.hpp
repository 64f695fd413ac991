// Stand-in for boost/algorithm/string.hpp (included, not used, by private/clsim/I3CLSimModuleHelper.cxx).
