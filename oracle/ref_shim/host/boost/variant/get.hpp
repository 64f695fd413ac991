// Stand-in for boost/variant/get.hpp (included, not used, by private/clsim/I3CLSimModuleHelper.cxx).
