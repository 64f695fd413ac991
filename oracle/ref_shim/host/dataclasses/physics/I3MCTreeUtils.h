// Stand-in for dataclasses/physics/I3MCTreeUtils.h (included, not used, by private/clsim/I3CLSimModuleHelper.cxx).
