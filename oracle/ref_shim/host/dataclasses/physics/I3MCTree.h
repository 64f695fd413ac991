// Stand-in for dataclasses/physics/I3MCTree.h (included, not used, by private/clsim/I3CLSimModuleHelper.cxx).
