// Stand-in for dataclasses/physics/I3MCTree.h (included by the converter's header, not used).
