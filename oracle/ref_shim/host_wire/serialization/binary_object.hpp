// Stand-in: make_binary_object lives in the wire build's icetray/serialization.h.
#include "icetray/serialization.h"
