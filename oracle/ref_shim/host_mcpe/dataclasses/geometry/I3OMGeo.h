#include "dataclasses/geometry/I3Geometry.h"
