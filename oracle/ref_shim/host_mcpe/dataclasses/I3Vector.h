#include "dataclasses/I3Map.h"
