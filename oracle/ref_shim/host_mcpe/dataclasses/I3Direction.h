#include "dataclasses/I3Position.h"
