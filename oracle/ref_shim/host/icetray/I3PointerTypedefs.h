#include "icetray/I3TrayHeaders.h"
