#include "simclasses/I3Photon.h"
