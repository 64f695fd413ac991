// Stand-in for phys-services/I3SummaryService.h: the summary is an I3MapStringDouble (dataclasses/I3Map.h here).
#include "dataclasses/I3Map.h"
