// Library identification.
#include "common.cuh"
#include "../../include/pfpp.h"

extern "C" int pfpp_version(void) { return PFPP_VERSION; }
