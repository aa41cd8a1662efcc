#include "../../include/pfpp.h"

#ifndef PFPP_BUILD_ID
#define PFPP_BUILD_ID 0ull
#endif

extern "C" int pfpp_version(void) { return PFPP_VERSION; }
// FNV-1a 64 over the library's sources, computed by build.py at build time (-DPFPP_BUILD_ID=...)
extern "C" unsigned long long pfpp_build_id(void) { return PFPP_BUILD_ID; }
