// Library-level entry points of libocc_b200.so.
#include "occ_common.cuh"

extern "C" int occ_version() { return 100; }  // 0.1.0
