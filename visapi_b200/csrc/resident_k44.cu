// resident search kernel, variant <MAXW = 4, MAXEW = 4> (see resident_variant.cuh)
#define BP_RESIDENT_MAXW 4
#define BP_RESIDENT_MAXEW 4
#include "resident_variant.cuh"
