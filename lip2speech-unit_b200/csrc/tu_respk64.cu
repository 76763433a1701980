// Translation unit that owns the C = 64 instantiations of the time-packed whole-ResBlock kernel.
#define L2S_TU_RESPK_C 64
#include <vector>
#include "respk_tc.cuh"
