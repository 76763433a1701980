// Translation unit that owns the C = 32 instantiations of the time-packed whole-ResBlock kernel.
#define L2S_TU_RESPK_C 32
#include <vector>
#include "respk_tc.cuh"
