// Translation unit that owns the res_tc kernel instantiations (built in parallel with the others by build.py).
#define L2S_TU_RES_TC
#include "res_tc.cuh"
