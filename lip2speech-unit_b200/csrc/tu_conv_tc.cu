// Translation unit that owns the conv_tc kernel instantiations (built in parallel with the others by build.py).
#define L2S_TU_CONV_TC
#include "conv_tc.cuh"
