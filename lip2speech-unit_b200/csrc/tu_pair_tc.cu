// Translation unit that owns the pair_tc kernel instantiations (built in parallel with the others by build.py).
#define L2S_TU_PAIR_TC
#include "pair_tc.cuh"
