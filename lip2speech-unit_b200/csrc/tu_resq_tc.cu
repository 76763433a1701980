// Translation unit that owns the resq_tc kernel instantiations (the skewed whole-ResBlock schedule).
#define L2S_TU_RESQ_TC
#include "resq_tc.cuh"
