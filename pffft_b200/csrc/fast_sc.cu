// fast_sc.cu -- warp-per-transform float kernels, scalar-arithmetic build (see fast.h)
#define PF_NO_PACKED_F32 1
#define PF_FAST_PART 0
#include "fast_impl.cuh"
