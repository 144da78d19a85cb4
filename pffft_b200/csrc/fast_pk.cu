// fast_pk.cu -- warp-per-transform float kernels, packed f32x2 build (see fast.h)
#define PF_FAST_PART 1
#include "fast_impl.cuh"
