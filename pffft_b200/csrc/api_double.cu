// api_double.cu -- pffftd_* : the double-precision C-ABI (ref include/pffft/pffft_double.h:129-245).
// All sizes run on the generic Stockham kernels instantiated for double.
#include "../../include/pffft/pffft_b200.h"
#include "api_impl.cuh"

PF_API(pffftd_, pffftdb_, PFFFTD_Setup, double, pf::FastHooks<double>, doubles_per_transform)
