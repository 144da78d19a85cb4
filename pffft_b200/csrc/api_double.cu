// api_double.cu -- pffftd_* : the double-precision C-ABI (ref include/pffft/pffft_double.h:129-245).
// Complex cores of 512..4096 points run on the 16x16xC CTA kernels, every other size on the generic Stockham kernels.
#include "../../include/pffft/pffft_b200.h"
#include "api_impl.cuh"
#include "cta_hooks.cuh"

PF_API(pffftd_, pffftdb_, PFFFTD_Setup, double, pf::CtaOnlyHooks<double>, doubles_per_transform)
