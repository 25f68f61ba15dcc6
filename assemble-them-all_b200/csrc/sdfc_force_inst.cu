// One explicit instantiation of the force kernel per translation unit; the
// variant index comes from -DSDFC_FORCE_VARIANT=n (bit 2 = DERIV, bit 1 = ANYFRIC,
// bit 0 = ANYDAMP). Keeps the 8 heavy instantiations compiling in parallel.
#include "sdfc_force.cuh"

#ifndef SDFC_FORCE_VARIANT
#error "compile with -DSDFC_FORCE_VARIANT=0..7"
#endif

namespace sdfc {
template void launch_force_t<((SDFC_FORCE_VARIANT >> 2) & 1) != 0, ((SDFC_FORCE_VARIANT >> 1) & 1) != 0,
                             (SDFC_FORCE_VARIANT & 1) != 0>(const BatchDev&, const double*, const double*, cudaStream_t);
}  // namespace sdfc
