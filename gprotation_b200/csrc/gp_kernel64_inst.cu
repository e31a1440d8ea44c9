// Explicit instantiation of the two-problems-per-SM kernel (gp_kernel64.cuh): -DGPROT_INST_EXACT=0|1
#include "gp_kernel64.cuh"

#if !defined(GPROT_INST_EXACT)
#error "define GPROT_INST_EXACT"
#endif

namespace gprot {
namespace k64 {
template __global__ void gp_lnlike64_kernel<(GPROT_INST_EXACT != 0)>(const Params);
}
}
