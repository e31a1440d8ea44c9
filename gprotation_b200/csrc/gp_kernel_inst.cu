// One explicit instantiation of gp_lnlike_kernel per translation unit, so that the eight variants
// (EXACT x CTAs per problem) compile in parallel: nvcc ... -DGPROT_INST_EXACT=0|1 -DGPROT_INST_CL=1|2|4|8 -c
#include "gp_kernel.cuh"

#if !defined(GPROT_INST_EXACT) || !defined(GPROT_INST_CL)
#error "define GPROT_INST_EXACT and GPROT_INST_CL"
#endif

namespace gprot {
template __global__ void gp_lnlike_kernel<(GPROT_INST_EXACT != 0), GPROT_INST_CL>(const Params);
}
