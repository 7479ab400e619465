// tcgen05 GEMM instantiations, part 0 of 3 (see the NOVA_GEMM_TU note in gemm_tcgen05.cuh).
#define NOVA_GEMM_TU 0
#include "gemm_tcgen05.cuh"
