// tcgen05 GEMM instantiations, part 3 of 4: the gate GEMM with the block tail in its epilogue (see gemm_tcgen05.cuh).
#define NOVA_GEMM_TU 3
#include "gemm_tcgen05.cuh"
