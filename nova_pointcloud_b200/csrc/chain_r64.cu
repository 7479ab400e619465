// Chain-kernel instantiations, 64 rows per cluster (see chain_tcgen05.cuh).
#include "chain_tcgen05.cuh"

namespace nova {
namespace chain {
int launch_rows64(const ChainParams& p, const bf16* w_stack, cudaStream_t stream) { return launch_rows<64>(p, w_stack, stream); }
int max_clusters64(int D) { return max_clusters_rows<64>(D); }
}  // namespace chain
}  // namespace nova
