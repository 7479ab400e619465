// Row-kernel instantiations, bf16 wide (small-M) dataflow (see the note at the end of rowwise.cuh).
#include "rowwise.cuh"

namespace nova {
namespace rw {
int launch_row_bf16(const RowParams& p, bool has_prev, int out, cudaStream_t stream) {
  return launch_row<bf16>(p, has_prev, out, stream);
}
}  // namespace rw
}  // namespace nova
