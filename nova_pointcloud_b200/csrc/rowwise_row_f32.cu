// Row-kernel instantiations, fp32 parity handle (see the note at the end of rowwise.cuh).
#include "rowwise.cuh"

namespace nova {
namespace rw {
int launch_row_f32(const RowParams& p, bool has_prev, int out, cudaStream_t stream) {
  return launch_row<float>(p, has_prev, out, stream);
}
}  // namespace rw
}  // namespace nova
