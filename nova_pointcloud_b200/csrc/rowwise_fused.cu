// Row-kernel instantiations of the fused-AdaLN dataflow: patch embed, block tail, velocity head + Euler
// (see the note at the end of rowwise.cuh).
#include "rowwise.cuh"

namespace nova {
namespace rw {
int embed_bf16(const float* x_tok, int64_t x_rows, const float* WpT, const float* bp, bf16* x_out, float* rowstats,
               int64_t M, int D, int T, cudaStream_t stream) {
  return dispatch_vpl<bf16, EmbedLauncher>(D, x_tok, x_rows, WpT, bp, x_out, rowstats, M, D, T, stream);
}
int resid_bf16(const bf16* u, const bf16* x_in, const bf16* gate, const float* gamma, const float* beta, bf16* x_out,
               float* rowstats, int64_t M, int D, int reverse, cudaStream_t stream) {
  return dispatch_vpl<bf16, ResidLauncher>(D, u, x_in, gate, gamma, beta, x_out, rowstats, M, D, reverse, stream);
}
int headout_bf16(const bf16* y, const float* Wh, const float* bh, float* v_out, const float* xt_in, float* xt_out, float dt,
                 int64_t M, int D, int T, cudaStream_t stream) {
  return dispatch_vpl<bf16, HeadoutLauncher>(D, y, Wh, bh, v_out, xt_in, xt_out, dt, M, D, T, stream);
}
int headout_cfg_bf16(const bf16* y, const float* Wh, const float* bh, float* x_sel, float dt, int64_t Mx, int D, int passes,
                     int mode, float scale, float scale3, cudaStream_t stream) {
  return dispatch_vpl<bf16, HeadoutCfgLauncher>(D, y, Wh, bh, x_sel, dt, Mx, D, passes, mode, scale, scale3, stream);
}
}  // namespace rw
}  // namespace nova
