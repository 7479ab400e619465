// Host interface of the cluster chain kernel (chain_tcgen05.cu): one launch runs, for every 128-row block, the whole
// block chain of a diffusion step that follows the AdaLN statistics GEMM.
#pragma once

#include "common.cuh"

namespace nova {
namespace chain {

struct ChainParams {
  int64_t M;          // head rows
  int D, T, depth;
  bf16* x;            // [M, D] residual stream (scratch)
  bf16* h;            // [M, D] modulated activations (scratch; A operand of fc1)
  bf16* u1;           // [M, D] fc1 output (scratch; A operand of fc2)
  bf16* u2;           // [M, D] fc2 output (scratch)
  const bf16* st;     // [M, ldst] AdaLN statistics of this step (scale | shift | gate per block, then final scale | shift)
  int64_t ldst;
  // per block i, D floats each, contiguous: b_fc1 | b_fc2 | norm2 weight | norm2 bias  at (4 i + k) D
  const float* fc_params;
  const float* x_tok;  // [x_rows, T] latent of the selected tokens
  int64_t x_rows;
  const float* Wp;     // [D, T] token-order patch-embed weight
  const float* WpT;    // [T, D] its transpose
  const float* bp;     // [D]
  const float* Wh;     // [T, D]
  const float* bh;     // [T]
  float* v_out;        // [M, T] or nullptr
  float* xt_out;       // Euler-updated latent [M, T] or nullptr
  float dt;
  // diagnostic (nova_debug_chain_timeline): SM-clock stamps of cluster 0 / CTA 0, 8 slots per stage; nullptr = off
  long long* timeline;
};

bool supported(int D);
// rows up to which the chain kernel beats the separate launches: all ceil(M / 64) clusters resident at once
int64_t profitable_rows(int D);
// w_stack: the fc weights of all blocks as one K-major matrix [2 depth D, D]: fc1_0 | fc2_0 | fc1_1 | ...
int launch(const ChainParams& p, const bf16* w_stack, cudaStream_t stream);
constexpr int TIMELINE_SLOTS = 8 * 64;
long long* timeline_buffer(bool create);  // device buffer [TIMELINE_SLOTS], created on request (NOVA_B200_CHAIN_TIMELINE=1)

}  // namespace chain
}  // namespace nova
