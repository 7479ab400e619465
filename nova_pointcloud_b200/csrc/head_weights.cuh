// Read-only view of a loaded head handle's packed parameters, for the training-mode kernels (train_bwd.cu).
// Filled by head_weights_view (head.cu), where struct nova_head lives.
#pragma once

#include "common.cuh"

namespace nova {

constexpr int HW_MAX_DEPTH = 16;

struct HeadWeightsView {
  int D = 0, Dc = 0, T = 0, depth = 0, channels = 0, dtype = 0;
  bool use_simt_gemm = false;
  // GEMM operands in the handle's element type, nn.Linear layout (out, in)
  const void* w_c1 = nullptr;   // [D, Dc]
  const void* w_c2 = nullptr;   // [D, D]
  const void* w_ada = nullptr;  // [(3 depth + 2) D, D]: per block scale | shift | gate, then final scale | shift
  const void* w_fc1[HW_MAX_DEPTH] = {};
  const void* w_fc2[HW_MAX_DEPTH] = {};
  // fp32 parameters
  const float *b_c1 = nullptr, *b_c2 = nullptr, *b_ada = nullptr;
  const float *b_fc1[HW_MAX_DEPTH] = {}, *b_fc2[HW_MAX_DEPTH] = {}, *gamma[HW_MAX_DEPTH] = {}, *beta[HW_MAX_DEPTH] = {};
  const float *w_t1 = nullptr, *b_t1 = nullptr, *w_t2 = nullptr, *b_t2 = nullptr;  // [D, 256], [D, D]
  const float *w_patch = nullptr, *b_patch = nullptr;                              // [D, T] token order (p, p, C)
  const float *w_head = nullptr, *b_head = nullptr;                                // [T, D]
};

}  // namespace nova

struct nova_head;
int head_weights_view(const nova_head* h, nova::HeadWeightsView* out, const char* who);
