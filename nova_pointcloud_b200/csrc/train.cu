// Training-mode arithmetic either side of the head, forward only, sm_100a:
//
//   nova_add_noise   x_t = sigma[i] * noise + (1 - sigma[i]) * x,  t = timesteps[i]  per token
//                    -- FlowMatchEulerDiscreteScheduler.add_noise, diffnext/schedulers/scheduling_cfm.py:106-117
//   nova_flow_loss   loss = sum_tok mean_T((pred - (noise - x))^2) * w / (sum(w) + 1e-5)
//                    -- Transformer3DModel.get_losses, diffnext/models/transformers/transformer_3d.py:91-95
//
// Element-wise / reduction work, HBM-bound and tiny next to the head (16 B per latent element).  The roundings
// follow the reference's op order (mul, mul, add -- no FMA contraction; mean over T then * w then / (sum w + 1e-5)),
// and both reductions are single-block trees in a fixed order, so results are deterministic run to run.
#include "common.cuh"

namespace nova {
namespace train {

constexpr int RED_THREADS = 1024;

__global__ void add_noise_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                 const float* __restrict__ sigma_table, const float* __restrict__ t_table,
                                 const int64_t* __restrict__ t_idx, int64_t tokens, int T, int n_train,
                                 float* __restrict__ x_t, float* __restrict__ t_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tokens * T) return;
  const int64_t tok = i / T;
  int64_t k = t_idx[tok];
  k = k < 0 ? 0 : (k >= n_train ? n_train - 1 : k);  // range is validated on the host side of the mirror
  const float s = sigma_table[k];
  x_t[i] = __fadd_rn(__fmul_rn(s, noise[i]), __fmul_rn(__fsub_rn(1.0f, s), x[i]));
  if (t_out != nullptr && i == tok * T) t_out[tok] = t_table[k];
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// one block: out[0] = sum(w) (or the token count when w == nullptr)
__global__ void __launch_bounds__(RED_THREADS) weight_sum_kernel(const float* __restrict__ w, int64_t tokens,
                                                                 float* __restrict__ out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < tokens; i += RED_THREADS) acc += w ? w[i] : 1.0f;
  const float r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = r;
}

__global__ void token_loss_kernel(const float* __restrict__ pred, const float* __restrict__ noise,
                                  const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ wsum, int64_t tokens, int T,
                                  float* __restrict__ loss_tok) {
  const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= tokens) return;
  float acc = 0.f;
  for (int c = 0; c < T; ++c) {
    const int64_t i = tok * T + c;
    const float d = __fsub_rn(pred[i], __fsub_rn(noise[i], x[i]));
    acc = __fadd_rn(acc, __fmul_rn(d, d));
  }
  const float m = acc / static_cast<float>(T);
  loss_tok[tok] = (m * (w ? w[tok] : 1.0f)) / (wsum[0] + 1e-5f);
}

__global__ void __launch_bounds__(RED_THREADS) total_kernel(const float* __restrict__ loss_tok, int64_t tokens,
                                                            float* __restrict__ out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < tokens; i += RED_THREADS) acc += loss_tok[i];
  const float r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = r;
}

}  // namespace train
}  // namespace nova

extern "C" int nova_add_noise(const float* x, const float* noise, const float* sigma_table, const float* t_table,
                              const int64_t* t_idx, int64_t tokens, int32_t T, int32_t n_train, float* x_t,
                              float* t_out, void* stream) {
  using namespace nova;
  NOVA_REQUIRE(x && noise && sigma_table && t_idx && x_t, "nova_add_noise: null pointer");
  NOVA_REQUIRE(t_out == nullptr || t_table != nullptr, "nova_add_noise: t_out needs t_table");
  NOVA_REQUIRE(tokens >= 0 && T > 0 && n_train > 0, "nova_add_noise: bad sizes tokens=%lld T=%d n_train=%d",
               (long long)tokens, (int)T, (int)n_train);
  if (tokens == 0) return NOVA_OK;
  const int64_t n = tokens * T;
  train::add_noise_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, noise, sigma_table, t_table, t_idx, tokens, T, n_train, x_t, t_out);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

extern "C" int nova_flow_loss(const float* pred, const float* noise, const float* x, const float* weight,
                              int64_t tokens, int32_t T, float* loss_tok, float* scratch2, void* stream) {
  using namespace nova;
  NOVA_REQUIRE(pred && noise && x && loss_tok && scratch2, "nova_flow_loss: null pointer");
  NOVA_REQUIRE(tokens > 0 && T > 0, "nova_flow_loss: bad sizes tokens=%lld T=%d", (long long)tokens, (int)T);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  train::weight_sum_kernel<<<1, train::RED_THREADS, 0, s>>>(weight, tokens, scratch2 + 1);
  NOVA_CHECK_LAUNCH();
  train::token_loss_kernel<<<(unsigned)ceil_div(tokens, 256), 256, 0, s>>>(pred, noise, x, weight, scratch2 + 1, tokens,
                                                                          T, loss_tok);
  NOVA_CHECK_LAUNCH();
  train::total_kernel<<<1, train::RED_THREADS, 0, s>>>(loss_tok, tokens, scratch2);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}
