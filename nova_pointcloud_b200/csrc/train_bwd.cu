// Training-mode head: forward with saved activations and the full backward pass, sm_100a.
//
// What the reference does with autograd -- Transformer3DModel.get_losses -> loss.backward()
// (diffnext/models/transformers/transformer_3d.py:79-100) through DiffusionMLP.forward with PER-TOKEN timesteps
// (diffnext/models/diffusion_mlp.py:56-99, normalization.py:24-36) -- is written out here by hand:
//
//   forward   f = [cos(t w), sin(t w)]                      t1p = f Wt1^T + bt1   t1 = silu(t1p)   temb = t1 Wt2^T + bt2
//             c1p = z Wc1^T + bc1   c1 = silu(c1p)          c = c1 Wc2^T + bc2    zt = c + temb    a = silu(zt)
//             st = a W_ada^T + b_ada  [M, (3L+2) D]         x_0 = x_tok We^T + be
//             block i:  h_i = LN0(x_i)(1 + scale_i) + shift_i        p1_i = h_i P1^T + b1     u1_i = silu(p1_i)
//                       u2_i = u1_i P2^T + b2                         x_{i+1} = (LN(u2_i) gamma + beta) gate_i + x_i
//             y = LN0(x_L)(1 + scale_f) + shift_f           v = y H^T + h0
//   backward  the chain rule over the same graph, in reverse.  Every product with a weight matrix is a GEMM on the
//             tensor cores (bf16 handle: the tcgen05 kernel of gemm_tcgen05.cuh; fp32 parity handle: the SIMT kernel):
//               dgrad   dX [M, K] = dY [M, N] W [N, K]          = gemm(dY, (W^T)[K, N])      W^T built once per call
//               wgrad   dW [N, K] = dY^T [N, M] X [M, K]        = gemm(dY^T, X^T), reduction over the M rows
//             A weight gradient has few output tiles (9 for D = 768) and a long reduction (M = 65 536), so the
//             M-reduction is split S ways: the transposes are written split-major ([S][N][M / S]) and ONE batched launch
//             (tc::launch_batched) computes the S partial products, which a small kernel sums in fp32.
//             LayerNorm / modulation / SiLU derivatives and the bias / gamma / beta column sums are HBM-bound row
//             kernels (one warp per row, fp32 arithmetic, deterministic two-stage column reductions -- no atomics).
//
// Gradients are WRITTEN (not accumulated) as fp32 in the reference's state_dict shapes; the residual-stream gradient
// is carried in fp32.  Nothing here is on the sampling path; the product never falls back to the CPU.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_api.cuh"
#include "head_weights.cuh"

using namespace nova;

namespace nova {
namespace tb {

constexpr int WARPS = 8, THREADS = 256;
constexpr float NEG_LOG_THETA_OVER_HALF = -9.210340371976184f / 128.0f;  // diffusion_mlp.py:67, 256-wide embedding

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dsilu(float p) {  // d/dp (p sigmoid(p))
  const float s = sigmoidf_acc(p);
  return s * (1.0f + p * (1.0f - s));
}

// ------------------------------------------------------------------ element-wise forward pieces
template <typename AT>
__global__ void freq_kernel(const float* __restrict__ t, int64_t M, AT* __restrict__ f) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 128) return;
  const int64_t r = i >> 7;
  const int k = static_cast<int>(i & 127);
  const float w = expf(static_cast<float>(k) * NEG_LOG_THETA_OVER_HALF);
  float s, c;
  sincosf(t[r] * w, &s, &c);
  f[r * 256 + k] = from_float<AT>(c);
  f[r * 256 + 128 + k] = from_float<AT>(s);
}

template <typename AT>
__global__ void embed_fwd_kernel(const float* __restrict__ xt, const float* __restrict__ Wp, const float* __restrict__ bp,
                                 AT* __restrict__ x0, int64_t M, int D, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const int64_t r = i / D;
  const int d = static_cast<int>(i - r * D);
  float acc = bp[d];
  for (int k = 0; k < T; ++k) acc = fmaf(xt[r * T + k], Wp[(int64_t)d * T + k], acc);
  x0[i] = from_float<AT>(acc);
}

// element-wise kernels: 8 elements per thread (16-byte accesses for bf16); n = M * D with D % 256 == 0
template <typename AT>
__global__ void silu_fwd_kernel(const AT* __restrict__ p, AT* __restrict__ u, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  float v[8];
  load8(p + i, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = silu_accurate(v[e]);
  store8(u + i, v);
}

template <typename AT>
__global__ void add_silu_kernel(const AT* __restrict__ c, const AT* __restrict__ temb, AT* __restrict__ zt,
                                AT* __restrict__ a, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  float x[8], y[8], z[8];
  load8(c + i, x);
  load8(temb + i, y);
#pragma unroll
  for (int e = 0; e < 8; ++e) z[e] = to_float(from_float<AT>(x[e] + y[e]));  // the stored (rounded) value is what the
  store8(zt + i, z);                                                         // backward differentiates
#pragma unroll
  for (int e = 0; e < 8; ++e) z[e] = silu_accurate(z[e]);
  store8(a + i, z);
}

// dp = du * silu'(p)   (in place allowed: dp == du)
template <typename AT>
__global__ void silu_bwd_kernel(const AT* __restrict__ du, const AT* __restrict__ p, AT* __restrict__ dp, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  float g[8], v[8];
  load8(du + i, g);
  load8(p + i, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) g[e] *= dsilu(v[e]);
  store8(dp + i, g);
}

template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = from_float<TD>(to_float(src[i]));
}

// ------------------------------------------------------------------ row kernels: one warp per row, D % 256 == 0
// h = LN0(x)(1 + scale) + shift  (normalization.py:34-36, eps 1e-6, no affine);  sx = (mean, rstd)
template <typename AT>
__global__ void __launch_bounds__(THREADS)
ln_mod_fwd_kernel(const AT* __restrict__ x, const AT* __restrict__ st, int64_t ldst, int64_t scale_off, AT* __restrict__ h,
                  float2* __restrict__ sx, int64_t M, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= M) return;
  const AT* xr = x + row * D;
  float s = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float v[8];
    load8(xr + d, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float v[8];
    load8(xr + d, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) q = fmaf(v[e] - mean, v[e] - mean, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + 1e-6f);
  if (lane == 0) sx[row] = make_float2(mean, rstd);
  const AT* sc = st + row * ldst + scale_off;
  const AT* sh = sc + D;
  for (int d = lane * 8; d < D; d += 256) {
    float v[8], a[8], b[8], o[8];
    load8(xr + d, v);
    load8(sc + d, a);
    load8(sh + d, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = fmaf((v[e] - mean) * rstd, 1.0f + a[e], b[e]);
    store8(h + row * D + d, o);
  }
}

// x' = (LN(u2) gamma + beta) gate + x   (diffusion_mlp.py:53, eps 1e-5);  su = (mean, rstd) of u2
template <typename AT>
__global__ void __launch_bounds__(THREADS)
tail_fwd_kernel(const AT* __restrict__ u2, const AT* __restrict__ x, const AT* __restrict__ st, int64_t ldst, int64_t gate_off,
                const float* __restrict__ gamma, const float* __restrict__ beta, AT* __restrict__ xo, float2* __restrict__ su,
                int64_t M, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= M) return;
  const AT* ur = u2 + row * D;
  float s = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float v[8];
    load8(ur + d, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float v[8];
    load8(ur + d, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) q = fmaf(v[e] - mean, v[e] - mean, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + 1e-5f);
  if (lane == 0) su[row] = make_float2(mean, rstd);
  const AT* gt = st + row * ldst + gate_off;
  for (int d = lane * 8; d < D; d += 256) {
    float v[8], g[8], xv[8], ga[8], be[8], o[8];
    load8(ur + d, v);
    load8(gt + d, g);
    load8(x + row * D + d, xv);
    load8(gamma + d, ga);
    load8(beta + d, be);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = fmaf(fmaf((v[e] - mean) * rstd, ga[e], be[e]), g[e], xv[e]);
    store8(xo + row * D + d, o);
  }
}

// v[r, k] = bh[k] + sum_d y[r, d] Wh[k, d]   (T small)
template <typename AT>
__global__ void __launch_bounds__(THREADS)
head_fwd_kernel(const AT* __restrict__ y, const float* __restrict__ Wh, const float* __restrict__ bh, float* __restrict__ v,
                int64_t M, int D, int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= M) return;
  for (int k = 0; k < T; ++k) {
    float acc = 0.f;
    for (int d = lane * 8; d < D; d += 256) {
      float a[8], w[8];
      load8(y + row * D + d, a);
      load8(Wh + (int64_t)k * D + d, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(a[e], w[e], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) v[row * T + k] = acc + bh[k];
  }
}

// dy[r, d] = sum_k dv[r, k] Wh[k, d]
template <typename AT>
__global__ void head_bwd_dy_kernel(const float* __restrict__ dv, const float* __restrict__ Wh, AT* __restrict__ dy, int64_t M,
                                   int D, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const int64_t r = i / D;
  const int d = static_cast<int>(i - r * D);
  float acc = 0.f;
  for (int k = 0; k < T; ++k) acc = fmaf(dv[r * T + k], Wh[(int64_t)k * D + d], acc);
  dy[i] = from_float<AT>(acc);
}

// Backward of h = xn (1 + scale) + shift, xn = (x - mean) rstd:
//   dscale = dh xn, dshift = dh  -> the scale / shift slots of dst (the gradient of the statistics GEMM's output)
//   g = dh (1 + scale);  dx (+)= rstd (g - mean(g) - xn mean(g xn))
template <typename AT>
__global__ void __launch_bounds__(THREADS)
ln_mod_bwd_kernel(const AT* __restrict__ dh, const AT* __restrict__ x, const float2* __restrict__ sx, const AT* __restrict__ st,
                  int64_t ldst, int64_t scale_off, AT* __restrict__ dst, float* __restrict__ dx, int accumulate, int64_t M,
                  int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= M) return;
  const float2 ms = sx[row];
  const AT* sc = st + row * ldst + scale_off;
  AT* dsc = dst + row * ldst + scale_off;
  AT* dsh = dsc + D;
  float s1 = 0.f, s2 = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float g[8], xv[8], a[8], o[8];
    load8(dh + row * D + d, g);
    load8(x + row * D + d, xv);
    load8(sc + d, a);
    store8(dsh + d, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xn = (xv[e] - ms.x) * ms.y;
      o[e] = g[e] * xn;
      const float ge = g[e] * (1.0f + a[e]);
      s1 += ge;
      s2 = fmaf(ge, xn, s2);
    }
    store8(dsc + d, o);
  }
  const float m1 = warp_sum(s1) / static_cast<float>(D), m2 = warp_sum(s2) / static_cast<float>(D);
  for (int d = lane * 8; d < D; d += 256) {
    float g[8], xv[8], a[8], o[8];
    load8(dh + row * D + d, g);
    load8(x + row * D + d, xv);
    load8(sc + d, a);
    if (accumulate) load8(dx + row * D + d, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xn = (xv[e] - ms.x) * ms.y;
      const float v = ms.y * (g[e] * (1.0f + a[e]) - m1 - xn * m2);
      o[e] = accumulate ? o[e] + v : v;
    }
    store8(dx + row * D + d, o);
  }
}

// Backward of x' = ln gate + x, ln = un gamma + beta, un = (u2 - mean) rstd:
//   dgate = dx' ln -> the gate slot of dst;  dln = dx' gate;  tmp_g = dln un, tmp_b = dln (their column sums are
//   dgamma / dbeta);  dun = dln gamma;  du2 = rstd (dun - mean(dun) - un mean(dun un)).   dx' passes to dx unchanged.
template <typename AT>
__global__ void __launch_bounds__(THREADS)
tail_bwd_kernel(const float* __restrict__ dxo, const AT* __restrict__ u2, const float2* __restrict__ su,
                const AT* __restrict__ st, int64_t ldst, int64_t gate_off, const float* __restrict__ gamma,
                const float* __restrict__ beta, AT* __restrict__ dst, AT* __restrict__ du2, AT* __restrict__ tmp_g,
                AT* __restrict__ tmp_b, int64_t M, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= M) return;
  const float2 ms = su[row];
  const AT* gt = st + row * ldst + gate_off;
  AT* dgt = dst + row * ldst + gate_off;
  float s1 = 0.f, s2 = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float dxv[8], uv[8], g[8], ga[8], be[8], o1[8], o2[8], o3[8];
    load8(dxo + row * D + d, dxv);
    load8(u2 + row * D + d, uv);
    load8(gt + d, g);
    load8(gamma + d, ga);
    load8(beta + d, be);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float un = (uv[e] - ms.x) * ms.y;
      o1[e] = dxv[e] * fmaf(un, ga[e], be[e]);  // dgate
      const float dln = dxv[e] * g[e];
      o2[e] = dln * un;
      o3[e] = dln;
      const float dun = dln * ga[e];
      s1 += dun;
      s2 = fmaf(dun, un, s2);
    }
    store8(dgt + d, o1);
    store8(tmp_g + row * D + d, o2);
    store8(tmp_b + row * D + d, o3);
  }
  const float m1 = warp_sum(s1) / static_cast<float>(D), m2 = warp_sum(s2) / static_cast<float>(D);
  for (int d = lane * 8; d < D; d += 256) {
    float dxv[8], uv[8], g[8], ga[8], o[8];
    load8(dxo + row * D + d, dxv);
    load8(u2 + row * D + d, uv);
    load8(gt + d, g);
    load8(gamma + d, ga);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float un = (uv[e] - ms.x) * ms.y;
      o[e] = ms.y * (dxv[e] * g[e] * ga[e] - m1 - un * m2);
    }
    store8(du2 + row * D + d, o);
  }
}

// ------------------------------------------------------------------ reductions over the M rows (deterministic)
// partial[chunk][n] = sum over the chunk's rows of Y[r, n].  A warp covers 256 consecutive columns of one row (16-byte
// loads), the 8 warps of a block take every 8th row of the chunk; N % 8 == 0 and 16-byte aligned rows (vector form),
// any N otherwise (scalar form, the [M, T] inputs).
template <typename TI>
__global__ void __launch_bounds__(THREADS)
colsum_partial_kernel(const TI* __restrict__ Y, int64_t ld, int64_t M, int N, int64_t rows_per_chunk, float* __restrict__ partial) {
  __shared__ float sh[WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 256 + lane * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < M ? r0 + rows_per_chunk : M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n0 + 8 <= N) {
    for (int64_t r = r0 + warp; r < r1; r += WARPS) {
      float v[8];
      load8(Y + r * ld + n0, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sh[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < WARPS; ++k) s += sh[k][threadIdx.x];
    partial[(int64_t)blockIdx.y * N + n] = s;
  }
}
template <typename TI>
__global__ void __launch_bounds__(THREADS)
colsum_partial_scalar_kernel(const TI* __restrict__ Y, int64_t ld, int64_t M, int N, int64_t rows_per_chunk, float* __restrict__ partial) {
  __shared__ float sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < M ? r0 + rows_per_chunk : M;
  float acc = 0.f;
  if (n < N)
    for (int64_t r = r0 + ty; r < r1; r += 8) acc += to_float(Y[r * ld + n]);
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sh[k][tx];
    partial[(int64_t)blockIdx.y * N + n] = s;
  }
}
// out[i] = sum_c partial[c][i]   (fixed order)
__global__ void reduce_chunks_kernel(const float* __restrict__ partial, int chunks, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[(int64_t)c * n + i];
  out[i] = s;
}
// partial[chunk][k][d] = sum over the chunk's rows of S[r, k] Wd[r, d]   (S skinny fp32 [M, T], Wd wide [M, D])
template <typename WT>
__global__ void __launch_bounds__(THREADS)
skinny_wgrad_kernel(const float* __restrict__ S, int T, const WT* __restrict__ Wd, int D, int64_t M, int64_t rows_per_chunk,
                    float* __restrict__ partial) {
  const int d = blockIdx.x * THREADS + threadIdx.x;
  const int k = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.z * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < M ? r0 + rows_per_chunk : M;
  if (d >= D) return;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc = fmaf(S[r * T + k], to_float(Wd[r * D + d]), acc);
  partial[((int64_t)blockIdx.z * T + k) * D + d] = acc;
}
// out[i] = sum_s partial[s][i], partials in the activation type (the batched weight-gradient GEMM's outputs)
template <typename AT>
__global__ void splitk_reduce_kernel(const AT* __restrict__ partial, int S, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += to_float(partial[(int64_t)k * n + i]);
  out[i] = s;
}
// out [S][N][Mc] <- Y [M, N] (row stride ld): out[s][n][mc] = Y[s Mc + mc][n], zero beyond row M.
// 64 x 64 tiles through shared memory; a thread moves two adjacent elements each way, so a warp reads 64 consecutive
// columns of a row and writes 64 consecutive rows of an output line (128 B each for bf16).  Mc % 64 == 0: a tile never
// straddles two splits.
template <typename T>
struct Pair {
  T a, b;
};
template <typename TI, typename AT>
__global__ void __launch_bounds__(256)
transpose_split_kernel(const TI* __restrict__ Y, int64_t ld, int64_t M, int N, AT* __restrict__ out, int64_t Mc, int S) {
  __shared__ float tile[64][65];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const bool pair_in = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(Y) & (2 * sizeof(TI) - 1)) == 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t r = r0 + ty + 8 * j;
    const int n = n0 + 2 * tx;
    float v0 = 0.f, v1 = 0.f;
    if (r < M) {
      if (pair_in && n + 1 < N) {
        const Pair<TI> pr = *reinterpret_cast<const Pair<TI>*>(Y + r * ld + n);
        v0 = to_float(pr.a);
        v1 = to_float(pr.b);
      } else {
        if (n < N) v0 = to_float(Y[r * ld + n]);
        if (n + 1 < N) v1 = to_float(Y[r * ld + n + 1]);
      }
    }
    tile[ty + 8 * j][2 * tx] = v0;
    tile[ty + 8 * j][2 * tx + 1] = v1;
  }
  __syncthreads();
  const int64_t Mp = Mc * S;
  const int64_t s = r0 / Mc, mc0 = r0 - s * Mc;  // the whole tile lies in split s
  if (r0 >= Mp) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + ty + 8 * j;
    if (n < N) {
      Pair<AT> pr;
      pr.a = from_float<AT>(tile[2 * tx][ty + 8 * j]);
      pr.b = from_float<AT>(tile[2 * tx + 1][ty + 8 * j]);
      *reinterpret_cast<Pair<AT>*>(out + (s * N + n) * Mc + mc0 + 2 * tx) = pr;
    }
  }
}
// patch-embed weight gradient: token order [T = (ph, pw, c)][D] -> the reference's Conv2d layout (D, C, p, p)
__global__ void emit_patch_grad_kernel(const float* __restrict__ g_td, float* __restrict__ out, int D, int C, int p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int T = C * p * p;
  if (i >= (int64_t)D * T) return;
  const int d = static_cast<int>(i / T), r = static_cast<int>(i % T);  // r indexes (c, ph, pw) of the output
  const int pw = r % p, ph = (r / p) % p, c = r / (p * p);
  const int k = (ph * p + pw) * C + c;
  out[i] = g_td[(int64_t)k * D + d];
}

// ------------------------------------------------------------------ host side
inline unsigned blocks_for(int64_t n, int per = 256) { return static_cast<unsigned>(ceil_div(n, per)); }
inline unsigned row_blocks(int64_t M) { return static_cast<unsigned>(ceil_div(M, WARPS)); }

template <typename AT>
int gemm_nt(bool simt, const AT* A, int64_t lda, const AT* W, int64_t ldw, const float* bias, AT* C, int64_t ldc, int64_t M,
            int N, int K, cudaStream_t s);
template <>
int gemm_nt<float>(bool, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                   int64_t M, int N, int K, cudaStream_t s) {
  return simt::launch<float, float, true>(A, lda, W, ldw, bias, C, ldc, (int)M, N, K, EPI_BIAS, s);
}
template <>
int gemm_nt<bf16>(bool simt_path, const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C,
                  int64_t ldc, int64_t M, int N, int K, cudaStream_t s) {
  if (simt_path) return simt::launch<bf16, bf16, false>(A, lda, W, ldw, bias, C, ldc, (int)M, N, K, EPI_BIAS, s);
  return tc::launch(A, lda, W, ldw, bias, C, ldc, (int)M, N, K, EPI_BIAS, s);
}

// pre = A W^T + bias (kept for the backward pass) and act = silu(pre).  bf16 handle on the tensor cores: ONE launch whose
// epilogue stores both (tc::launch_silu_dual); otherwise the GEMM followed by the element-wise kernel.
template <typename AT>
int gemm_nt_silu(bool simt_path, const AT* A, int64_t lda, const AT* W, int64_t ldw, const float* bias, AT* pre, AT* act,
                 int64_t ldc, int64_t M, int N, int K, cudaStream_t s);
template <>
int gemm_nt_silu<float>(bool simt_path, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* pre,
                        float* act, int64_t ldc, int64_t M, int N, int K, cudaStream_t s) {
  NOVA_PROPAGATE(gemm_nt<float>(simt_path, A, lda, W, ldw, bias, pre, ldc, M, N, K, s));
  silu_fwd_kernel<float><<<blocks_for(M * N / 8), 256, 0, s>>>(pre, act, M * N);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}
template <>
int gemm_nt_silu<bf16>(bool simt_path, const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* pre,
                       bf16* act, int64_t ldc, int64_t M, int N, int K, cudaStream_t s) {
  static const bool dual = [] { const char* e = std::getenv("NOVA_B200_TRAIN_DUAL_SILU"); return e == nullptr || e[0] != '0'; }();
  if (!simt_path && dual) return tc::launch_silu_dual(A, lda, W, ldw, bias, pre, ldc, act, ldc, (int)M, N, K, s);
  NOVA_PROPAGATE(gemm_nt<bf16>(simt_path, A, lda, W, ldw, bias, pre, ldc, M, N, K, s));
  silu_fwd_kernel<bf16><<<blocks_for(M * N / 8), 256, 0, s>>>(pre, act, M * N);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

struct Carve {
  uint8_t* base;
  size_t off = 0;
  explicit Carve(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(count * sizeof(T), 256);
    return r;
  }
};

constexpr int COL_CHUNK_ROWS = 512;  // rows per first-stage block of the column reductions

// Everything a training step keeps between forward and backward (the "saved tensors" of autograd), then the
// backward's own scratch.  One workspace, carved identically by both calls.
template <typename AT>
struct Plan {
  // saved by the forward
  AT *f, *t1p, *t1, *c1p, *c1, *zt, *a, *st, *y;
  AT *x[HW_MAX_DEPTH + 1], *h[HW_MAX_DEPTH], *p1[HW_MAX_DEPTH], *u1[HW_MAX_DEPTH], *u2[HW_MAX_DEPTH];
  float2 *sx[HW_MAX_DEPTH + 1], *su[HW_MAX_DEPTH];
  AT *wt1, *wt2;  // time-MLP weights in the activation type (the handle keeps them in fp32 for the sampling path)
  // backward scratch
  AT *s0, *s1;             // two [M, D] activation-gradient buffers; the forward borrows them for temb / c
  AT *tmp_g, *tmp_b, *dst; // [M, D], [M, D], [M, n_ada]
  float* dx;               // [M, D] residual-stream gradient
  AT *yt, *xt;             // split-major transposes [rows][Mp]
  AT* partial;             // batched weight-gradient outputs [S][N][K]
  float* colpart;          // first-stage column sums / skinny weight gradients
  float* small;            // fp32 staging for gradients that need a re-layout
  float* wstage;           // [n_ada, D] fp32: the AdaLN weight gradients of all layers before they are split up
  AT *fc1T[HW_MAX_DEPTH], *fc2T[HW_MAX_DEPTH], *adaT, *t2T, *c2T, *c1T;  // transposed weights for the dgrads
  int64_t Mp;  // padded row count of the transposes: Smax * 64-aligned chunk
  size_t bytes;
};

// NOVA_B200_WGRAD_TRANSPOSE=1 restores the K-major weight gradients (split-major transposed copies + launch_batched)
inline bool mn_major_wgrad() {
  static const bool on = [] {
    const char* e = std::getenv("NOVA_B200_WGRAD_TRANSPOSE");
    return e == nullptr || std::atoi(e) == 0;
  }();
  return on;
}

inline int pick_split(bool tensor_path, int64_t M, int n_rows, int k_out) {
  if (!tensor_path) return 1;
  const int64_t tiles = ceil_div(n_rows, 256) * ceil_div(k_out, 256);
  int64_t S = ceil_div(74, tiles);
  const int64_t smax = ceil_div(M, 512);  // at least 512 rows of reduction per split
  if (S > smax) S = smax;
  if (S > 16) S = 16;
  return S < 1 ? 1 : static_cast<int>(S);
}

template <typename AT>
Plan<AT> make_plan(const HeadWeightsView& w, void* base, int64_t M) {
  Plan<AT> p{};
  const size_t D = w.D, Dc = w.Dc, L = w.depth, n_ada = (3 * L + 2) * D, m = static_cast<size_t>(M > 0 ? M : 1);
  Carve cv(base);
  p.f = cv.take<AT>(m * 256);
  p.t1p = cv.take<AT>(m * D); p.t1 = cv.take<AT>(m * D); p.c1p = cv.take<AT>(m * D); p.c1 = cv.take<AT>(m * D);
  p.zt = cv.take<AT>(m * D); p.a = cv.take<AT>(m * D); p.st = cv.take<AT>(m * n_ada); p.y = cv.take<AT>(m * D);
  for (size_t i = 0; i <= L; ++i) { p.x[i] = cv.take<AT>(m * D); p.sx[i] = cv.take<float2>(m); }
  for (size_t i = 0; i < L; ++i) {
    p.h[i] = cv.take<AT>(m * D); p.p1[i] = cv.take<AT>(m * D); p.u1[i] = cv.take<AT>(m * D); p.u2[i] = cv.take<AT>(m * D);
    p.su[i] = cv.take<float2>(m);
  }
  p.wt1 = cv.take<AT>(D * 256); p.wt2 = cv.take<AT>(D * D);
  p.s0 = cv.take<AT>(m * D); p.s1 = cv.take<AT>(m * D);
  p.tmp_g = cv.take<AT>(m * D); p.tmp_b = cv.take<AT>(m * D); p.dst = cv.take<AT>(m * n_ada);
  p.dx = cv.take<float>(m * D);
  p.Mp = static_cast<int64_t>(align_up(m, 64) + 64 * 16);  // any split S <= 16 with 64-aligned chunks fits
  const size_t wide = n_ada, narrow = D > Dc ? (D > 256 ? D : 256) : (Dc > 256 ? Dc : 256);
  // split-major transposed copies: only the K-major weight-gradient path needs them (SIMT GEMMs, or
  // NOVA_B200_WGRAD_TRANSPOSE=1); the MN-major tensor-core path reads dY and X as they lie
  const bool need_t = !(std::is_same<AT, bf16>::value && !w.use_simt_gemm && mn_major_wgrad());
  p.yt = cv.take<AT>(need_t ? wide * p.Mp : 8);
  p.xt = cv.take<AT>(need_t ? narrow * p.Mp : 8);
  // S partial products of an [n_rows, k_out] gradient with S <= ceil(74 / tiles), tiles = (n_rows / 256) ceil(k_out / 256):
  // S n_rows k_out <= 74 * 256 * 256 + n_rows k_out for every shape
  p.partial = cv.take<AT>((size_t)75 * 65536 + n_ada * narrow);
  const size_t chunks = ceil_div(m, COL_CHUNK_ROWS);
  const size_t wideT = (size_t)w.T * D > n_ada ? (size_t)w.T * D : n_ada;
  p.colpart = cv.take<float>(chunks * wideT);
  p.small = cv.take<float>((size_t)w.T * D);
  p.wstage = cv.take<float>(n_ada * D);
  for (size_t i = 0; i < L; ++i) { p.fc1T[i] = cv.take<AT>(D * D); p.fc2T[i] = cv.take<AT>(D * D); }
  p.adaT = cv.take<AT>(D * n_ada); p.t2T = cv.take<AT>(D * D); p.c2T = cv.take<AT>(D * D); p.c1T = cv.take<AT>(Dc * D);
  p.bytes = cv.off;
  return p;
}

template <typename AT>
int colsum(const AT* Y, int64_t ld, int64_t M, int N, float* colpart, float* out, cudaStream_t s) {
  if (out == nullptr) return NOVA_OK;
  const int chunks = static_cast<int>(ceil_div(M, COL_CHUNK_ROWS));
  const bool vec = N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(Y) & 31) == 0;
  if (vec)
    colsum_partial_kernel<AT><<<dim3(blocks_for(N, 256), (unsigned)chunks), THREADS, 0, s>>>(Y, ld, M, N, COL_CHUNK_ROWS, colpart);
  else
    colsum_partial_scalar_kernel<AT><<<dim3(blocks_for(N, 32), (unsigned)chunks), THREADS, 0, s>>>(Y, ld, M, N, COL_CHUNK_ROWS, colpart);
  NOVA_CHECK_LAUNCH();
  reduce_chunks_kernel<<<blocks_for(N), 256, 0, s>>>(colpart, chunks, N, out);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

// out[k][d] = sum_r S[r, k] Wd[r, d]
template <typename WT>
int skinny_wgrad(const float* S, int T, const WT* Wd, int D, int64_t M, float* colpart, float* out, cudaStream_t s) {
  const int chunks = static_cast<int>(ceil_div(M, COL_CHUNK_ROWS));
  skinny_wgrad_kernel<WT><<<dim3(blocks_for(D), (unsigned)T, (unsigned)chunks), THREADS, 0, s>>>(S, T, Wd, D, M, COL_CHUNK_ROWS, colpart);
  NOVA_CHECK_LAUNCH();
  reduce_chunks_kernel<<<blocks_for((int64_t)T * D), 256, 0, s>>>(colpart, chunks, (int64_t)T * D, out);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

template <typename TI, typename AT>
int transpose_split(const TI* Y, int64_t ld, int64_t M, int N, AT* out, int64_t Mc, int S, cudaStream_t s) {
  transpose_split_kernel<TI, AT><<<dim3(blocks_for(Mc * S, 64), blocks_for(N, 64)), 256, 0, s>>>(Y, ld, M, N, out, Mc, S);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

// dW [n_rows, k_out] (fp32, written) = dY^T X over the M rows; dY [M, n_rows] (row stride ldy), X [M, k_out] (stride ldx).
// xt_ready: p.xt already holds X^T for split S_given (the caller reuses one transpose for two gradients).
template <typename AT>
int wgrad(const HeadWeightsView& w, const Plan<AT>& p, const AT* dY, int64_t ldy, int n_rows, const AT* X, int64_t ldx, int k_out,
          int64_t M, float* out, cudaStream_t s, bool yt_ready = false) {
  if (out == nullptr) return NOVA_OK;
  const bool tensor_path = std::is_same<AT, bf16>::value && !w.use_simt_gemm;
  const int S = pick_split(tensor_path, M, n_rows, k_out);
  const int64_t Mc = static_cast<int64_t>(align_up(static_cast<size_t>(ceil_div(M, S)), 64));
  if (tensor_path && n_rows % 256 == 0 && mn_major_wgrad()) {
    // MN-major operands: the tensor cores read dY [M, n_rows] and X [M, k_out] as they lie (TMA boxes of 64 columns x
    // 64 rows, the reduction running over the rows) -- no transposed copies at all
    NOVA_PROPAGATE(tc::launch_batched_mn(reinterpret_cast<const bf16*>(dY), ldy, reinterpret_cast<const bf16*>(X), ldx,
                                         reinterpret_cast<bf16*>(p.partial), k_out, (int)M, n_rows, k_out, (int)Mc, S, s));
    splitk_reduce_kernel<AT><<<blocks_for((int64_t)n_rows * k_out), 256, 0, s>>>(p.partial, S, (int64_t)n_rows * k_out, out);
    NOVA_CHECK_LAUNCH();
    return NOVA_OK;
  }
  if (!yt_ready) NOVA_PROPAGATE((transpose_split<AT, AT>(dY, ldy, M, n_rows, p.yt, Mc, S, s)));
  NOVA_PROPAGATE((transpose_split<AT, AT>(X, ldx, M, k_out, p.xt, Mc, S, s)));
  if (tensor_path && n_rows % 256 == 0) {
    NOVA_PROPAGATE(tc::launch_batched(reinterpret_cast<const bf16*>(p.yt), Mc, reinterpret_cast<const bf16*>(p.xt), Mc,
                                      reinterpret_cast<bf16*>(p.partial), k_out, S * n_rows, k_out, (int)Mc, n_rows, s));
  } else {
    for (int k = 0; k < S; ++k)
      NOVA_PROPAGATE(gemm_nt<AT>(true, p.yt + (size_t)k * n_rows * Mc, Mc, p.xt + (size_t)k * k_out * Mc, Mc, nullptr,
                                 p.partial + (size_t)k * n_rows * k_out, k_out, n_rows, k_out, (int)Mc, s));
  }
  splitk_reduce_kernel<AT><<<blocks_for((int64_t)n_rows * k_out), 256, 0, s>>>(p.partial, S, (int64_t)n_rows * k_out, out);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}
// the split wgrad() would choose, so a caller can share dY^T between two gradients of the same shape
template <typename AT>
int wgrad_split(const HeadWeightsView& w, int64_t M, int n_rows, int k_out) {
  return pick_split(std::is_same<AT, bf16>::value && !w.use_simt_gemm, M, n_rows, k_out);
}

template <typename AT>
int train_forward(const HeadWeightsView& w, const float* x_tok, const float* t, const AT* z, int64_t M, float* v_out,
                  void* ws, cudaStream_t s) {
  const int D = w.D, Dc = w.Dc, T = w.T, L = w.depth, n_ada = (3 * L + 2) * D;
  const bool simt_path = w.use_simt_gemm;
  Plan<AT> p = make_plan<AT>(w, ws, M);
  const int64_t MD = M * D;
  const AT *wt1, *wt2;
  if (std::is_same<AT, float>::value) {
    wt1 = reinterpret_cast<const AT*>(w.w_t1);
    wt2 = reinterpret_cast<const AT*>(w.w_t2);
  } else {
    convert_kernel<float, AT><<<blocks_for((int64_t)D * 256), 256, 0, s>>>(w.w_t1, p.wt1, (int64_t)D * 256);
    NOVA_CHECK_LAUNCH();
    convert_kernel<float, AT><<<blocks_for((int64_t)D * D), 256, 0, s>>>(w.w_t2, p.wt2, (int64_t)D * D);
    NOVA_CHECK_LAUNCH();
    wt1 = p.wt1; wt2 = p.wt2;
  }
  // time / condition embedding (diffusion_mlp.py:65-75)
  freq_kernel<AT><<<blocks_for(M * 128), 256, 0, s>>>(t, M, p.f);
  NOVA_CHECK_LAUNCH();
  NOVA_PROPAGATE(gemm_nt_silu<AT>(simt_path, p.f, 256, wt1, 256, w.b_t1, p.t1p, p.t1, D, M, D, 256, s));
  NOVA_PROPAGATE(gemm_nt<AT>(simt_path, p.t1, D, wt2, D, w.b_t2, p.s0, D, M, D, D, s));  // temb
  NOVA_PROPAGATE(gemm_nt_silu<AT>(simt_path, z, Dc, static_cast<const AT*>(w.w_c1), Dc, w.b_c1, p.c1p, p.c1, D, M, D, Dc, s));
  NOVA_PROPAGATE(gemm_nt<AT>(simt_path, p.c1, D, static_cast<const AT*>(w.w_c2), D, w.b_c2, p.s1, D, M, D, D, s));  // c
  add_silu_kernel<AT><<<blocks_for(MD / 8), 256, 0, s>>>(p.s1, p.s0, p.zt, p.a, MD);
  NOVA_CHECK_LAUNCH();
  // all AdaLN statistics in one GEMM (normalization.py:34)
  NOVA_PROPAGATE(gemm_nt<AT>(simt_path, p.a, D, static_cast<const AT*>(w.w_ada), D, w.b_ada, p.st, n_ada, M, n_ada, D, s));
  embed_fwd_kernel<AT><<<blocks_for(MD), 256, 0, s>>>(x_tok, w.w_patch, w.b_patch, p.x[0], M, D, T);
  NOVA_CHECK_LAUNCH();
  for (int i = 0; i < L; ++i) {
    ln_mod_fwd_kernel<AT><<<row_blocks(M), THREADS, 0, s>>>(p.x[i], p.st, n_ada, (int64_t)3 * i * D, p.h[i], p.sx[i], M, D);
    NOVA_CHECK_LAUNCH();
    NOVA_PROPAGATE(gemm_nt_silu<AT>(simt_path, p.h[i], D, static_cast<const AT*>(w.w_fc1[i]), D, w.b_fc1[i], p.p1[i], p.u1[i], D, M,
                                    D, D, s));
    NOVA_PROPAGATE(gemm_nt<AT>(simt_path, p.u1[i], D, static_cast<const AT*>(w.w_fc2[i]), D, w.b_fc2[i], p.u2[i], D, M, D, D, s));
    tail_fwd_kernel<AT><<<row_blocks(M), THREADS, 0, s>>>(p.u2[i], p.x[i], p.st, n_ada, (int64_t)3 * i * D + 2 * D, w.gamma[i],
                                                         w.beta[i], p.x[i + 1], p.su[i], M, D);
    NOVA_CHECK_LAUNCH();
  }
  ln_mod_fwd_kernel<AT><<<row_blocks(M), THREADS, 0, s>>>(p.x[L], p.st, n_ada, (int64_t)3 * L * D, p.y, p.sx[L], M, D);
  NOVA_CHECK_LAUNCH();
  head_fwd_kernel<AT><<<row_blocks(M), THREADS, 0, s>>>(p.y, w.w_head, w.b_head, v_out, M, D, T);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

struct GradTable {
  std::map<std::string, float*> m;
  float* get(const std::string& k) const {
    auto it = m.find(k);
    return it == m.end() ? nullptr : it->second;
  }
};

template <typename AT>
int train_backward(const HeadWeightsView& w, const float* dv, const float* x_tok, const AT* z, int64_t M, const GradTable& g,
                   AT* dz_out, void* ws, cudaStream_t s) {
  const int D = w.D, Dc = w.Dc, T = w.T, L = w.depth, n_ada = (3 * L + 2) * D;
  const bool simt_path = w.use_simt_gemm;
  Plan<AT> p = make_plan<AT>(w, ws, M);
  const int64_t MD = M * D;
  const AT *wt1 = std::is_same<AT, float>::value ? reinterpret_cast<const AT*>(w.w_t1) : p.wt1;
  const AT *wt2 = std::is_same<AT, float>::value ? reinterpret_cast<const AT*>(w.w_t2) : p.wt2;
  (void)wt1;
  // W^T of every weight a dgrad multiplies by ([in, out], K-major for the GEMM's second operand)
  for (int i = 0; i < L; ++i) {
    NOVA_PROPAGATE((transpose_split<AT, AT>(static_cast<const AT*>(w.w_fc1[i]), D, D, D, p.fc1T[i], D, 1, s)));
    NOVA_PROPAGATE((transpose_split<AT, AT>(static_cast<const AT*>(w.w_fc2[i]), D, D, D, p.fc2T[i], D, 1, s)));
  }
  NOVA_PROPAGATE((transpose_split<AT, AT>(static_cast<const AT*>(w.w_ada), D, n_ada, D, p.adaT, n_ada, 1, s)));
  NOVA_PROPAGATE((transpose_split<AT, AT>(wt2, D, D, D, p.t2T, D, 1, s)));
  NOVA_PROPAGATE((transpose_split<AT, AT>(static_cast<const AT*>(w.w_c2), D, D, D, p.c2T, D, 1, s)));
  NOVA_PROPAGATE((transpose_split<AT, AT>(static_cast<const AT*>(w.w_c1), Dc, D, Dc, p.c1T, D, 1, s)));

  // ---- head: v = y H^T + h0 (diffusion_mlp.py:98)
  if (float* o = g.get("head.weight")) NOVA_PROPAGATE(skinny_wgrad<AT>(dv, T, p.y, D, M, p.colpart, o, s));
  NOVA_PROPAGATE(colsum<float>(dv, T, M, T, p.colpart, g.get("head.bias"), s));
  AT* dh = p.s0;  // gradient w.r.t. a modulated activation (y, then h_i)
  head_bwd_dy_kernel<AT><<<blocks_for(MD), 256, 0, s>>>(dv, w.w_head, dh, M, D, T);
  NOVA_CHECK_LAUNCH();
  // ---- final AdaLN (diffusion_mlp.py:97): first writer of dx
  ln_mod_bwd_kernel<AT><<<row_blocks(M), THREADS, 0, s>>>(dh, p.x[L], p.sx[L], p.st, n_ada, (int64_t)3 * L * D, p.dst, p.dx, 0, M, D);
  NOVA_CHECK_LAUNCH();
  for (int i = L - 1; i >= 0; --i) {
    const std::string blk = "blocks." + std::to_string(i) + ".";
    AT* du2 = p.s1;
    tail_bwd_kernel<AT><<<row_blocks(M), THREADS, 0, s>>>(p.dx, p.u2[i], p.su[i], p.st, n_ada, (int64_t)3 * i * D + 2 * D,
                                                         w.gamma[i], w.beta[i], p.dst, du2, p.tmp_g, p.tmp_b, M, D);
    NOVA_CHECK_LAUNCH();
    NOVA_PROPAGATE(colsum<AT>(p.tmp_g, D, M, D, p.colpart, g.get(blk + "norm2.weight"), s));
    NOVA_PROPAGATE(colsum<AT>(p.tmp_b, D, M, D, p.colpart, g.get(blk + "norm2.bias"), s));
    // fc2: u2 = u1 P2^T + b2
    NOVA_PROPAGATE(colsum<AT>(du2, D, M, D, p.colpart, g.get(blk + "proj.fc2.bias"), s));
    NOVA_PROPAGATE(wgrad<AT>(w, p, du2, D, D, p.u1[i], D, D, M, g.get(blk + "proj.fc2.weight"), s));
    AT* du1 = p.s0;  // dh of the block above has been consumed
    NOVA_PROPAGATE(gemm_nt<AT>(simt_path, du2, D, p.fc2T[i], D, nullptr, du1, D, M, D, D, s));
    silu_bwd_kernel<AT><<<blocks_for(MD / 8), 256, 0, s>>>(du1, p.p1[i], du1, MD);  // dp1 in place
    NOVA_CHECK_LAUNCH();
    // fc1: p1 = h P1^T + b1
    NOVA_PROPAGATE(colsum<AT>(du1, D, M, D, p.colpart, g.get(blk + "proj.fc1.bias"), s));
    NOVA_PROPAGATE(wgrad<AT>(w, p, du1, D, D, p.h[i], D, D, M, g.get(blk + "proj.fc1.weight"), s));
    dh = p.s1;  // du2 has been consumed
    NOVA_PROPAGATE(gemm_nt<AT>(simt_path, du1, D, p.fc1T[i], D, nullptr, dh, D, M, D, D, s));
    ln_mod_bwd_kernel<AT><<<row_blocks(M), THREADS, 0, s>>>(dh, p.x[i], p.sx[i], p.st, n_ada, (int64_t)3 * i * D, p.dst, p.dx, 1, M, D);
    NOVA_CHECK_LAUNCH();
  }
  // ---- patch embed: x_0 = x_tok We^T + be (embeddings.py:160-166); dx is now the gradient w.r.t. x_0
  NOVA_PROPAGATE(colsum<float>(p.dx, D, M, D, p.colpart, g.get("patch_embed.proj.bias"), s));
  if (float* o = g.get("patch_embed.proj.weight")) {
    NOVA_PROPAGATE(skinny_wgrad<float>(x_tok, T, p.dx, D, M, p.colpart, p.small, s));
    const int C = w.channels, pp = static_cast<int>(std::lround(std::sqrt(static_cast<double>(T / C))));
    emit_patch_grad_kernel<<<blocks_for((int64_t)D * T), 256, 0, s>>>(p.small, o, D, C, pp);
    NOVA_CHECK_LAUNCH();
  }
  // ---- the statistics GEMM st = a W_ada^T + b_ada: its output gradient dst is complete now; per layer, because the
  // reference keeps one nn.Linear per AdaLayerNormZero (normalization.py:28-32)
  {
    // ONE product for all (3L + 2) D statistics rows (dst^T a: 60 x 3 tiles at D = 768), staged in fp32; the rows of
    // each layer then go to that layer's gradient tensor (the reference keeps one nn.Linear per AdaLayerNormZero,
    // normalization.py:28-32)
    bool any_w = false;
    for (int i = 0; i <= L; ++i) {
      const bool fin = i == L;
      const std::string key = fin ? "norm.proj." : "blocks." + std::to_string(i) + ".norm1.proj.";
      any_w = any_w || g.get(key + "weight") != nullptr;
      NOVA_PROPAGATE(colsum<AT>(p.dst + (int64_t)3 * i * D, n_ada, M, fin ? 2 * D : 3 * D, p.colpart, g.get(key + "bias"), s));
    }
    if (any_w) {
      NOVA_PROPAGATE(wgrad<AT>(w, p, p.dst, n_ada, n_ada, p.a, D, D, M, p.wstage, s));
      for (int i = 0; i <= L; ++i) {
        const bool fin = i == L;
        float* out = g.get((fin ? std::string("norm.proj.") : "blocks." + std::to_string(i) + ".norm1.proj.") + "weight");
        if (out == nullptr) continue;
        NOVA_CHECK_CUDA(cudaMemcpyAsync(out, p.wstage + (size_t)3 * i * D * D, (size_t)(fin ? 2 : 3) * D * D * sizeof(float),
                                        cudaMemcpyDeviceToDevice, s));
      }
    }
  }
  AT* da = p.s0;
  NOVA_PROPAGATE(gemm_nt<AT>(simt_path, p.dst, n_ada, p.adaT, n_ada, nullptr, da, D, M, D, n_ada, s));
  silu_bwd_kernel<AT><<<blocks_for(MD / 8), 256, 0, s>>>(da, p.zt, da, MD);  // dzt in place: zt = c + temb feeds both branches
  NOVA_CHECK_LAUNCH();
  AT* dzt = da;
  // ---- timestep_proj (diffusion_mlp.py:59-60,74): temb = t1 Wt2^T + bt2, t1 = silu(f Wt1^T + bt1)
  NOVA_PROPAGATE(colsum<AT>(dzt, D, M, D, p.colpart, g.get("time_cond_embed.timestep_proj.fc2.bias"), s));
  NOVA_PROPAGATE(colsum<AT>(dzt, D, M, D, p.colpart, g.get("time_cond_embed.condition_proj.fc2.bias"), s));
  NOVA_PROPAGATE(wgrad<AT>(w, p, dzt, D, D, p.t1, D, D, M, g.get("time_cond_embed.timestep_proj.fc2.weight"), s));
  // same dY and the same split: dzt^T is already in place if the gradient above was computed
  NOVA_PROPAGATE(wgrad<AT>(w, p, dzt, D, D, p.c1, D, D, M, g.get("time_cond_embed.condition_proj.fc2.weight"), s,
                           g.get("time_cond_embed.timestep_proj.fc2.weight") != nullptr));
  AT* d1 = p.s1;
  NOVA_PROPAGATE(gemm_nt<AT>(simt_path, dzt, D, p.t2T, D, nullptr, d1, D, M, D, D, s));
  silu_bwd_kernel<AT><<<blocks_for(MD / 8), 256, 0, s>>>(d1, p.t1p, d1, MD);
  NOVA_CHECK_LAUNCH();
  NOVA_PROPAGATE(colsum<AT>(d1, D, M, D, p.colpart, g.get("time_cond_embed.timestep_proj.fc1.bias"), s));
  NOVA_PROPAGATE(wgrad<AT>(w, p, d1, D, D, p.f, 256, 256, M, g.get("time_cond_embed.timestep_proj.fc1.weight"), s));
  // ---- condition_proj (diffusion_mlp.py:61,75): c = c1 Wc2^T + bc2, c1 = silu(z Wc1^T + bc1)
  NOVA_PROPAGATE(gemm_nt<AT>(simt_path, dzt, D, p.c2T, D, nullptr, d1, D, M, D, D, s));
  silu_bwd_kernel<AT><<<blocks_for(MD / 8), 256, 0, s>>>(d1, p.c1p, d1, MD);
  NOVA_CHECK_LAUNCH();
  NOVA_PROPAGATE(colsum<AT>(d1, D, M, D, p.colpart, g.get("time_cond_embed.condition_proj.fc1.bias"), s));
  NOVA_PROPAGATE(wgrad<AT>(w, p, d1, D, D, z, Dc, Dc, M, g.get("time_cond_embed.condition_proj.fc1.weight"), s));
  if (dz_out != nullptr) NOVA_PROPAGATE(gemm_nt<AT>(simt_path, d1, D, p.c1T, D, nullptr, dz_out, Dc, M, Dc, D, s));
  return NOVA_OK;
}

}  // namespace tb
}  // namespace nova

// ------------------------------------------------------------------ C ABI
extern "C" size_t nova_head_train_bytes(const nova_head_t* h, int64_t rows) {
  HeadWeightsView w;
  if (rows < 0 || head_weights_view(h, &w, "nova_head_train_bytes") != NOVA_OK) return 0;
  return w.dtype == NOVA_F32 ? tb::make_plan<float>(w, nullptr, rows).bytes : tb::make_plan<bf16>(w, nullptr, rows).bytes;
}

static int check_train(const HeadWeightsView& w, int64_t rows, const void* ws, size_t ws_bytes, size_t need, const char* who) {
  NOVA_REQUIRE(rows >= 0 && rows < (1ll << 31), "%s: bad row count %lld", who, (long long)rows);
  NOVA_REQUIRE(w.D % 256 == 0, "%s: the training kernels need width %% 256 == 0 (got %d)", who, w.D);
  NOVA_REQUIRE(w.T <= 64 && w.depth <= HW_MAX_DEPTH, "%s: token_dim %d / depth %d out of range", who, w.T, w.depth);
  NOVA_REQUIRE(rows == 0 || (ws != nullptr && ws_bytes >= need), "%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need);
  NOVA_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  return NOVA_OK;
}

extern "C" int nova_head_train_forward(const nova_head_t* h, const float* x_tok, const float* t, const void* z, int64_t rows,
                                       float* v_out, void* workspace, size_t workspace_bytes, void* stream) {
  HeadWeightsView w;
  NOVA_PROPAGATE(head_weights_view(h, &w, "nova_head_train_forward"));
  NOVA_PROPAGATE(check_train(w, rows, workspace, workspace_bytes, nova_head_train_bytes(h, rows), "nova_head_train_forward"));
  if (rows == 0) return NOVA_OK;
  NOVA_REQUIRE(x_tok && t && z && v_out, "nova_head_train_forward: null argument");
  pdl_set_for_rows(1ll << 40);  // ordinary launches: no programmatic dependent launch on the training path
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return w.dtype == NOVA_F32 ? tb::train_forward<float>(w, x_tok, t, static_cast<const float*>(z), rows, v_out, workspace, s)
                             : tb::train_forward<bf16>(w, x_tok, t, static_cast<const bf16*>(z), rows, v_out, workspace, s);
}

extern "C" int nova_head_backward(const nova_head_t* h, const float* dv, const float* x_tok, const void* z, int64_t rows,
                                  int32_t n_grads, const char* const* names, float* const* grads, void* dz_out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  HeadWeightsView w;
  NOVA_PROPAGATE(head_weights_view(h, &w, "nova_head_backward"));
  NOVA_PROPAGATE(check_train(w, rows, workspace, workspace_bytes, nova_head_train_bytes(h, rows), "nova_head_backward"));
  NOVA_REQUIRE(n_grads >= 0 && (n_grads == 0 || (names && grads)), "nova_head_backward: null gradient table");
  NOVA_REQUIRE(rows > 0, "nova_head_backward: no rows (the caller zero-fills the gradients of an empty batch)");
  NOVA_REQUIRE(dv && x_tok && z, "nova_head_backward: null argument");
  tb::GradTable g;
  for (int k = 0; k < n_grads; ++k) {
    NOVA_REQUIRE(names[k] != nullptr, "nova_head_backward: null gradient name");
    if (grads[k] != nullptr) g.m[std::string(names[k])] = grads[k];
  }
  pdl_set_for_rows(1ll << 40);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return w.dtype == NOVA_F32
             ? tb::train_backward<float>(w, dv, x_tok, static_cast<const float*>(z), rows, g, static_cast<float*>(dz_out), workspace, s)
             : tb::train_backward<bf16>(w, dv, x_tok, static_cast<const bf16*>(z), rows, g, static_cast<bf16*>(dz_out), workspace, s);
}
