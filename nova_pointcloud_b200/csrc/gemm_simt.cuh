// SIMT GEMM: C[M,N] = epi(A[M,K] * W[N,K]^T + bias[N]), fp32 FMA accumulation.
//
// This is the arithmetic of the fp32 parity mode (<= 1e-5 vs the reference, which a
// tf32/bf16 tensor-core product cannot meet) and the isolation reference for the
// tcgen05 kernel in tests.  It is not the throughput path.
#pragma once

#include "common.cuh"

namespace nova {

// EPI_ADALN and EPI_TAIL exist on the tcgen05 kernel only (AdaLN statistics GEMM with the modulation fused in; gate
// GEMM with the block tail x += LN_aff(u) * gate fused in).
enum Epilogue : int { EPI_BIAS = 0, EPI_BIAS_SILU = 1, EPI_ADALN = 2, EPI_TAIL = 3,
                      EPI_BIAS_SILU_DUAL = 4 };  // tcgen05 only: also stores the pre-activation (training forward)

namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

template <typename TIn, typename TOut, int EPI, bool ACCURATE>
__global__ void __launch_bounds__(THREADS)
gemm_kernel(const TIn* __restrict__ A, int64_t lda, const TIn* __restrict__ W, int64_t ldw,
            const float* __restrict__ bias, TOut* __restrict__ C, int64_t ldc, int M, int N, int K) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row 0..63, k offset 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;       // compute: 4x4 micro tile

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float av[4], wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      av[i] = (m0 + lr < M && k < K) ? to_float(A[(int64_t)(m0 + lr) * lda + k]) : 0.f;
      wv[i] = (n0 + lr < N && k < K) ? to_float(W[(int64_t)(n0 + lr) * ldw + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = av[i];
      Ws[lk + i][lr] = wv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (EPI == EPI_BIAS_SILU) v = ACCURATE ? silu_accurate(v) : silu(v);
      C[(int64_t)m * ldc + n] = from_float<TOut>(v);
    }
  }
}

template <typename TIn, typename TOut, bool ACCURATE>
int launch(const TIn* A, int64_t lda, const TIn* W, int64_t ldw, const float* bias, TOut* C, int64_t ldc, int M,
           int N, int K, int epi, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return NOVA_OK;
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
  if (epi == EPI_BIAS)
    gemm_kernel<TIn, TOut, EPI_BIAS, ACCURATE><<<grid, THREADS, 0, stream>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  else
    gemm_kernel<TIn, TOut, EPI_BIAS_SILU, ACCURATE><<<grid, THREADS, 0, stream>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

}  // namespace simt
}  // namespace nova
