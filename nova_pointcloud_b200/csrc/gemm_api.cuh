// Host-side interface of the tcgen05 GEMM (definitions: gemm.cu = gemm_tcgen05.cuh, helpers in runtime.cu).
#pragma once

#include <cuda.h>

#include <atomic>

#include "common.cuh"

namespace nova {
namespace tc {

constexpr int BM = 128;       // rows per CTA tile
constexpr int BK = 64;        // K elements per pipeline stage (one 128 B swizzle atom of bf16)
constexpr int BN_FULL = 256;  // tile columns at large M (and the only width EPI_ADALN supports)

// K-major bf16 matrix [rows, K] with row stride ld (elements) -> 2D tiled map, 128 B swizzle,
// box = {64 elements of K, box_rows}.  Out-of-bounds elements are zero-filled by TMA.
int make_tmap_kmajor(CUtensorMap* map, const bf16* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows);
// The same encoding serves the output: [M, N] row-major, box = {64 columns, 32 rows} per TMA store.
uint32_t* debug_word();  // host-mapped [4] words written on barrier timeout (device pointer)
extern uint32_t* g_debug_host;  // the same words, host pointer
int num_sms();  // SM count of the current device
// sets cudaFuncAttributeMaxDynamicSharedMemorySize once per (kernel, device); thread-safe
int ensure_smem_attr(const void* func, int bytes, std::atomic<unsigned long long>* done_mask);
int default_cta_group(int M);  // env NOVA_B200_CTA_GROUP=1|2 overrides the heuristic
int pdl_epi_mask();            // env NOVA_B200_PDL_EPI_MASK: bit e = GEMMs with epilogue e take part in programmatic dependent launch (default 7: every kind but the tail GEMM, see launch_epi)
bool fixed_column_grid();      // env NOVA_B200_FIXED_N=1 turns the fixed-column grids of launch_epi on
int tile_columns_override();   // env NOVA_B200_TILE_N=64|128|256 forces the tile columns of the plain GEMMs (tests)

struct AdaLNArgs {
  const bf16* x = nullptr;   // [M, ldx]
  int64_t ldx = 0;
  const float* rowstats = nullptr;  // [M, 2]
  bf16* gate = nullptr;      // [M, ldg], columns N - 2 * features
  int64_t ldg = 0;
  int features = 0;          // D: mod tiles cover 2 D weight rows
  // x came out of an EPI_TAIL epilogue: its LayerNorm statistics are `n_parts` partials [n_parts][M] of (mean_t, M2_t)
  // (one per 256-column tile) instead of rowstats
  const float2* parts = nullptr;
  int n_parts = 0;
};

// The block tail fused into the gate GEMM (diffusion_mlp.py:53):  x <- x + (LN(u; eps 1e-5) * gamma + beta) * g,
// g = a Wg^T + bg.  u's statistics arrive as partials from the fc2 epilogue, the new x's leave as partials.
struct TailArgs {
  const bf16* u = nullptr;    // [M, ldu] fc2 output
  int64_t ldu = 0;
  bf16* x = nullptr;          // [M, ldx] residual stream, updated in place
  int64_t ldx = 0;
  const float* gamma = nullptr;  // [N]
  const float* beta = nullptr;   // [N]
  const float2* u_parts = nullptr;  // [n_parts][M]: one per 256-column tile of fc2
  float2* x_parts = nullptr;        // [2 n_parts][M]: one per 128-column half tile (two epilogue warp sets)
  int n_parts = 0;                  // N / 256
};

// C[M,N] = epi(A[M,K] W[N,K]^T + bias), bf16 in / bf16 out, epi = EPI_BIAS | EPI_BIAS_SILU.
// cta_group: 0 = automatic (CTA pairs once there are at least 2 x 128 rows), 1, 2.
// reverse_m: walk the row blocks in descending order (see EpiParams::reverse_m).
// part_out (EPI_BIAS, N a multiple of 256): also emit [N / 256][M] partial LayerNorm statistics of the rounded outputs.
int launch(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc, int M, int N,
           int K, int epi, cudaStream_t stream, int cta_group = 0, bool reverse_m = false, float2* part_out = nullptr);

// pre = A W^T + bias and act = silu(pre), both [M, N] bf16, from one launch (the training forward keeps pre for the backward)
int launch_silu_dual(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* pre, int64_t ldpre,
                     bf16* act, int64_t ldact, int M, int N, int K, cudaStream_t stream);

// S = M / batch_m_rows independent GEMMs in one launch (the split-M weight gradients of train_bwd.cu):
//   C[b] [batch_m_rows, N] = A[b] [batch_m_rows, K] W[b] [N, K]^T, all three stacked along their rows.
int launch_batched(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, bf16* C, int64_t ldc, int M, int N, int K,
                   int batch_m_rows, cudaStream_t stream);

// The same products from MN-major operands (no transposed copies): C[b] [n_rows, N] = Y[b]^T X[b], Y = source
// [src_rows, n_rows], X = source [src_rows, N], batch b reducing over source rows [b k_len, (b + 1) k_len).
int launch_batched_mn(const bf16* Y, int64_t ldy, const bf16* X, int64_t ldx, bf16* C, int64_t ldc, int src_rows, int n_rows,
                      int N, int k_len, int batches, cudaStream_t stream);

// gate GEMM with the block tail in its epilogue (CTA pairs, 256-column tiles; M > 128, N a multiple of 256)
int launch_tail(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, const TailArgs& tail, int M, int N,
                int K, cudaStream_t stream, bool reverse_m = false);

// AdaLN statistics GEMM with the modulation fused into the epilogue:
//   W [2 features + gate_cols, K] packed per 128 features as [scale | shift], then gate rows;
//   h [M, features] = LN(x)(1 + scale) + shift,  gate [M, gate_cols] = a W_gate^T + b_gate.
int launch_adaln(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* h_out, int64_t ldh,
                 const AdaLNArgs& ada, int M, int N, int K, cudaStream_t stream, int cta_group = 0,
                 bool reverse_m = false);

}  // namespace tc
}  // namespace nova
