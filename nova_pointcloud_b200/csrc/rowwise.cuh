// Row-wise (HBM-bound) kernels of the diffusion head: everything between the GEMMs.
// One warp owns one token row; a row of D = VPL*256 values lives in registers
// (VPL 16-byte vectors per lane), so each activation is read once and written once.
//
// Reference arithmetic (file:line into /root/reference):
//   time embedding        diffnext/models/diffusion_mlp.py:65-75
//   AdaLN-zero modulate   diffnext/models/normalization.py:34-36   (LN eps 1e-6, no affine)
//   block tail            diffnext/models/diffusion_mlp.py:53      (LN eps 1e-5 affine, * gate + x)
//   head + Euler          diffnext/models/diffusion_mlp.py:98, diffnext/schedulers/scheduling_cfm.py:136
#pragma once

#include "common.cuh"

namespace nova {
namespace rw {

constexpr int WARPS = 8;           // rows per CTA
constexpr int THREADS = WARPS * 32;
constexpr int MAX_T = 64;
constexpr int MAX_STEPS = 256;
struct StepList {
  float v[MAX_STEPS];
};

// ------------------------------------------------------------------ time embedding
// hidden[r, d] = silu(b1[d] + sum_k emb(t_r)[k] * W1[d, k]),  emb = [cos(t f), sin(t f)], 256 wide
static __global__ void __launch_bounds__(THREADS)
temb_fc1_kernel(const float* __restrict__ t, int R, const float* __restrict__ W1, const float* __restrict__ b1,
                float* __restrict__ hidden, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.y;
  const int d = blockIdx.x * WARPS + warp;
  if (d >= D) return;
  const float tv = t[r];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = lane + 32 * i;  // 0..127
    const float f = expf(static_cast<float>(k) * (-9.210340371976184f / 128.0f));
    float s, c;
    sincosf(tv * f, &s, &c);
    acc = fmaf(c, W1[(int64_t)d * 256 + k], acc);
    acc = fmaf(s, W1[(int64_t)d * 256 + 128 + k], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) hidden[(int64_t)r * D + d] = silu_accurate(acc + b1[d]);
}
// temb[r, d] = b2[d] + sum_k hidden[r, k] * W2[d, k]
static __global__ void __launch_bounds__(THREADS)
temb_fc2_kernel(const float* __restrict__ hidden, int R, const float* __restrict__ W2,
                const float* __restrict__ b2, float* __restrict__ temb, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.y;
  const int d = blockIdx.x * WARPS + warp;
  if (d >= D) return;
  float acc = 0.f;
  for (int k = lane * 4; k < D; k += 128) {
    const float4 h = *reinterpret_cast<const float4*>(hidden + (int64_t)r * D + k);
    const float4 w = *reinterpret_cast<const float4*>(W2 + (int64_t)d * D + k);
    acc = fmaf(h.x, w.x, acc);
    acc = fmaf(h.y, w.y, acc);
    acc = fmaf(h.z, w.z, acc);
    acc = fmaf(h.w, w.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) temb[(int64_t)r * D + d] = acc + b2[d];
}

// ------------------------------------------------------------------ gather rows by pred_ids
// The ids of cloud b, token slot j: ptr[(b % batch) * ld + j].  A plain (B, n) pred_ids tensor is {ptr, n, B}; the
// on-device set scheduler passes a window of the (Bx, N) generation order: {order + first, N, Bx}, which also serves
// the guidance passes (clouds Bx.. repeat the ids of clouds 0..Bx-1).
struct IdsView {
  const int64_t* ptr;
  int64_t ld, batch;
  __device__ __forceinline__ int64_t at(int64_t b, int64_t j, int64_t fallback) const {
    return ptr ? ptr[(b % batch) * ld + j] : fallback;
  }
};

// dst[b*n + j, :] = src[b*N + ids[b*n + j], :]   (W elements per row, 8 per thread)
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, const IdsView ids, T* __restrict__ dst,
                                   int64_t B, int64_t N, int64_t n, int W, uint32_t* bad_ids) {
  const int64_t vec_per_row = W / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n * vec_per_row) return;
  const int64_t row = i / vec_per_row, v = i % vec_per_row;
  const int64_t b = row / n;
  int64_t tok = ids.at(b, row - b * n, row % n);
  if (tok < 0 || tok >= N) {  // out-of-range id: never index with it (the reference's gather raises); flag and read token 0
    if (bad_ids && v == 0) *bad_ids = 0xBAD1D5u;
    tok = 0;
  }
  float tmp[8];
  load8(src + (b * N + tok) * W + v * 8, tmp);
  store8(dst + row * W + v * 8, tmp);
}
// small fp32 rows (token latent, T values): dst[b*n+j, :] = src[(b % Bx)*N + ids[b*n+j], :]
static __global__ void
gather_tok_kernel(const float* __restrict__ src, const IdsView ids,
                                  float* __restrict__ dst, int64_t B, int64_t Bx, int64_t N, int64_t n, int T,
                                  uint32_t* bad_ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n * T) return;
  const int64_t row = i / T, c = i % T;
  const int64_t b = row / n;
  int64_t tok = ids.at(b, row - b * n, row % n);
  if (tok < 0 || tok >= N) {
    if (bad_ids && c == 0) *bad_ids = 0xBAD1D5u;
    tok = 0;
  }
  dst[i] = src[((b % Bx) * N + tok) * T + c];
}

// ------------------------------------------------------------------ step prologue
// a[m, :] = silu(c[m, :] + temb[tsel(m), :])   -> A operand of the AdaLN GEMM;  one warp per row
// tsel(m) = m / rows_per_t + t_offset
template <typename AT, bool ACCURATE>
__global__ void __launch_bounds__(THREADS)
prep_kernel(const AT* __restrict__ c, const float* __restrict__ temb, int64_t rows_per_t, int64_t t_offset,
            AT* __restrict__ a, int64_t M, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * WARPS + warp;
  pdl_trigger();
  if (m >= M) return;
  pdl_wait();
  const float* te = temb + ((m / rows_per_t) + t_offset) * D;
  const AT* cr = c + m * D;
  AT* ar = a + m * D;
  for (int e = lane * 8; e < D; e += 256) {
    float cv[8], tv[8], av[8];
    load8(cr + e, cv);
    load8(te + e, tv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = cv[j] + tv[j];
      av[j] = ACCURATE ? silu_accurate(s) : silu(s);
    }
    store8(ar + e, av);
  }
}

// ------------------------------------------------------------------ the fused row kernel
struct RowParams {
  int64_t M;
  int D, T;
  const void* x_in;    // [M, D]   residual stream (HAS_PREV)
  void* x_out;         // [M, D]   residual stream out
  // !HAS_PREV: the residual stream starts as PatchEmbed(x_tok): x = bp + Wp x_tok[m % x_rows]
  const float* x_tok;  // [x_rows, T]
  int64_t x_rows;
  const float* Wp;     // [D, T] token-order patch-embed weight
  const float* WpT;    // [T, D] its transpose: vector loads (same fmaf chain per feature: bias, then t = 0, 1, ..)
  const float* bp;     // [D]
  const void* x_emb = nullptr;  // !HAS_PREV, row kernels only: [M, D] rows that are ALREADY embedded (PatchEmbed.forward passes
                                // a 3-D input through, embeddings.py:165); nullptr: embed x_tok
  const void* u;       // [M, D]   fc2 output of the finished block (HAS_PREV)
  const void* st;      // [M, ldst] all AdaLN statistics of this step
  int64_t ldst;
  int64_t gate_off;    // column of the finished block's gate
  int64_t scale_off;   // column of the next modulation's scale; shift follows at +D
  const float* gamma;  // norm2 weight / bias of the finished block
  const float* beta;
  void* h_out;         // [M, D]   modulated activations (OUT == 0)
  // OUT == 1: head + Euler
  const float* Wh;     // [T, D]
  const float* bh;     // [T]
  float* v_out;        // [M, T] or nullptr
  const float* xt_in;  // [M, T] latent (Euler) or nullptr
  float* xt_out;       // [M, T]
  float dt;
};

template <int VPL>
__device__ __forceinline__ void row_stats(const float (&v)[VPL][8], float inv_d, float eps, float& mean,
                                          float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      q = fmaf(d, d, q);
    }
  rstd = rsqrtf(warp_sum(q) * inv_d + eps);
}

// HAS_PREV: x <- LN_affine(u; 1e-5) * gate + x        (tail of the block that just ran)
// then      y  = LN(x; 1e-6) * (1 + scale) + shift     (AdaLN of the next consumer)
// OUT == 0: h_out <- y.   OUT == 1: v = Wh y + bh; optional Euler update of the latent.
// COHERENT (the cluster chain kernel, chain_tcgen05.cu): u and x were written earlier in the SAME kernel, u by other
// CTAs of the cluster, so they are read with L2-coherent loads (ld.global.cg) instead of through L1.
__device__ __forceinline__ void load8_cg(const bf16* p, float (&v)[8]) {
  uint4 raw;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p));
  unpack8(raw, v);
}
__device__ __forceinline__ void load8_cg(const float* p, float (&v)[8]) { load8(p, v); }

template <typename AT, int VPL, bool HAS_PREV, int OUT, bool COHERENT = false, bool EMBEDDED = false>
__device__ __forceinline__ void row_body(const RowParams& p, const int64_t row, const int lane) {
  const int D = p.D;
  const float inv_d = 1.0f / static_cast<float>(D);
  const AT* st = static_cast<const AT*>(p.st) + row * p.ldst;

  // bf16 handle (the latency-bound small-M dataflow): this row's gate / scale / shift slices of the statistics are
  // requested now, kept packed (4 registers per 8 values), so that their L2 round trip overlaps the two LayerNorm
  // reductions instead of following them (the loads below sit behind stores the compiler must assume may alias).
  constexpr bool PREFETCH = sizeof(AT) == 2;
  uint4 pg[PREFETCH && HAS_PREV ? VPL : 1], psc[PREFETCH ? VPL : 1], psh[PREFETCH ? VPL : 1];
  if (PREFETCH) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int e = (i * 32 + lane) * 8;
      if (HAS_PREV) pg[i] = *reinterpret_cast<const uint4*>(st + p.gate_off + e);
      psc[i] = *reinterpret_cast<const uint4*>(st + p.scale_off + e);
      psh[i] = *reinterpret_cast<const uint4*>(st + p.scale_off + D + e);
    }
  }

  // same for the fp32 LayerNorm affine of the block tail while the registers allow it (D <= 1024)
  constexpr bool PREFETCH_AFFINE = PREFETCH && HAS_PREV && VPL <= 4;
  float pga[PREFETCH_AFFINE ? VPL : 1][8], pbe[PREFETCH_AFFINE ? VPL : 1][8];
  if (PREFETCH_AFFINE) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      load8(p.gamma + (i * 32 + lane) * 8, pga[i]);
      load8(p.beta + (i * 32 + lane) * 8, pbe[i]);
    }
  }

  float x[VPL][8];
  if (HAS_PREV) {
    const AT* xin = static_cast<const AT*>(p.x_in) + row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (COHERENT) load8_cg(xin + (i * 32 + lane) * 8, x[i]); else load8(xin + (i * 32 + lane) * 8, x[i]);
    }
  } else if (EMBEDDED) {
    // the caller's rows are the embedded tokens already; the residual stream starts as a copy of them
    const AT* xe = static_cast<const AT*>(p.x_emb) + row * D;
    AT* xout = static_cast<AT*>(p.x_out) + row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      load8(xe + (i * 32 + lane) * 8, x[i]);
      if (OUT == 0) store8(xout + (i * 32 + lane) * 8, x[i]);
    }
  } else {
    // PatchEmbed with K = T (embeddings.py:146,165): too skinny for a GEMM, fused here
    const float* xt = p.x_tok + (row % p.x_rows) * p.T;
    AT* xout = static_cast<AT*>(p.x_out) + row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) load8(p.bp + (i * 32 + lane) * 8, x[i]);
    for (int t = 0; t < p.T; ++t) {
      const float xv = xt[t];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float w[8];
        load8(p.WpT + (int64_t)t * D + (i * 32 + lane) * 8, w);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[i][j] = fmaf(xv, w[j], x[i][j]);
      }
    }
    if (OUT == 0) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int e = (i * 32 + lane) * 8;
        store8(xout + e, x[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[i][j] = to_float(from_float<AT>(x[i][j]));  // continue from the stored (rounded) residual
      }
    }
  }

  if (HAS_PREV) {
    const AT* uin = static_cast<const AT*>(p.u) + row * D;
    float u[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (COHERENT) load8_cg(uin + (i * 32 + lane) * 8, u[i]); else load8(uin + (i * 32 + lane) * 8, u[i]);
    }
    float mean, rstd;
    row_stats<VPL>(u, inv_d, 1e-5f, mean, rstd);
    AT* xout = static_cast<AT*>(p.x_out) + row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int e = (i * 32 + lane) * 8;
      float g[8], ga[8], be[8];
      if (PREFETCH) unpack8(pg[i], g); else load8(st + p.gate_off + e, g);
      if (!PREFETCH_AFFINE) {
        load8(p.gamma + e, ga);
        load8(p.beta + e, be);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ln = PREFETCH_AFFINE ? fmaf((u[i][j] - mean) * rstd, pga[i][j], pbe[i][j])
                                         : fmaf((u[i][j] - mean) * rstd, ga[j], be[j]);
        x[i][j] = fmaf(ln, g[j], x[i][j]);
      }
      if (OUT == 0) store8(xout + e, x[i]);
    }
  }

  float mean, rstd;
  row_stats<VPL>(x, inv_d, 1e-6f, mean, rstd);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int e = (i * 32 + lane) * 8;
    float sc[8], sh[8];
    if (PREFETCH) {
      unpack8(psc[i], sc);
      unpack8(psh[i], sh);
    } else {
      load8(st + p.scale_off + e, sc);
      load8(st + p.scale_off + D + e, sh);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) x[i][j] = fmaf((x[i][j] - mean) * rstd, 1.0f + sc[j], sh[j]);
    if (OUT == 0) store8(static_cast<AT*>(p.h_out) + row * D + e, x[i]);
  }

  if (OUT == 1) {
    // velocity head: outputs in groups of 3 (T = 3 for xyz tokens) whose weight loads and warp reductions overlap;
    // every output keeps its own fmaf chain and butterfly, so the values do not depend on the grouping
    for (int t0 = 0; t0 < p.T; t0 += 3) {
      float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (t0 + g < p.T) {
          const float* w = p.Wh + (int64_t)(t0 + g) * D;
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            float wv[8];
            load8(w + (i * 32 + lane) * 8, wv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[g] = fmaf(x[i][j], wv[j], acc[g]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < 3; ++g) acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], o);
      }
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int t = t0 + g;
        if (lane == g && t < p.T) {
          const float v = acc[g] + p.bh[t];
          if (p.v_out) p.v_out[row * p.T + t] = v;
          if (p.xt_out) p.xt_out[row * p.T + t] = __fadd_rn(__fmul_rn(v, p.dt), p.xt_in[row * p.T + t]);
        }
      }
    }
  }
}

template <typename AT, int VPL, bool HAS_PREV, int OUT>
__global__ void __launch_bounds__(THREADS)
row_kernel(const RowParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= p.M) return;
  if (!HAS_PREV && p.x_emb != nullptr) row_body<AT, VPL, HAS_PREV, OUT, false, !HAS_PREV>(p, row, lane);
  else row_body<AT, VPL, HAS_PREV, OUT>(p, row, lane);
}

template <typename AT, int VPL>
int launch_row_vpl(const RowParams& p, bool has_prev, int out, cudaStream_t stream) {
  const unsigned grid = (unsigned)ceil_div(p.M, WARPS);
  if (has_prev && out == 0) row_kernel<AT, VPL, true, 0><<<grid, THREADS, 0, stream>>>(p);
  else if (has_prev && out == 1) row_kernel<AT, VPL, true, 1><<<grid, THREADS, 0, stream>>>(p);
  else if (!has_prev && out == 0) row_kernel<AT, VPL, false, 0><<<grid, THREADS, 0, stream>>>(p);
  else row_kernel<AT, VPL, false, 1><<<grid, THREADS, 0, stream>>>(p);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

template <typename AT>
int launch_row(const RowParams& p, bool has_prev, int out, cudaStream_t stream) {
  if (p.M <= 0) return NOVA_OK;
  switch (p.D / 256) {
    case 1: return launch_row_vpl<AT, 1>(p, has_prev, out, stream);
    case 2: return launch_row_vpl<AT, 2>(p, has_prev, out, stream);
    case 3: return launch_row_vpl<AT, 3>(p, has_prev, out, stream);
    case 4: return launch_row_vpl<AT, 4>(p, has_prev, out, stream);
    case 5: return launch_row_vpl<AT, 5>(p, has_prev, out, stream);
    case 6: return launch_row_vpl<AT, 6>(p, has_prev, out, stream);
    case 7: return launch_row_vpl<AT, 7>(p, has_prev, out, stream);
    case 8: return launch_row_vpl<AT, 8>(p, has_prev, out, stream);
    default: break;
  }
  set_error("unsupported head width %d (multiple of 256, <= 2048)", p.D);
  return NOVA_ERR_INVALID;
}

// ------------------------------------------------------------------ row kernels of the fused-AdaLN path
// With the modulation fused into the AdaLN GEMM epilogue (gemm_tcgen05.cuh, EPI_ADALN) the row-wise work
// shrinks to: start the residual stream, finish a block (LN * gate + x), and the velocity head.  Each also
// leaves (mean, rstd) of the stored residual row for the next AdaLN epilogue.

// x0 = bp + Wp x_tok  (PatchEmbed with K = T; WpT is [T, D]);  rowstats = LN statistics of the stored row
template <typename AT, int VPL>
__global__ void __launch_bounds__(THREADS)
embed_kernel(const float* __restrict__ x_tok, int64_t x_rows, const float* __restrict__ WpT,
             const float* __restrict__ bp, AT* __restrict__ x_out, float* __restrict__ rowstats, int64_t M, int D,
             int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  pdl_trigger();
  if (row >= M) return;
  pdl_wait();
  const float* xt = x_tok + (row % x_rows) * T;
  float x[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) load8(bp + (i * 32 + lane) * 8, x[i]);
  for (int t = 0; t < T; ++t) {
    const float xv = xt[t];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float w[8];
      load8(WpT + (int64_t)t * D + (i * 32 + lane) * 8, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[i][j] = fmaf(xv, w[j], x[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[i][j] = to_float(from_float<AT>(x[i][j]));  // statistics of what is stored
    store8(x_out + row * D + (i * 32 + lane) * 8, x[i]);
  }
  float mean, rstd;
  row_stats<VPL>(x, 1.0f / static_cast<float>(D), 1e-6f, mean, rstd);
  if (lane == 0) *reinterpret_cast<float2*>(rowstats + 2 * row) = make_float2(mean, rstd);
}

// x <- LN_affine(u; 1e-5) * gate + x   (diffusion_mlp.py:53);  rowstats = LN statistics (1e-6) of the new row
//
// bf16 only (the fused-AdaLN dataflow).  The row stays PACKED in registers (VPL x uint4 each for u and x),
// unpacked on the fly for every pass, so the kernel needs <= 64 registers per thread and 128-thread CTAs:
// several of them fit next to a persistent tcgen05 GEMM CTA (which owns ~230 KB of shared memory but only
// 28-49 K registers), so on a second stream this HBM-bound work overlaps the other half's MMAs.
constexpr int RESID_WARPS = 4;
constexpr int RESID_THREADS = RESID_WARPS * 32;

__device__ __forceinline__ float bf16x2_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16x2_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  v[0] = bf16x2_lo(q.x); v[1] = bf16x2_hi(q.x); v[2] = bf16x2_lo(q.y); v[3] = bf16x2_hi(q.y);
  v[4] = bf16x2_lo(q.z); v[5] = bf16x2_hi(q.z); v[6] = bf16x2_lo(q.w); v[7] = bf16x2_hi(q.w);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <int VPL>
__device__ __forceinline__ void packed_stats(const uint4 (&q)[VPL], float inv_d, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float v[8];
    unpack8(q[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  mean = warp_sum(s) * inv_d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float v[8];
    unpack8(q[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[j] - mean;
      ss = fmaf(d, d, ss);
    }
  }
  rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
}

template <typename AT, int VPL>
__global__ void __launch_bounds__(RESID_THREADS, (VPL <= 4 ? 8 : 4))  // 64 registers up to D = 1024, 128 above
resid_kernel(const AT* __restrict__ u, const AT* __restrict__ x_in, const AT* __restrict__ gate,
             const float* __restrict__ gamma, const float* __restrict__ beta, AT* __restrict__ x_out,
             float* __restrict__ rowstats, int64_t M, int D, int reverse) {
  static_assert(sizeof(AT) == 2, "resid_kernel is the bf16 fused-path kernel");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * RESID_WARPS + warp;
  pdl_trigger();
  if (row >= M) return;
  if (reverse) row = M - 1 - row;  // start on the rows the producer wrote last (still in L2)
  pdl_wait();
  const float inv_d = 1.0f / static_cast<float>(D);
  const uint4* up = reinterpret_cast<const uint4*>(u + row * D) + lane;
  const uint4* xp = reinterpret_cast<const uint4*>(x_in + row * D) + lane;
  const uint4* gp = reinterpret_cast<const uint4*>(gate + row * D) + lane;
  uint4 uq[VPL], xq[VPL], gq[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) uq[i] = up[i * 32];  // every load of the row in flight before the first use
#pragma unroll
  for (int i = 0; i < VPL; ++i) xq[i] = xp[i * 32];
#pragma unroll
  for (int i = 0; i < VPL; ++i) gq[i] = gp[i * 32];
  float mean, rstd;
  packed_stats<VPL>(uq, inv_d, 1e-5f, mean, rstd);
  uint4* op = reinterpret_cast<uint4*>(x_out + row * D) + lane;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int e = (i * 32 + lane) * 8;
    float uu[8], xx[8], gg[8], ga[8], be[8];
    unpack8(uq[i], uu);
    unpack8(xq[i], xx);
    unpack8(gq[i], gg);
    load8(gamma + e, ga);
    load8(beta + e, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ln = fmaf((uu[j] - mean) * rstd, ga[j], be[j]);
      xx[j] = fmaf(ln, gg[j], xx[j]);
    }
    xq[i] = make_uint4(pack2(xx[0], xx[1]), pack2(xx[2], xx[3]), pack2(xx[4], xx[5]), pack2(xx[6], xx[7]));
    op[i * 32] = xq[i];
  }
  packed_stats<VPL>(xq, inv_d, 1e-6f, mean, rstd);  // statistics of what was stored (rounded)
  if (lane == 0) *reinterpret_cast<float2*>(rowstats + 2 * row) = make_float2(mean, rstd);
}

// v = Wh y + bh  (diffusion_mlp.py:98);  optional Euler update of the fp32 latent (scheduling_cfm.py:136)
template <typename AT, int VPL>
__global__ void __launch_bounds__(THREADS)
headout_kernel(const AT* __restrict__ y, const float* __restrict__ Wh, const float* __restrict__ bh,
               float* __restrict__ v_out, const float* __restrict__ xt_in, float* __restrict__ xt_out, float dt,
               int64_t M, int D, int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  pdl_trigger();
  if (row >= M) return;
  pdl_wait();
  float x[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) load8(y + row * D + (i * 32 + lane) * 8, x[i]);
  for (int t = 0; t < T; ++t) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float w[8];
      load8(Wh + (int64_t)t * D + (i * 32 + lane) * 8, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(x[i][j], w[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + bh[t];
      if (v_out) v_out[row * T + t] = v;
      if (xt_out) xt_out[row * T + t] = __fadd_rn(__fmul_rn(v, dt), xt_in[row * T + t]);
    }
  }
}

// ------------------------------------------------------------------ xyz-token (T = 3) fast paths, bf16
// Point tokens have T = 3, so the patch-embed / velocity-head weights of one lane's features fit in
// registers (D <= 1024) or L1 (above).  Each warp walks rows grid-stride with the next row's inputs already
// in flight, so these kernels run at HBM speed instead of one latency-bound row per warp.
// Arithmetic order is identical to embed_kernel / headout_kernel (same fmaf chains) => bit-identical output.
constexpr int ROWLOOP_CTAS_PER_SM = 2;

// x0 = bp + Wp x_tok for T = 3; weights [3, D] and bias [D] staged once per CTA in shared memory, two rows per
// iteration share every weight load, so the kernel keeps 16+ warps per SM and runs close to its HBM write time
// (the round-1 register-resident version had 8 warps per SM and was issue-bound at 55 us for M = 65 536).
constexpr int EMBED3_CTAS_PER_SM = 2;
template <int VPL>
__global__ void __launch_bounds__(THREADS, EMBED3_CTAS_PER_SM)
embed3_kernel(const float* __restrict__ x_tok, int64_t x_rows, const float* __restrict__ WpT,
              const float* __restrict__ bp, bf16* __restrict__ x_out, float* __restrict__ rowstats, int64_t M,
              int D) {
  extern __shared__ float emb_s[];  // [w0 | w1 | w2 | b], D floats each
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  for (int e = threadIdx.x * 4; e < 3 * D; e += THREADS * 4)
    *reinterpret_cast<float4*>(emb_s + e) = *reinterpret_cast<const float4*>(WpT + e);
  for (int e = threadIdx.x * 4; e < D; e += THREADS * 4)
    *reinterpret_cast<float4*>(emb_s + 3 * D + e) = *reinterpret_cast<const float4*>(bp + e);
  __syncthreads();
  const float inv_d = 1.0f / static_cast<float>(D);
  const int64_t stride = (int64_t)gridDim.x * WARPS * 2;
  for (int64_t row = ((int64_t)blockIdx.x * WARPS + warp) * 2; row < M; row += stride) {
    const bool two = row + 1 < M;
    const int64_t ra = row >= x_rows ? row % x_rows : row;
    const int64_t rb = two ? ((row + 1) >= x_rows ? (row + 1) % x_rows : row + 1) : ra;
    const float a0 = x_tok[ra * 3], a1 = x_tok[ra * 3 + 1], a2 = x_tok[ra * 3 + 2];
    const float b0 = x_tok[rb * 3], b1 = x_tok[rb * 3 + 1], b2 = x_tok[rb * 3 + 2];
    uint4 qa[VPL], qb[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int e = (i * 32 + lane) * 8;
      float w0[8], w1[8], w2[8], bb[8], xa[8], xb[8];
      load8(emb_s + e, w0);
      load8(emb_s + D + e, w1);
      load8(emb_s + 2 * D + e, w2);
      load8(emb_s + 3 * D + e, bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // same fmaf chain as embed_kernel: bias, then t = 0, 1, 2
        xa[j] = fmaf(a2, w2[j], fmaf(a1, w1[j], fmaf(a0, w0[j], bb[j])));
        xb[j] = fmaf(b2, w2[j], fmaf(b1, w1[j], fmaf(b0, w0[j], bb[j])));
      }
      qa[i] = make_uint4(pack2(xa[0], xa[1]), pack2(xa[2], xa[3]), pack2(xa[4], xa[5]), pack2(xa[6], xa[7]));
      qb[i] = make_uint4(pack2(xb[0], xb[1]), pack2(xb[2], xb[3]), pack2(xb[4], xb[5]), pack2(xb[6], xb[7]));
      reinterpret_cast<uint4*>(x_out + row * D)[i * 32 + lane] = qa[i];
      if (two) reinterpret_cast<uint4*>(x_out + (row + 1) * D)[i * 32 + lane] = qb[i];
    }
    float mean, rstd;
    packed_stats<VPL>(qa, inv_d, 1e-6f, mean, rstd);  // statistics of what is stored
    if (lane == 0) *reinterpret_cast<float2*>(rowstats + 2 * row) = make_float2(mean, rstd);
    if (two) {
      packed_stats<VPL>(qb, inv_d, 1e-6f, mean, rstd);
      if (lane == 0) *reinterpret_cast<float2*>(rowstats + 2 * (row + 1)) = make_float2(mean, rstd);
    }
  }
}

template <int VPL, bool WREG>
__global__ void __launch_bounds__(THREADS, ROWLOOP_CTAS_PER_SM)
headout3_kernel(const bf16* __restrict__ y, const float* __restrict__ Wh, const float* __restrict__ bh,
                float* __restrict__ v_out, const float* __restrict__ xt_in, float* __restrict__ xt_out, float dt,
                int64_t M, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * WARPS;
  int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  pdl_trigger();
  pdl_wait();
  extern __shared__ float wh_s[];  // !WREG: [3, D] head weights staged once per CTA
  float w[WREG ? 3 : 1][VPL][8];
  if (WREG) {
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int i = 0; i < VPL; ++i) load8(Wh + (int64_t)t * D + (i * 32 + lane) * 8, w[t][i]);
  } else {
    for (int e = threadIdx.x * 4; e < 3 * D; e += THREADS * 4)
      *reinterpret_cast<float4*>(wh_s + e) = *reinterpret_cast<const float4*>(Wh + e);
    __syncthreads();
  }
  if (row >= M) return;
  const float bias = lane < 3 ? bh[lane] : 0.f;
  uint4 cur[VPL], nxt[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) cur[i] = reinterpret_cast<const uint4*>(y + row * D)[i * 32 + lane];
  for (; row < M; row += stride) {
    const bool more = row + stride < M;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      nxt[i] = more ? reinterpret_cast<const uint4*>(y + (row + stride) * D)[i * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
    const float xin = (xt_out != nullptr && lane < 3) ? xt_in[row * 3 + lane] : 0.f;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float x[8];
      unpack8(cur[i], x);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        if (WREG) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t] = fmaf(x[j], w[t][i][j], acc[t]);
        } else {
          float wv[8];
          load8(wh_s + t * D + (i * 32 + lane) * 8, wv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t] = fmaf(x[j], wv[j], acc[t]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
      acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], o);
      acc[2] += __shfl_xor_sync(0xffffffffu, acc[2], o);
    }
    if (lane < 3) {  // lane t finishes output t
      const float v = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : acc[2]) + bias;
      if (v_out) v_out[row * 3 + lane] = v;
      if (xt_out) xt_out[row * 3 + lane] = __fadd_rn(__fmul_rn(v, dt), xin);
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) cur[i] = nxt[i];
  }
}

// Velocity head of a GUIDED step with the guidance combine and the Euler update in the same kernel (no renorm):
// row r of the conditional pass, row r + Mx of the unconditional pass (and r + 2 Mx of the third pass) share latent
// row r, so one warp forms the two (three) head outputs, combines them as GuidanceScaler.scale does
// (guidance_scaler.py:74-87; the formulas of cfg_euler_kernel with ratio == 1) and steps the latent
// (scheduling_cfm.py:136).  Each head output is the same fmaf chain and butterfly as in headout3_kernel, so the result
// is bit-identical to headout3 + cfg_euler_kernel; the [passes, Mx, 3] velocity tensor is never written.
template <int VPL, bool WREG>
__global__ void __launch_bounds__(THREADS, ROWLOOP_CTAS_PER_SM)
headout3_cfg_kernel(const bf16* __restrict__ y, const float* __restrict__ Wh, const float* __restrict__ bh,
                    float* __restrict__ x_sel, float dt, int64_t Mx, int D, int passes, int mode, float scale,
                    float scale3) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * WARPS;
  pdl_trigger();
  pdl_wait();
  extern __shared__ float wh_s[];  // !WREG: [3, D] head weights staged once per CTA
  float w[WREG ? 3 : 1][VPL][8];
  if (WREG) {
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int i = 0; i < VPL; ++i) load8(Wh + (int64_t)t * D + (i * 32 + lane) * 8, w[t][i]);
  } else {
    for (int e = threadIdx.x * 4; e < 3 * D; e += THREADS * 4)
      *reinterpret_cast<float4*>(wh_s + e) = *reinterpret_cast<const float4*>(Wh + e);
    __syncthreads();
  }
  const float bias = lane < 3 ? bh[lane] : 0.f;
  for (int64_t row = (int64_t)blockIdx.x * WARPS + warp; row < Mx; row += stride) {
    uint4 q[3][VPL];  // the passes' rows of y, all requested before the first is used
#pragma unroll
    for (int ps = 0; ps < 3; ++ps) {
      if (ps < passes) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) q[ps][i] = reinterpret_cast<const uint4*>(y + (row + ps * Mx) * D)[i * 32 + lane];
      }
    }
    const float xin = lane < 3 ? x_sel[row * 3 + lane] : 0.f;
    float vp[3] = {0.f, 0.f, 0.f};  // lane t < 3: output t of pass ps
#pragma unroll
    for (int ps = 0; ps < 3; ++ps) {
      if (ps < passes) {
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          float x[8];
          unpack8(q[ps][i], x);
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            if (WREG) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[t] = fmaf(x[j], w[t][i][j], acc[t]);
            } else {
              float wv[8];
              load8(wh_s + t * D + (i * 32 + lane) * 8, wv);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[t] = fmaf(x[j], wv[j], acc[t]);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
          acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], o);
          acc[2] += __shfl_xor_sync(0xffffffffu, acc[2], o);
        }
        vp[ps] = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : acc[2]) + bias;
      }
    }
    if (lane < 3) {
      const float c = vp[0], u = vp[1], t3 = vp[2];
      float v;
      if (mode == 0) v = fmaf(c - u, scale, u);
      else if (mode == 1) v = fmaf(t3 - u, scale3, fmaf(c - t3, scale, u));
      else v = fmaf(c - t3, scale3, fmaf(c - u, scale, u));
      x_sel[row * 3 + lane] = __fadd_rn(__fmul_rn(v, dt), xin);
    }
  }
}

int num_sms_rw();  // SM count of the current device (runtime.cu)
inline unsigned rowloop_grid(int64_t M, int ctas_per_sm) {
  const int64_t want = ceil_div(M, WARPS), cap = (int64_t)num_sms_rw() * ctas_per_sm;
  return (unsigned)(want < cap ? want : cap);
}

template <typename AT, template <typename, int> class Launcher, typename... Args>
int dispatch_vpl(int D, Args... args) {
  switch (D / 256) {
    case 1: return Launcher<AT, 1>::run(args...);
    case 2: return Launcher<AT, 2>::run(args...);
    case 3: return Launcher<AT, 3>::run(args...);
    case 4: return Launcher<AT, 4>::run(args...);
    case 5: return Launcher<AT, 5>::run(args...);
    case 6: return Launcher<AT, 6>::run(args...);
    case 7: return Launcher<AT, 7>::run(args...);
    case 8: return Launcher<AT, 8>::run(args...);
    default: break;
  }
  set_error("unsupported head width %d (multiple of 256, <= 2048)", D);
  return NOVA_ERR_INVALID;
}
template <typename AT, int VPL>
struct EmbedLauncher {
  static int run(const float* x_tok, int64_t x_rows, const float* WpT, const float* bp, AT* x_out, float* rowstats,
                 int64_t M, int D, int T, cudaStream_t s) {
    if (T == 3 && sizeof(AT) == 2)
      launch_pdl(embed3_kernel<VPL>, dim3(rowloop_grid(ceil_div(M, 2), EMBED3_CTAS_PER_SM)), dim3(THREADS),
                 4 * D * sizeof(float), s, x_tok, x_rows, WpT, bp, reinterpret_cast<bf16*>(x_out), rowstats, M, D);
    else
      launch_pdl(embed_kernel<AT, VPL>, dim3((unsigned)ceil_div(M, WARPS)), dim3(THREADS), 0, s, x_tok, x_rows, WpT, bp,
                 x_out, rowstats, M, D, T);
    NOVA_CHECK_LAUNCH();
    return NOVA_OK;
  }
};
template <typename AT, int VPL>
struct ResidLauncher {
  static int run(const AT* u, const AT* x_in, const AT* gate, const float* gamma, const float* beta, AT* x_out,
                 float* rowstats, int64_t M, int D, int reverse, cudaStream_t s) {
    launch_pdl(resid_kernel<AT, VPL>, dim3((unsigned)ceil_div(M, RESID_WARPS)), dim3(RESID_THREADS), 0, s, u, x_in, gate,
               gamma, beta, x_out, rowstats, M, D, reverse);
    NOVA_CHECK_LAUNCH();
    return NOVA_OK;
  }
};
template <typename AT, int VPL>
struct HeadoutLauncher {
  static int run(const AT* y, const float* Wh, const float* bh, float* v_out, const float* xt_in, float* xt_out,
                 float dt, int64_t M, int D, int T, cudaStream_t s) {
    if (T == 3 && sizeof(AT) == 2)
      launch_pdl(headout3_kernel<VPL, (VPL <= 3)>, dim3(rowloop_grid(M, ROWLOOP_CTAS_PER_SM)), dim3(THREADS),
                 (VPL <= 3) ? 0 : 3 * D * sizeof(float), s, reinterpret_cast<const bf16*>(y), Wh, bh, v_out, xt_in,
                 xt_out, dt, M, D);
    else
      launch_pdl(headout_kernel<AT, VPL>, dim3((unsigned)ceil_div(M, WARPS)), dim3(THREADS), 0, s, y, Wh, bh, v_out,
                 xt_in, xt_out, dt, M, D, T);
    NOVA_CHECK_LAUNCH();
    return NOVA_OK;
  }
};

// (mean, rstd) of every row of x [M, D]; generic two-pass form used by the debug hook only
template <typename AT>
__global__ void __launch_bounds__(THREADS)
rowstats_kernel(const AT* __restrict__ x, float* __restrict__ rowstats, int64_t M, int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= M) return;
  const AT* xr = x + row * D;
  float s = 0.f;
  for (int e = lane; e < D; e += 32) s += to_float(xr[e]);
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
  for (int e = lane; e < D; e += 32) {
    const float d = to_float(xr[e]) - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
  if (lane == 0) *reinterpret_cast<float2*>(rowstats + 2 * row) = make_float2(mean, rstd);
}

// Pre-embedded rows (DiffusionMLP.forward with a 3-D x: PatchEmbed passes it through, embeddings.py:165) entering the
// fused dataflow: the residual stream starts as a copy of the caller's rows, with their LayerNorm statistics
template <typename AT>
__global__ void __launch_bounds__(THREADS)
adopt_rows_kernel(const AT* __restrict__ x_emb, AT* __restrict__ x_out, float* __restrict__ rowstats, int64_t M, int D,
                  float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= M) return;
  const AT* xr = x_emb + row * D;
  float s = 0.f;
  for (int e = lane; e < D; e += 32) {
    const AT v = xr[e];
    x_out[row * D + e] = v;
    s += to_float(v);
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
  for (int e = lane; e < D; e += 32) {
    const float d = to_float(xr[e]) - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
  if (lane == 0) *reinterpret_cast<float2*>(rowstats + 2 * row) = make_float2(mean, rstd);
}

// AdaLN projection rows packed for EPI_ADALN: per 128 features [128 scale rows | 128 shift rows], then the
// gate rows unchanged.  src is the reference layout [scale (D) | shift (D) | gate (D, optional)] x row_len.
template <typename TS, typename TD>
__global__ void pack_adaln_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t rows, int64_t row_len,
                                  int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * row_len) return;
  const int64_t dr = i / row_len, c = i % row_len;
  int64_t sr = dr;
  if (dr < 2 * (int64_t)D) {
    const int64_t grp = dr / 256, r = dr % 256;
    sr = r < 128 ? grp * 128 + r : (int64_t)D + grp * 128 + (r - 128);
  }
  dst[i] = from_float<TD>(to_float(src[sr * row_len + c]));
}

// ------------------------------------------------------------------ latent bookkeeping
// Guided Euler update over one cloud per CTA (needs per-cloud norms for renorm):
//   v = vu + (vc - vu) * s;  v *= clamp(|vc| / |v|, renorm, 1) if renorm < 1;  x += dt * v
// vc = v2[b], vu = v2[B + b], each [n*T]; x_sel [B, n*T].
// The reference takes the renorm norms over the head's full (N, T) output, in which tokens outside
// pred_ids carry the latent itself for both passes (diffusion_mlp.py:99): `extra_sumsq[b] * extra_scale`
// adds that contribution (sum of squares of the unpredicted latent rows at this step) to both norms.
// Three-pass forms (guidance_scaler.py:78-85), v3 = v2[2B + b]:
//   mode 1 (image):          v = renorm(vu + (vc - v3) * s) + (v3 - vu) * s3
//   mode 2 (spatiotemporal): v = renorm(vu + (vc - vu) * s) + (vc - v3) * s3
// On rows outside pred_ids all passes carry the latent, so the extra term vanishes there.
static __global__ void __launch_bounds__(256)
cfg_euler_kernel(const float* __restrict__ v2, float* __restrict__ x_sel, int64_t B, int64_t len, float scale,
                 float renorm, float dt, float* __restrict__ extra_sumsq, float* __restrict__ ratio_out, int mode,
                 float scale3) {
  __shared__ float red[2][8];
  pdl_trigger();
  pdl_wait();
  const int64_t b = blockIdx.x;
  const float* vc = v2 + b * len;
  const float* vu = v2 + (B + b) * len;
  const float* v3 = v2 + (2 * B + b) * len;  // only read when mode != 0
  float* x = x_sel + b * len;
  float ratio = 1.0f;
  if (renorm < 1.0f) {
    float sc = 0.f, sv = 0.f;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
      const float c = vc[i], u = vu[i];
      const float v = fmaf(c - (mode == 1 ? v3[i] : u), scale, u);
      sc = fmaf(c, c, sc);
      sv = fmaf(v, v, sv);
    }
    sc = warp_sum(sc);
    sv = warp_sum(sv);
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = sc;
      red[1][threadIdx.x >> 5] = sv;
    }
    __syncthreads();
    float tc = 0.f, tv = 0.f;
    for (int w = 0; w < 8; ++w) {
      tc += red[0][w];
      tv += red[1][w];
    }
    if (extra_sumsq != nullptr) {
      tc += extra_sumsq[b];
      tv += extra_sumsq[b];
    }
    ratio = fminf(fmaxf(sqrtf(tc) / sqrtf(tv), renorm), 1.0f);
    __syncthreads();
    if (threadIdx.x == 0) {
      // the reference scales the WHOLE head output by ratio, including the rows that only carry the
      // latent: those rows then move as x <- x + dt*ratio*x, and so does their sum of squares.
      if (extra_sumsq != nullptr) {
        const float gr = 1.0f + dt * ratio;
        extra_sumsq[b] *= gr * gr;
      }
      if (ratio_out != nullptr) ratio_out[b] = ratio;
    }
  }
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
    const float c = vc[i], u = vu[i];
    float v;
    if (mode == 0) {
      v = fmaf(c - u, scale, u) * ratio;
    } else {
      const float t = v3[i];
      v = mode == 1 ? fmaf(t - u, scale3, fmaf(c - t, scale, u) * ratio) : fmaf(c - t, scale3, fmaf(c - u, scale, u) * ratio);
    }
    x[i] = __fadd_rn(__fmul_rn(v, dt), x[i]);
  }
}

// extra[b] = sum(noise[b]^2 over all N*T) - sum(x_sel[b]^2 over the n*T selected), in double
static __global__ void __launch_bounds__(256)
unpred_sumsq_kernel(const float* __restrict__ noise, const float* __restrict__ x_sel, int64_t all_len,
                    int64_t sel_len, float* __restrict__ extra) {
  __shared__ double red[8];
  const int64_t b = blockIdx.x;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < all_len; i += blockDim.x) {
    const double v = noise[b * all_len + i];
    acc += v * v;
  }
  for (int64_t i = threadIdx.x; i < sel_len; i += blockDim.x) {
    const double v = x_sel[b * sel_len + i];
    acc -= v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    extra[b] = static_cast<float>(t > 0.0 ? t : 0.0);
  }
}

// out[b, tok, :] for predicted tokens <- x_sel;  (pred_ids == nullptr: plain copy)
static __global__ void
scatter_tok_kernel(const float* __restrict__ x_sel, const IdsView ids,
                                   float* __restrict__ out, int64_t Bx, int64_t N, int64_t n, int T, uint32_t* bad_ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bx * n * T) return;
  const int64_t row = i / T, c = i % T;
  const int64_t b = row / n;
  const int64_t tok = ids.at(b, row - b * n, row % n);
  if (tok < 0 || tok >= N) {  // never write through an out-of-range id
    if (bad_ids && c == 0) *bad_ids = 0xBAD1D5u;
    return;
  }
  out[(b * N + tok) * T + c] = x_sel[i];
}
// Tokens outside pred_ids: the head returns its own input there, so the reference's loop
// does x <- x*dt + x every step (two roundings); reproduce that recurrence exactly.
static __global__ void
unpredicted_kernel(const float* __restrict__ noise, float* __restrict__ out, int64_t numel, int64_t per_cloud,
                   int64_t Bx, const StepList dts_list, const float* __restrict__ ratios, int S) {
  const float* dts = dts_list.v;  // by value: this kernel also runs for an EMPTY set, when there is no workspace
  // With guidance renorm the head output (these rows included) is first scaled by the per-cloud
  // ratio of that step: ratios[s * Bx + b]; nullptr means 1.
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numel) return;
  const int64_t b = i / per_cloud;
  float x = noise[i];
  for (int s = 0; s < S; ++s) {
    const float v = ratios ? __fmul_rn(x, ratios[s * Bx + b]) : x;
    x = __fadd_rn(__fmul_rn(v, dts[s]), x);
  }
  out[i] = x;
}
static __global__ void fill_kernel(float* __restrict__ dst, int64_t n, float value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

template <typename T>
__global__ void euler_kernel(const T* __restrict__ v, const T* __restrict__ x, T* __restrict__ out, int64_t numel,
                             float dt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numel) return;
  // model_output.mul(dt).add_(sample): the product is rounded to the tensor dtype first.
  const T prod = from_float<T>(__fmul_rn(to_float(v[i]), dt));
  out[i] = from_float<T>(__fadd_rn(to_float(prod), to_float(x[i])));
}

// fp32 -> AT / permuting copies used when packing weights
template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t numel) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < numel) dst[i] = from_float<TD>(to_float(src[i]));
}
// [D, T] -> [T, D]
static __global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int D, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)D * T) return;
  const int t = static_cast<int>(i / D), d = static_cast<int>(i % D);
  dst[i] = src[(int64_t)d * T + t];
}
// Conv2d weight (D, C, p, p) -> token order (D, p, p, C)
template <typename TS>
__global__ void permute_patch_kernel(const TS* __restrict__ src, float* __restrict__ dst, int D, int C, int p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int T = C * p * p;
  if (i >= (int64_t)D * T) return;
  const int d = static_cast<int>(i / T), r = static_cast<int>(i % T);
  const int c = r % C, pw = (r / C) % p, ph = r / (C * p);
  dst[i] = to_float(src[((int64_t)(d * C + c) * p + ph) * p + pw]);
}


template <typename AT, int VPL>
struct HeadoutCfgLauncher {
  static int run(const AT* y, const float* Wh, const float* bh, float* x_sel, float dt, int64_t Mx, int D, int passes,
                 int mode, float scale, float scale3, cudaStream_t s) {
    launch_pdl(headout3_cfg_kernel<VPL, (VPL <= 3)>, dim3(rowloop_grid(Mx, ROWLOOP_CTAS_PER_SM)), dim3(THREADS),
               (VPL <= 3) ? 0 : 3 * D * sizeof(float), s, reinterpret_cast<const bf16*>(y), Wh, bh, x_sel, dt, Mx, D, passes,
               mode, scale, scale3);
    NOVA_CHECK_LAUNCH();
    return NOVA_OK;
  }
};

// ------------------------------------------------------------------ out-of-line entry points
// The heavy instantiations (8 widths x kernel variants) live in their own translation units so that they build
// in parallel with head.cu:  rowwise_row_f32.cu, rowwise_row_bf16.cu (launch_row) and rowwise_fused.cu (the
// fused-dataflow kernels).  head.cu calls only these.
int launch_row_f32(const RowParams& p, bool has_prev, int out, cudaStream_t stream);
int launch_row_bf16(const RowParams& p, bool has_prev, int out, cudaStream_t stream);
int embed_bf16(const float* x_tok, int64_t x_rows, const float* WpT, const float* bp, bf16* x_out, float* rowstats,
               int64_t M, int D, int T, cudaStream_t stream);
int resid_bf16(const bf16* u, const bf16* x_in, const bf16* gate, const float* gamma, const float* beta, bf16* x_out,
               float* rowstats, int64_t M, int D, int reverse, cudaStream_t stream);
int headout_bf16(const bf16* y, const float* Wh, const float* bh, float* v_out, const float* xt_in, float* xt_out, float dt,
                 int64_t M, int D, int T, cudaStream_t stream);
// T == 3 only: head + guidance combine (passes = 2 | 3 over Mx latent rows each) + Euler step of x_sel in place
int headout_cfg_bf16(const bf16* y, const float* Wh, const float* bh, float* x_sel, float dt, int64_t Mx, int D, int passes,
                     int mode, float scale, float scale3, cudaStream_t stream);

}  // namespace rw
}  // namespace nova
