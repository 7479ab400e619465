// The block chain of one diffusion step as ONE cluster kernel (small / medium M: the set-by-set pattern).
//
// After the AdaLN statistics GEMM of a step (st [M, 20 D], one big tcgen05 GEMM), the rest of the step is a chain of
// 19 dependent stages over the SAME rows (reference: diffnext/models/diffusion_mlp.py:48-53,89-99):
//   R0            x = PatchEmbed(x_tok);  h = LN(x)(1+scale_0)+shift_0
//   per block i   F1: u1 = silu(h P1^T + p1)     F2: u2 = u1 P2^T + p2
//                 R : x += LN_aff(u2) * gate_i;  h = LN(x)(1+scale_{i+1})+shift_{i+1}
//   last R        y = LN(x)(1+scale_f)+shift_f;  v = y H^T + h0;  x_tok += dt v     (Euler, scheduling_cfm.py:136)
// As separate launches each stage costs 5-8 us whatever M is (launch + prologue + first-operand latency + drain), i.e.
// ~110-180 us per step for 32..1632 rows.  Rows are independent, so here a CLUSTER of 8 CTAs owns 128 rows for the
// whole chain and nothing but cluster barriers (~0.2 us) separates the stages; there is no grid-wide dependency, so
// any number of clusters may run in any order (no co-residency requirement, no grid barrier to deadlock on).
//
//   F stages: CTA c of the cluster computes output columns [c w, (c+1) w), w = D / 8, of the 128 rows:
//             tcgen05.mma cta_group::1, M = 128, N = w, accumulators in TMEM; A (h or u1, all K) and its W slice are
//             streamed by TMA through separate rings.  The W ring runs AHEAD across stage boundaries (weights do not
//             depend on the previous stage): while a stage drains, the next stage's weights are already landing.
//   R stages: the 64 warps of the cluster take the 128 rows (rw::row_body, the arithmetic of the fused row kernel).
// Activations (x, h, u1, u2: 128 x D bf16 each per cluster) stay in L2; stores are made visible to the other CTAs'
// TMA loads by fence.proxy.async + barrier.cluster (release / acquire).
//
// This header is compiled by chain_r64.cu and chain_r128.cu (rows per cluster; NOVA_CHAIN_ROWS selects which).
#pragma once

#include <cuda.h>

#define NOVA_GEMM_TU 99  // PTX wrappers only, no GEMM instantiations
#include "gemm_tcgen05.cuh"
#include "rowwise.cuh"
#include "chain_api.cuh"

namespace nova {
namespace chain {

using namespace nova::tc;

constexpr int CL = 8;                  // CTAs per cluster
constexpr int CH_THREADS = 256;
constexpr int SMEM_LIMIT = 227 * 1024;

// ROWS = rows per cluster: 128, or 64 (twice the clusters, half the activation bytes every CTA has to take in per
// stage and one row per warp in the R stages: the better choice while 8 * ceil(M / 64) CTAs still fit on the chip).
// The A operand of a stage (h or u1, ROWS x D) is needed by all 8 CTAs: the k-blocks are dealt round-robin to the
// CTAs, each loads its k-blocks ONCE and MULTICASTS them into the same shared-memory offset of all 8 CTAs (a TMA
// issue costs the issuing thread ~250 cycles, so 12 / 8 issues per CTA and stage instead of 12 matter).
template <int VPL, int ROWS>
struct ChainPlan {
  static_assert(ROWS == 64 || ROWS == 128, "rows per cluster");
  static constexpr int W_COLS = 32 * VPL;             // output columns per CTA (D / 8)
  static constexpr int NUM_K = 4 * VPL;               // D / 64 k-blocks per stage
  // The single-thread producer / MMA loops pay ~100-250 cycles per mbarrier wait, tcgen05.commit and TMA issue, far
  // more than the 4 MMAs of one k-block take, so the rings are managed in GROUPS of KG k-blocks: one wait, one
  // expect_tx and one commit per group (measured: 845 cycles per k-block with per-k-block barriers).
  static constexpr int KG_A = VPL <= 3 ? 4 : 2;       // k-blocks per A group
  static constexpr int KG_W = 2;                      // k-blocks per W group (finer: the W ring is what runs ahead)
  static constexpr int GPS_A = NUM_K / KG_A;          // groups per stage
  static constexpr int GPS_W = NUM_K / KG_W;
  static constexpr int A_KB_BYTES = ROWS * BK * 2;    // one k-block of A: 16 KB / 8 KB
  static constexpr int W_KB_BYTES = W_COLS * BK * 2;  // one k-block of the W slice (multiple of 1024)
  static constexpr int A_BYTES = KG_A * A_KB_BYTES;   // one group
  static constexpr int W_BYTES = KG_W * W_KB_BYTES;
  static constexpr int PAD = 0;
  static constexpr int AVAIL = SMEM_LIMIT - 3072 - PAD;
  // A ring: a whole stage if that is <= 96 KB (64 rows) / 128 KB (128 rows), so that every load of a stage is issued
  // the moment its barrier opens; the W ring takes the rest (at least two groups) and runs ahead of the stage boundary
  static constexpr int A_BUDGET = (ROWS == 64 ? 96 : 128) * 1024 / A_BYTES;
  static constexpr int A_FIT = (AVAIL - 2 * W_BYTES) / A_BYTES;
  static constexpr int A_WANT = A_BUDGET < A_FIT ? A_BUDGET : A_FIT;
  static constexpr int A_STAGES = A_WANT < GPS_A ? A_WANT : GPS_A;   // ring depths in groups
  static constexpr int W_FIT = (AVAIL - A_STAGES * A_BYTES) / W_BYTES;
  static constexpr int W_STAGES = W_FIT < GPS_W ? W_FIT : GPS_W;
  static constexpr int OFF_W = A_STAGES * A_BYTES + PAD;
  static constexpr int OFF_BIAS = OFF_W + W_STAGES * W_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 1024;     // w <= 256 floats
  static constexpr int SMEM_BYTES = OFF_BAR + 512;
  static constexpr int TMEM_COLS_ALLOC = W_COLS <= 32 ? 32 : W_COLS <= 64 ? 64 : W_COLS <= 128 ? 128 : 256;
  static_assert(NUM_K % KG_A == 0 && KG_A % KG_W == 0, "k-blocks per stage must be whole groups");
  static_assert(A_STAGES >= 1 && (GPS_A == 1 || A_STAGES >= 2) && W_STAGES >= 2, "rings too shallow");
  static_assert(SMEM_BYTES <= SMEM_LIMIT, "shared memory plan exceeds 227 KB");
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 2D tiled TMA load multicast to the CTAs in `mask`: data and complete_tx land at the same offsets in each of them
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* t, uint32_t bar, uint32_t dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(t)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once all prior MMAs of this thread have retired
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

// one lane of a converged warp (the others skip); keeps the surrounding address arithmetic warp-uniform
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Stage boundary: every global store of this CTA becomes visible to the whole cluster, generic and async proxy alike.
// The proxy fence sits on the WRITER side only (generic-proxy stores -> the TMA loads another CTA issues after the
// barrier), as in any st.shared -> fence.proxy.async -> barrier -> TMA-store epilogue; a second fence behind the barrier
// cost 1.2-1.5 % of a call (2.43 -> 2.40 ms at 32 rows) and orders nothing that is read through the async proxy before being written by it.
__device__ __forceinline__ void stage_barrier() {
  __syncwarp();
  fence_proxy_async_all();
  cluster_sync_all();  // arrive.release + wait.acquire, all threads of all 8 CTAs
#ifdef NOVA_CHAIN_DOUBLE_FENCE
  fence_proxy_async_all();
#endif
}

template <int VPL, int ROWS>
__global__ void __launch_bounds__(CH_THREADS, 1)
chain_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_u1,
             const __grid_constant__ CUtensorMap tmap_w, const ChainParams p, uint32_t* dbg) {
  using P = ChainPlan<VPL, ROWS>;
  constexpr int W_COLS = P::W_COLS, KG_A = P::KG_A, KG_W = P::KG_W, GPS_A = P::GPS_A, GPS_W = P::GPS_W;
  constexpr int W_STAGES = P::W_STAGES, A_STAGES = P::A_STAGES;
  constexpr int A_BYTES = P::A_BYTES;
  constexpr uint16_t ALL_CTAS = (1u << CL) - 1u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) {
    if (dbg && threadIdx.x == 0) { dbg[0] = 0xDEAD0A12u; dbg[1] = base; __threadfence_system(); }
    __trap();
  }
  const uint32_t bar_base = base + P::OFF_BAR;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (A_STAGES + s); };
  auto fullW = [&](int s) { return bar_base + 8u * (2 * A_STAGES + s); };
  auto emptyW = [&](int s) { return bar_base + 8u * (2 * A_STAGES + W_STAGES + s); };
  const uint32_t tfull = bar_base + 8u * (2 * A_STAGES + 2 * W_STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P::OFF_BAR + 8 * (2 * A_STAGES + 2 * W_STAGES + 1));
  float* bias_s = reinterpret_cast<float*>(smem + P::OFF_BIAS);

  pdl_trigger();
  const long long tl_t0 = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // timeline slot (stage, k): stage 0 = prologue + R0, stage 1 + f = F stage f (and the R stage that follows an F2)
  //   k = 0 barrier left (thread 0) | 1 all A loads issued (thread 0) | 2 first operands landed (MMA thread)
  //   3 MMAs issued (MMA thread) | 4 accumulator complete (epilogue thread) | 5 epilogue stores done
  //   6 R stage entered (thread 0) | 7 R stage rows done (thread 0)
  auto stamp = [&](int stage, int k) {
    if (p.timeline != nullptr && blockIdx.x == 0 && stage < 64) p.timeline[stage * 8 + k] = clock64() - tl_t0;
  };
  const int64_t row0 = static_cast<int64_t>(blockIdx.x / CL) * ROWS;  // first row of this cluster
  const int D = p.D, depth = p.depth;
  const int n_f = 2 * depth;           // F stages of the step
  const int total_ga = n_f * GPS_A, total_gw = n_f * GPS_W;  // k-block groups over all F stages

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_h);
    prefetch_tmap(&tmap_u1);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    // emptyA: one multicast tcgen05.commit arrival from each of the 8 CTAs (a slot is free once ALL have consumed it)
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), CL); }
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(fullW(s), 1); mbar_init(emptyW(s), 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(const_cast<uint32_t*>(tmem_slot)), P::TMEM_COLS_ALLOC);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- W producer state (warp 0, lane 0): weights never depend on this or the preceding kernel
  int w_issued = 0;
  auto issue_w = [&](int g) {  // global W group index over all F stages
    const int f = g / GPS_W, j = g - f * GPS_W, s = g % W_STAGES;
    mbar_wait(emptyW(s), (static_cast<uint32_t>(g / W_STAGES) & 1u) ^ 1u, dbg, 0x500u | s);
    mbar_expect_tx(fullW(s), P::W_BYTES);
#pragma unroll
    for (int k = 0; k < KG_W; ++k)
      tma_load_2d(&tmap_w, fullW(s), base + P::OFF_W + s * P::W_BYTES + k * P::W_KB_BYTES, (j * KG_W + k) * BK,
                  f * D + static_cast<int>(rank) * W_COLS);
  };
  if (warp == 0 && lane == 0) {
    const int lim = total_gw < W_STAGES ? total_gw : W_STAGES;
    while (w_issued < lim) issue_w(w_issued++);
  }
  pdl_wait();  // from here on: st (the statistics GEMM of this step) and x_tok (the previous step) are read
  if (threadIdx.x == 0) stamp(0, 6);

  // ---- row stages: warp (rank, warp) of the cluster takes rows g (and g + 64 when the cluster owns 128 rows)
  rw::RowParams rp{};
  rp.M = p.M; rp.D = D; rp.T = p.T;
  rp.x_in = p.x; rp.x_out = p.x; rp.u = p.u2; rp.st = p.st; rp.ldst = p.ldst; rp.h_out = p.h;
  rp.x_tok = p.x_tok; rp.x_rows = p.x_rows; rp.Wp = p.Wp; rp.WpT = p.WpT; rp.bp = p.bp; rp.Wh = p.Wh; rp.bh = p.bh;
  rp.v_out = p.v_out; rp.xt_in = p.x_tok; rp.xt_out = p.xt_out; rp.dt = p.dt;
  const int64_t final_off = static_cast<int64_t>(3) * depth * D;
  const int wg = static_cast<int>(rank) * (CH_THREADS / 32) + warp;  // 0..63

  {  // R0: patch embed + first modulation (or, depth == 0, straight to the head)
    rp.scale_off = depth > 0 ? 0 : final_off;
    for (int r = wg; r < ROWS; r += CL * (CH_THREADS / 32)) {
      const int64_t row = row0 + r;
      if (row < p.M) {
        if (depth > 0) rw::row_body<bf16, VPL, false, 0, true>(rp, row, lane);
        else rw::row_body<bf16, VPL, false, 1, true>(rp, row, lane);
      }
    }
    if (threadIdx.x == 0) stamp(0, 7);
  }

  for (int f = 0; f < n_f; ++f) {
    stage_barrier();  // the A operand of stage f (h or u1) is complete and visible cluster-wide
    if (threadIdx.x == 0) stamp(1 + f, 0);
    const int blk = f >> 1;
    const bool second = (f & 1) != 0;  // F2 of the block
    if (warp == 0) {
      if (lane == 0) {  // ------------------------------------------------ TMA producer
        const CUtensorMap* ta = second ? &tmap_u1 : &tmap_h;
        for (int j = 0; j < GPS_A; ++j) {
          const int g = f * GPS_A + j, s = g % A_STAGES;
          mbar_wait(emptyA(s), (static_cast<uint32_t>(g / A_STAGES) & 1u) ^ 1u, dbg, 0x100u | s);
          mbar_expect_tx(fullA(s), A_BYTES);  // KG_A k-blocks, whoever of the 8 CTAs multicasts them
#pragma unroll
          for (int k = 0; k < KG_A; ++k) {
            const int kb = j * KG_A + k;
            if (static_cast<uint32_t>(kb + f) % CL == rank)  // my turn (rotated by stage: 12 k-blocks over 8 CTAs)
              tma_load_2d_mc(ta, fullA(s), base + s * A_BYTES + k * P::A_KB_BYTES, kb * BK, static_cast<int>(row0), ALL_CTAS);
          }
          // the W groups this A group is multiplied with, if the run-ahead has not issued them yet
          const int gw_need = f * GPS_W + (j + 1) * (KG_A / KG_W);
          while (w_issued < gw_need) issue_w(w_issued++);
        }
        stamp(1 + f, 1);
        // run ahead: weights of the following stage(s) into the slots this stage's MMAs free
        const int ahead = (f + 1) * GPS_W + W_STAGES;
        const int lim = total_gw < ahead ? total_gw : ahead;
        while (w_issued < lim) issue_w(w_issued++);
      }
    } else if (warp == 1) {
      {  // ---------------------------------------------------------------- MMA issuer
        // The whole warp walks the loop (waits, descriptor arithmetic: warp-uniform, so it can stay in uniform
        // registers) and one elected lane issues; with a single-lane loop every tcgen05.mma cost ~61 cycles of
        // R2UR traffic + issue, more than the instruction takes on the tensor pipe at these tile sizes.
        constexpr uint32_t idesc = make_idesc_bf16(ROWS, W_COLS);
        tcgen05_fence_after();
        for (int j = 0; j < GPS_W; ++j) {
          const int gw = f * GPS_W + j, sw = gw % W_STAGES;
          const int kb0 = j * KG_W;                       // first k-block of this W group
          const int ga = f * GPS_A + kb0 / KG_A, sa = ga % A_STAGES;
          if (kb0 % KG_A == 0) mbar_wait(fullA(sa), static_cast<uint32_t>(ga / A_STAGES) & 1u, dbg, 0x300u | sa);
          mbar_wait(fullW(sw), static_cast<uint32_t>(gw / W_STAGES) & 1u, dbg, 0x600u | sw);
          tcgen05_fence_after();
          if (j == 0 && lane == 0) stamp(1 + f, 2);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < KG_W; ++kk) {
              const uint64_t a_desc = make_smem_desc_sw128(base + sa * A_BYTES + (kb0 % KG_A + kk) * P::A_KB_BYTES);
              const uint64_t b_desc = make_smem_desc_sw128(base + P::OFF_W + sw * P::W_BYTES + kk * P::W_KB_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_f16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, idesc, (j > 0 || kk > 0 || k > 0) ? 1u : 0u);
            }
            // a slot nobody will refill needs no release (and no arrival may reach a CTA that has already exited)
            if (gw + W_STAGES < total_gw) umma_commit(emptyW(sw));
            if ((kb0 + KG_W) % KG_A == 0 && ga + A_STAGES < total_ga) umma_commit_mc(emptyA(sa), ALL_CTAS);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
        if (lane == 0) stamp(1 + f, 3);
      }
    } else if (warp >= EPI_WARP0) {  // ------------------------------------ epilogue: thread = row
      const int q = warp & 3, tid_e = threadIdx.x - EPI_WARP0 * 32;
      const float* bias = p.fc_params + (static_cast<int64_t>(4) * blk + (second ? 1 : 0)) * D + static_cast<int>(rank) * W_COLS;
      for (int j = tid_e; j < W_COLS; j += 128) bias_s[j] = __ldg(bias + j);
      epi_bar_sync();
      // accumulator rows -> TMEM lanes: M = 128: row r in lane r; M = 64: row r in lane (r % 16) + 32 (r / 16), i.e.
      // every epilogue warp holds ROWS / 4 rows in its first lanes (cute::UMMA tmem_frg_1sm, "half subpartitions")
      constexpr int RPW = ROWS / 4;
      const int64_t row = row0 + q * RPW + lane;
      const bool have_row = lane < RPW && row < p.M;
      bf16* out = (second ? p.u2 : p.u1) + row * D + static_cast<int>(rank) * W_COLS;
      mbar_wait(tfull, static_cast<uint32_t>(f) & 1u, dbg, 0x400u);
      if (tid_e == 0) stamp(1 + f, 4);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < W_COLS / 32; ++cc) {
        uint32_t ra[32];
        tmem_ld_32x32(t_row + cc * 32, ra);
        tmem_ld_wait();
        if (have_row) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              const int e = c * 8 + 2 * hh;
              float x0 = __uint_as_float(ra[e]) + bias_s[cc * 32 + e];
              float x1 = __uint_as_float(ra[e + 1]) + bias_s[cc * 32 + e + 1];
              if (!second) { x0 = silu(x0); x1 = silu(x1); }
              w[hh] = pack_bf16x2(x0, x1);
            }
            st_global_v4(out + cc * 32 + c * 8, w[0], w[1], w[2], w[3]);
          }
        }
      }
      tcgen05_fence_before();
      if (tid_e == 0) stamp(1 + f, 5);
    }
    if (second) {  // R after the block: tail + next modulation, or tail + final modulation + head + Euler
      stage_barrier();  // u2 complete
      if (threadIdx.x == 0) stamp(1 + f, 6);
      const bool last = blk + 1 == depth;
      rp.gate_off = static_cast<int64_t>(3) * blk * D + 2 * D;
      rp.gamma = p.fc_params + (static_cast<int64_t>(4) * blk + 2) * D;
      rp.beta = p.fc_params + (static_cast<int64_t>(4) * blk + 3) * D;
      rp.scale_off = last ? final_off : static_cast<int64_t>(3) * (blk + 1) * D;
      for (int r = wg; r < ROWS; r += CL * (CH_THREADS / 32)) {
        const int64_t row = row0 + r;
        if (row < p.M) {
          if (last) rw::row_body<bf16, VPL, true, 1, true>(rp, row, lane);
          else rw::row_body<bf16, VPL, true, 0, true>(rp, row, lane);
        }
      }
      if (threadIdx.x == 0) stamp(1 + f, 7);
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();  // no CTA exits while a peer may still be reading what it wrote (and TMEM is idle)
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, P::TMEM_COLS_ALLOC);
  }
}

// ---------------------------------------------------------------- host side
template <int VPL, int ROWS>
int set_attrs() {
  static std::atomic<unsigned long long> done{0ull};  // one bit per device
  return ensure_smem_attr(reinterpret_cast<const void*>(chain_kernel<VPL, ROWS>), ChainPlan<VPL, ROWS>::SMEM_BYTES, &done);
}

template <int VPL, int ROWS>
int launch_vpl(const ChainParams& p, const bf16* w_stack, int64_t w_rows, cudaStream_t stream) {
  using P = ChainPlan<VPL, ROWS>;
  NOVA_PROPAGATE((set_attrs<VPL, ROWS>()));
  CUtensorMap th, tu, tw;
  NOVA_PROPAGATE(make_tmap_kmajor(&th, p.h, p.M, p.D, p.D, ROWS));
  NOVA_PROPAGATE(make_tmap_kmajor(&tu, p.u1, p.M, p.D, p.D, ROWS));
  NOVA_PROPAGATE(make_tmap_kmajor(&tw, w_stack, w_rows > 0 ? w_rows : P::W_COLS, p.D, p.D, P::W_COLS));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ceil_div(p.M, ROWS) * CL));
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = P::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  NOVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, chain_kernel<VPL, ROWS>, th, tu, tw, p, debug_word()));
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

// clusters of this kernel the device can hold at once (cached per device); 0 if the query fails
template <int VPL, int ROWS>
int max_clusters_vpl() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    if (set_attrs<VPL, ROWS>() != NOVA_OK) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * 64);
    cfg.blockDim = dim3(CH_THREADS);
    cfg.dynamicSmemBytes = ChainPlan<VPL, ROWS>::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, chain_kernel<VPL, ROWS>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    cached[dev] = n > 0 ? n : -1;
  }
  return cached[dev] > 0 ? cached[dev] : 0;
}

template <int ROWS>
int max_clusters_rows(int D) {
  switch (D / 256) {
    case 1: return max_clusters_vpl<1, ROWS>();
    case 2: return max_clusters_vpl<2, ROWS>();
    case 3: return max_clusters_vpl<3, ROWS>();
    case 4: return max_clusters_vpl<4, ROWS>();
    case 5: return max_clusters_vpl<5, ROWS>();
    case 6: return max_clusters_vpl<6, ROWS>();
    case 7: return max_clusters_vpl<7, ROWS>();
    case 8: return max_clusters_vpl<8, ROWS>();
    default: break;
  }
  return 0;
}

template <int ROWS>
int launch_rows(const ChainParams& p, const bf16* w_stack, cudaStream_t stream) {
  const int64_t w_rows = static_cast<int64_t>(2) * p.depth * p.D;
  switch (p.D / 256) {
    case 1: return launch_vpl<1, ROWS>(p, w_stack, w_rows, stream);
    case 2: return launch_vpl<2, ROWS>(p, w_stack, w_rows, stream);
    case 3: return launch_vpl<3, ROWS>(p, w_stack, w_rows, stream);
    case 4: return launch_vpl<4, ROWS>(p, w_stack, w_rows, stream);
    case 5: return launch_vpl<5, ROWS>(p, w_stack, w_rows, stream);
    case 6: return launch_vpl<6, ROWS>(p, w_stack, w_rows, stream);
    case 7: return launch_vpl<7, ROWS>(p, w_stack, w_rows, stream);
    case 8: return launch_vpl<8, ROWS>(p, w_stack, w_rows, stream);
    default: break;
  }
  set_error("chain kernel: unsupported width %d", p.D);
  return NOVA_ERR_INVALID;
}

int launch_rows64(const ChainParams& p, const bf16* w_stack, cudaStream_t stream);
int max_clusters64(int D);
int launch_rows128(const ChainParams& p, const bf16* w_stack, cudaStream_t stream);

}  // namespace chain
}  // namespace nova
