// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epi(A[M,K] * W[N,K]^T + bias[N])
//
//   A, W : bf16, K-major (row-major with K contiguous) -- nn.Linear's (out,in) weight is
//          already the K-major "B" operand, activations are the K-major "A" operand.
//   acc  : fp32 in tensor memory, double buffered (2 x 256 columns = all 512 TMEM columns)
//   C    : bf16, staged per epilogue warp in swizzled shared memory and written by TMA stores
//          (full 128 B lines, out-of-bounds rows/columns clipped by the tensor map).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0    TMA producer   (one lane): cp.async.bulk.tensor 2D loads, 128B swizzle,
//                            4-stage ring of {A 128x64, W 256x64} tiles, mbarrier tx counts
//   warp 1    MMA issuer     (one lane): tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16,
//                            tcgen05.commit releases smem stages / publishes accumulators
//   warp 2    TMEM allocator (tcgen05.alloc / dealloc)
//   warps 4-7 epilogue       tcgen05.ld 32x32b -> bias (+SiLU) -> bf16 -> st.shared (128B swizzle)
//                            -> cp.async.bulk.tensor store of 32 x 64 sub-tiles, double buffered
//
// This header is compiled by gemm_bias.cu / gemm_silu.cu / gemm_adaln.cu only (each instantiates the kernel
// variants of one epilogue, NOVA_GEMM_TU selects which); other translation units include gemm_api.cuh.
//
// Every mbarrier wait is bounded: on timeout the kernel records where it was stuck in a
// host-mapped debug word and traps, so a protocol bug surfaces as a CUDA error, not a hang.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "gemm_api.cuh"
#include "gemm_simt.cuh"  // Epilogue enum

namespace nova {
namespace tc {

constexpr int UMMA_K = 16;  // BM, BK, BN_FULL: gemm_api.cuh; tile columns BN are a template parameter
constexpr int A_STAGE_BYTES = BM * BK * 2;                // 16 KB
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;
// EPI_TAIL runs 8 epilogue warps (two per TMEM lane quarter, each set taking one half of the tile's columns): its
// epilogue costs ~13 instructions per element.  The epilogue code below is written for either count; for EPI_ADALN
// 8 warps measured 11 % SLOWER (38.6 vs 34.7 ms of AdaLN GEMMs per cfg2 pass), so it stays at 4.
__host__ __device__ constexpr int epi_warps(int epi) { return epi == EPI_TAIL ? 8 : 4; }
// (EPI_TAIL with 11 warps -- TMEM allocator and chunk producer sharing a warp -- was tried to get past the 168 registers
// 384 threads allow: warps are allocated in fours, ptxas stays at 168.)
__host__ __device__ constexpr int epi_warp0(int) { return EPI_WARP0; }
__host__ __device__ constexpr int num_threads(int epi) { return 32 * (epi_warp0(epi) + epi_warps(epi)); }
constexpr int TMEM_COLS = 512;
constexpr int C_CHUNK = 64;                               // output columns per TMA store (128 B rows)
constexpr int C_BUF_BYTES = 32 * C_CHUNK * 2;             // one warp's 32 x 64 bf16 sub-tile: 4 KB

// Per-CTA shared-memory plan for cta_group CG (1: one CTA per 128x256 tile, 2: a CTA pair per
// 256x256 tile, each CTA holding its 128 rows of A and HALF of the W tile).
template <int CG, int BN, bool TAIL = false>
struct Plan {
  static_assert(BN == 64 || BN == 128 || BN == 256, "tile columns: 64 / 128 (small M: more, shorter tiles) or 256");
  static constexpr int B_ROWS = BN / CG;                         // W rows this CTA loads per stage
  static constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;          // BN = 256: 32 KB / 16 KB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // BN = 256: 48 KB / 32 KB
  static constexpr int RING = 192 * 1024 / STAGE_BYTES;          // up to 192 KB of operand ring ...
  // ... in at most 8 stages (4 / 6 at BN = 256).  EPI_TAIL gives 64 KB of the ring to the staged u / x chunks
  // (measured at cfg2: 4 stages instead of 6 for EVERY GEMM change nothing -- 34.1 / 23.8 ms of AdaLN / fc GEMMs per
  // pass against 35.4 / 23.3; a 3-stage ring with 3 staged chunk buffers was 3 % slower than 4 + 2).
  static constexpr int STAGES = TAIL ? (128 * 1024 / STAGE_BYTES) : (RING > 8 ? 8 : RING);
  // EPI_TAIL: T_BUFS buffers x {u chunk, x chunk} of 128 rows x 64 columns (16 KB each), filled by TMA.  The epilogue
  // works IN PLACE: a warp reads u and x of its 32 rows from the buffer, writes the new x over the old one and stores
  // its 32 x 64 slice from there by TMA -- no copy into registers (64 fewer), no separate C staging (32 KB that pay for
  // the third buffer).  A buffer returns to the producer when the four warps of its column half have had their stores
  // read it.
  static constexpr int T_CHUNK_BYTES = BM * C_CHUNK * 2;         // 16 KB
  static constexpr int T_BUF_BYTES = 2 * T_CHUNK_BYTES;          // u + x
  // three: with a 3-stage operand ring four would fit (two per column half), measured 6 % slower on the AdaLN class
  static constexpr int T_BUFS = 3;
  static constexpr int OFF_TAIL = STAGES * STAGE_BYTES;
  static constexpr int OFF_CSTAGE = OFF_TAIL + (TAIL ? T_BUFS * T_BUF_BYTES : 0);  // 4 warps x 2 buffers x 4 KB (not EPI_TAIL)
  static constexpr int OFF_BIAS = OFF_CSTAGE + (TAIL ? 0 : 4 * 2 * C_BUF_BYTES);  // one BN fp32 bias tile
  // EPI_TAIL: + one tile of gamma (fp32) and of beta as bf16 column pairs (exact for a bf16 head: its norm2 parameters
  // ARE bf16; all three tiles in fp32 would exceed the 227 KB by 136 B); with 225 KB of shared memory in use there is
  // next to no L1 left, so per-column constants must not come through it
  static constexpr int OFF_GAMMA = OFF_BIAS + BN_FULL * 4;
  static constexpr int OFF_BETA = OFF_GAMMA + BN_FULL * 4;
  static constexpr int OFF_BAR = OFF_BIAS + BN_FULL * 4 + (TAIL ? BN_FULL * 4 + BN_FULL * 2 : 0);
  // 230 656 B: with the 1 KB the hardware reserves per CTA this leaves >= 1 KB of the SM's 228 KB, so a
  // CTA of a shared-memory-free kernel (the HBM-bound row kernels, launched on a second stream) can be
  // co-resident with a persistent GEMM CTA and its memory traffic overlaps the MMAs.
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int UMMA_M = BM * CG;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory plan exceeds 227 KB");
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: after ~4 s without progress record (code, block) in the host-mapped debug
// words and trap, so a protocol bug is a CUDA error instead of a hung GPU.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, uint32_t* dbg, uint32_t code) {
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > 4000000000ull) {
      if (dbg) {
        dbg[0] = 0xDEAD0000u | code;
        dbg[1] = blockIdx.x;
        dbg[2] = parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* dbg, uint32_t code) {
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, dbg, code);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* t) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(t)) : "memory");
}
// 2D tiled TMA load: coordinates are (c0 = innermost/K element index, c1 = row index).
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* t, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(t)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// 2D tiled TMA store (shared -> global), bulk-group completion; out-of-bounds parts are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* t, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(t)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int THREADS = 128>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {  // issued by the same warp id in both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// ---- CTA-pair (cta_group::2) helpers; encodings follow cute/arch/copy_sm100_tma.hpp and
// cutlass/arch/barrier.h (SM100_TMA_2SM_LOAD_2D, umma_arrive_multicast_2x1SM, ClusterBarrier::arrive)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit: address of the pair's even CTA
// TMA load issued by either CTA of a pair; the transaction bytes land on the LEADER's mbarrier.
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* t, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(t)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs of the pair once prior MMAs retire
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  const uint16_t mask = 0x3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand tile written by TMA with 128B swizzle:
// rows are 128 B apart, 8-row groups 1024 B apart (SBO), LBO unused for swizzled K-major.
// Field layout follows cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// MN-major operand tile (the reduction index runs over the ROWS of the source matrix, as in a weight gradient
// dW = dY^T X read straight from dY [M, N] and X [M, K]): TMA boxes of {64 MN elements = 128 B, 64 reduction rows} with
// 128 B swizzle, one 8 KB box per 64 MN elements.  Canonical form (cute/atom/mma_traits_sm100.hpp, Major::MN, B128), in
// 16-byte units: ((8, n), (8, k)) : ((1, LBO), (8, SBO)) -- 8 reduction rows of 128 B are one 1024 B swizzle atom
// (SBO = 1024 B to the next 8 rows), the next 64 MN elements start LBO bytes on (8192 B: the next box).
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1,
// b=BF16 [10,13)=1, both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
constexpr uint32_t IDESC_A_MN_MAJOR = 1u << 15, IDESC_B_MN_MAJOR = 1u << 16;

struct EpiParams {
  const float* bias;  // [N] or nullptr
  int M, N, K;
  // EPI_ADALN: W rows are packed per 128 features as [128 scale | 128 shift] ("mod" tiles, the first
  // n_mod_tiles column tiles), followed by plain gate tiles.  A mod tile yields
  //   h[:, f0:f0+128] = (x - mean) * rstd * (1 + scale) + shift      (normalization.py:34-36)
  // through tmap_c; gate tiles go through tmap_c2 with bias only.
  const bf16* x;          // [M, ldx] residual stream
  int64_t ldx;
  const float* rowstats;  // [M, 2] = (mean, rstd) of x, eps 1e-6
  int n_mod_tiles;
  // Row-block order: consecutive kernels of a step alternate between ascending and descending row blocks, so a
  // kernel starts on the rows its producer wrote LAST -- the part of a 100 MB activation still in the 126 MB L2.
  int reverse_m;
  // ---- deferred LayerNorm statistics (the block tail fused into the gate GEMM, EPI_TAIL):
  // a row's statistics travel as `parts` partials [parts][M] of (mean_t, M2_t), one per 256-column tile of the
  // producer; the consumer merges them (Chan) -- no kernel ever needs a full row.
  //   EPI_BIAS  : part_out != nullptr: also emit the partials of the (bf16-rounded) outputs      (fc2 -> u statistics)
  //   EPI_ADALN : stats_parts > 0: rowstats holds `stats_parts` partials of x instead of (mean, rstd)
  //   EPI_TAIL  : g = a Wg^T + bg (rounded to bf16 like the stored gate it replaces);
  //               x <- x + (LN(u; part_in, eps 1e-5) * gamma + beta) * g   (diffusion_mlp.py:53), u and x staged by TMA
  //               through tmap_c2 / tmap_c3, x stored through tmap_c; part_out = partials of the new (rounded) x
  const float2* part_in;
  int stats_parts;
  float2* part_out;
  const float* gamma;  // [N] norm2 weight / bias (EPI_TAIL)
  const float* beta;
  // Batched form (weight gradients with the M-reduction split S ways, train_bwd.cu): A is S stacked [batch_m_rows, K]
  // operands, W is S stacked [batch_w_rows, K] operands, and row block m of the stacked output multiplies the W slab
  // of ITS batch.  batch_m_rows = 0: one ordinary GEMM.
  int batch_m_rows, batch_w_rows;
  // MN-major form of the batched GEMM (no transposed copies): C[b] [batch_m_rows, N] = A[b]^T W[b], where A is the source
  // matrix [rows, batch_m_rows] and W the source matrix [rows, N], both row-major, and batch b reduces over source rows
  // [b K, (b + 1) K) (rows beyond the matrices read as zero).  CTA pairs, 256-column tiles.
  int mn_major;
};

// merge `parts` equal-sized partials (mean_t, M2_t over n_t values each) of one row -> (mean, rstd).  Two halves so that
// the loads can be issued at the top of a tile and the arithmetic done after the epilogue warps' barriers.
template <int MAXP>
struct RowPartials {
  float2 v[MAXP];
};
template <int MAXP>
__device__ __forceinline__ void load_partials(const float2* part, int parts, int64_t M, int64_t row, RowPartials<MAXP>& rp) {
#pragma unroll
  for (int t = 0; t < MAXP; ++t)
    if (t < parts) rp.v[t] = part[static_cast<int64_t>(t) * M + row];
}
template <int MAXP>
__device__ __forceinline__ void finish_partials(const RowPartials<MAXP>& rp, int parts, float n_t, float eps, float& mean, float& rstd) {
  float m2 = 0.f, msum = 0.f;
#pragma unroll
  for (int t = 0; t < MAXP; ++t) {
    if (t < parts) {
      msum += rp.v[t].x;
      m2 += rp.v[t].y;
    }
  }
  mean = msum / static_cast<float>(parts);
#pragma unroll
  for (int t = 0; t < MAXP; ++t) {
    if (t < parts) {
      const float d = rp.v[t].x - mean;
      m2 = fmaf(n_t * d, d, m2);
    }
  }
  rstd = rsqrtf(m2 / (n_t * static_cast<float>(parts)) + eps);
}
__device__ __forceinline__ void merge_partials(const float2* part, int parts, int64_t M, int64_t row, float n_t, float eps,
                                               float& mean, float& rstd) {
  RowPartials<16> rp;
  load_partials(part, parts, M, row, rp);
  finish_partials(rp, parts, n_t, eps, mean, rstd);
}

__device__ __forceinline__ uint4 ld_global_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ---- packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issued instruction).
// The epilogues run on 4 or 8 warps next to the MMAs; what they cost is issue slots and dependent latency, so every
// per-element operation is done on (column e, column e + 1) pairs held in 64-bit registers.
typedef uint64_t f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t lo, uint32_t hi) {  // two accumulator words as loaded from TMEM
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t w) { return pk2u(w << 16, w & 0xFFFF0000u); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ f32x2 lds_f32x2(const float* p) { return *reinterpret_cast<const f32x2*>(p); }

template <int EPI, int CG, int BN>
__global__ void __launch_bounds__(num_threads(EPI), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2,
            const __grid_constant__ CUtensorMap tmap_c3, const EpiParams p, uint32_t* dbg) {
  using P = Plan<CG, BN, EPI == EPI_TAIL>;
  constexpr int EW0 = epi_warp0(EPI);  // first epilogue warp
  static_assert(EPI != EPI_ADALN || BN == BN_FULL, "the AdaLN epilogue pairs 128 scale + 128 shift columns per tile");
  static_assert(EPI != EPI_TAIL || BN == BN_FULL, "the tail epilogue works on 256-column tiles");
  constexpr int STAGES = P::STAGES, STAGE_BYTES = P::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw;
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) {
    if (dbg && threadIdx.x == 0) { dbg[0] = 0xDEAD0A11u; dbg[1] = base; __threadfence_system(); }
    __trap();
  }
  const uint32_t bar_base = base + P::OFF_BAR;
  // barrier block: full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P::OFF_BAR + 8 * (2 * STAGES + 4));
  // EPI_TAIL: full / empty barriers of the two staged {u, x} chunk buffers
  auto tail_full = [&](int b) { return bar_base + 8u * (2 * STAGES + 5 + b); };   // b < T_BUFS
  auto tail_empty = [&](int b) { return bar_base + 8u * (2 * STAGES + 9 + b); };
  // EPI_ADALN: epilogue warp q stages its 32 x 64 chunks of x through its two C staging buffers (TMA load, in-place
  // modulation, TMA store): one "landed" barrier per (warp, buffer)
  auto xfull = [&](int qq, int b) { return bar_base + 8u * (2 * STAGES + 5 + 2 * qq + b); };  // (never together with the tail barriers)

  pdl_trigger();  // the next kernel may be scheduled as soon as resources free up; it waits for our completion itself
#ifdef NOVA_GEMM_TIMELINE  // diagnostic build (scripts/profile_gemm_timeline.py): SM-clock stamps of CTA 0 in the debug words
  const long long tl_t0 = clock64();
#define NOVA_TL_STAMP(word) do { if (dbg && blockIdx.x == 0) { dbg[word] = (uint32_t)(clock64() - tl_t0); } } while (0)
#else
#define NOVA_TL_STAMP(word) do { } while (0)
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader of the CTA pair
  const int group = blockIdx.x / CG, num_groups = gridDim.x / CG;
  const int num_m = (p.M + BM * CG - 1) / (BM * CG), num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    if (EPI == EPI_ADALN || EPI == EPI_TAIL) {
      prefetch_tmap(&tmap_c2);
      prefetch_tmap(&tmap_c3);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // the leader's arrive.expect_tx; TMA bytes of the whole pair
      mbar_init(empty_bar(s), 1);  // one tcgen05.commit arrival (multicast to both CTAs when CG == 2)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), CG * epi_warps(EPI) * 32);  // every epilogue thread of every CTA of the group arrives
      if (EPI == EPI_TAIL) {
        for (int t = b; t < P::T_BUFS; t += 2) {
          mbar_init(tail_full(t), 1);   // the tail producer's arrive.expect_tx
          mbar_init(tail_empty(t), 4);  // one arrival per warp of the column half once its store has read the buffer
        }
      }
      if (EPI == EPI_ADALN)
        for (int qq = 0; qq < 4; ++qq) mbar_init(xfull(qq, b), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();  // peer barriers are initialised before any remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barriers, TMEM, tensor-map prefetch) is independent of the preceding kernel and overlaps
  // its tail; from here on global memory written by it is read (A via TMA, x / rowstats) or overwritten (C).
  pdl_wait();
  if (threadIdx.x == 0) NOVA_TL_STAMP(0);  // prologue done and the preceding kernel complete

  if (warp == 0) {
    if (lane == 0) {  // ---------------------------------------------- TMA producer (both CTAs of a pair)
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups) {
        const int m_tile = p.reverse_m ? num_m - 1 - tile / num_n : tile / num_n;
        const int m_idx = m_tile * (BM * CG) + static_cast<int>(rank) * BM;
        const int w_batch = p.batch_m_rows > 0 ? (m_tile * (BM * CG) / p.batch_m_rows) * p.batch_w_rows : 0;
        const int n_idx = (tile % num_n) * BN + static_cast<int>(rank) * P::B_ROWS + w_batch;  // CG == 2: my half of W
        // MN-major batched form: batch b = m_tile / (tiles per batch); operand columns instead of operand rows
        const int mn_tiles = (CG == 2 && p.mn_major) ? p.batch_m_rows / (BM * CG) : 1;
        const int mn_b = m_tile / mn_tiles;
        const int mn_a_col = (m_tile - mn_b * mn_tiles) * (BM * CG) + static_cast<int>(rank) * BM;
        const int mn_w_col = (tile % num_n) * BN + static_cast<int>(rank) * P::B_ROWS;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, dbg, 0x100u | stage);
          const uint32_t sa = base + stage * STAGE_BYTES;
          if (CG == 2 && p.mn_major) {
            const int row = mn_b * p.K + kb * BK;  // source rows of this k-block
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
#pragma unroll
            for (int at = 0; at < BM / 64; ++at)  // one 64 x 64 box (8 KB) per 64 MN elements
              tma_load_2d_2sm(&tmap_a, full_bar(stage), sa + at * 8192, mn_a_col + at * 64, row);
#pragma unroll
            for (int at = 0; at < P::B_ROWS / 64; ++at)
              tma_load_2d_2sm(&tmap_b, full_bar(stage), sa + A_STAGE_BYTES + at * 8192, mn_w_col + at * 64, row);
          } else if (CG == 1) {
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(&tmap_a, full_bar(stage), sa, kb * BK, m_idx);
            tma_load_2d(&tmap_b, full_bar(stage), sa + A_STAGE_BYTES, kb * BK, n_idx);
          } else {
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            tma_load_2d_2sm(&tmap_a, full_bar(stage), sa, kb * BK, m_idx);
            tma_load_2d_2sm(&tmap_b, full_bar(stage), sa + A_STAGE_BYTES, kb * BK, n_idx);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // --------------------------------- MMA issuer (leader CTA only)
      constexpr uint32_t idesc = make_idesc_bf16(P::UMMA_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t buf_phase = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups) {
        mbar_wait(tempty_bar(buf), buf_phase ^ 1u, dbg, 0x200u | buf);  // epilogues drained this buffer
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar(stage), phase, dbg, 0x300u | stage);  // TMA bytes landed (both CTAs)
          if (tile == group && kb == 0) NOVA_TL_STAMP(1);  // first operands landed
          tcgen05_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
          if (CG == 2 && p.mn_major) {
            const uint64_t a_desc = make_smem_desc_mn_sw128(sa, 8192u);
            const uint64_t b_desc = make_smem_desc_mn_sw128(sa + A_STAGE_BYTES, 8192u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // 16 reduction rows = two 1024 B swizzle atoms: +2048 B = +128 in 16 B units
              const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
              umma_f16_2sm(d_tmem, a_desc + 128u * k, b_desc + 128u * k, idesc | IDESC_A_MN_MAJOR | IDESC_B_MN_MAJOR, acc);
            }
          } else {
          const uint64_t a_desc = make_smem_desc_sw128(sa);
          const uint64_t b_desc = make_smem_desc_sw128(sa + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in 16 B units
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if (CG == 1) umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, acc);
            else umma_f16_2sm(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, acc);
          }
          }
          if (CG == 1) umma_commit(empty_bar(stage)); else umma_commit_2sm(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (CG == 1) umma_commit(tfull_bar(buf)); else umma_commit_2sm(tfull_bar(buf));  // accumulator complete
        if (++buf == 2) { buf = 0; buf_phase ^= 1u; }
      }
    }
  } else if (EPI == EPI_TAIL && warp == 3) {
    if (lane == 0) {  // ------------------------------------------------ tail producer: {u, x} chunks of my 128 rows
      int g = 0;  // running chunk index: chunk g lives in buffer g % T_BUFS
      int it = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups, ++it) {
        const int m_tile = p.reverse_m ? num_m - 1 - tile / num_n : tile / num_n;
        const int m_idx = m_tile * (BM * CG) + static_cast<int>(rank) * BM, n_idx = (tile % num_n) * BN;
        // consecutive chunks go to alternating column halves (left half: tile chunks 0, 1; right half: 2, 3), and the
        // half that is served first alternates from tile to tile, so neither half is always the one whose first chunk
        // had the shorter head start
        for (int k = 0; k < BN / C_CHUNK; ++k, ++g) {
          const int cc = (((k ^ it) & 1) * 2) + (k >> 1);
          const int tb = g % P::T_BUFS;
          mbar_wait(tail_empty(tb), ((static_cast<uint32_t>(g / P::T_BUFS) & 1u) ^ 1u), dbg, 0x700u | tb);
          const uint32_t dst = base + P::OFF_TAIL + static_cast<uint32_t>(tb) * P::T_BUF_BYTES;
          mbar_expect_tx(tail_full(tb), P::T_BUF_BYTES);
          tma_load_2d(&tmap_c2, tail_full(tb), dst, n_idx + cc * C_CHUNK, m_idx);                     // u chunk
          tma_load_2d(&tmap_c3, tail_full(tb), dst + P::T_CHUNK_BYTES, n_idx + cc * C_CHUNK, m_idx);  // x chunk
        }
      }
    }
  } else if (warp >= EW0) {  // ------------------------------------ epilogue (every CTA: its 128 rows)
    const int q = warp & 3;  // TMEM lane quarter this warp may access == its 32-row slice of the tile
    const int tid_e = threadIdx.x - EW0 * 32;
    // C staging: 4 warps x 2 buffers of 4 KB; EPI_TAIL: 8 warps x 1 buffer
    constexpr bool WIDE_EPI = epi_warps(EPI) == 8;
    const int half = WIDE_EPI ? (warp - EW0) >> 2 : 0;  // which half of the tile's columns this warp set takes
    const uint32_t cbuf = base + P::OFF_CSTAGE + (WIDE_EPI ? static_cast<uint32_t>(warp - EW0) * C_BUF_BYTES
                                                           : static_cast<uint32_t>(q) * 2u * C_BUF_BYTES);
    float* bias_all = reinterpret_cast<float*>(smem + P::OFF_BIAS);
    int buf = 0, cpar = 0;
    uint32_t buf_phase = 0;
    int tile_it = 0;      // EPI_TAIL: tiles done so far (chunk indices follow the producer's order)
    int prev_tb = -1;     // EPI_TAIL: staged buffer my last store still reads, -1 if none
    uint32_t xphase = 0u;  // EPI_ADALN: parities of my two x-staging barriers (bit b = buffer b)
#ifdef NOVA_TAIL_TIMELINE  // diagnostic build: where an EPI_TAIL epilogue warp spends its cycles (CTA 0, first warp)
    long long tt_full = 0, tt_acc = 0, tt_store = 0, tt_bar = 0;
    const long long tt_begin = clock64();
#define NOVA_TT(acc, stmt) do { const long long _t = clock64(); stmt; acc += clock64() - _t; } while (0)
#else
#define NOVA_TT(acc, stmt) do { stmt; } while (0)
#endif
    // What an epilogue needs from global memory per tile -- its slice of the bias (gamma, beta) tile and its row's
    // LayerNorm statistics -- is fetched ONE TILE AHEAD into registers: the round trips run under the previous tile's
    // arithmetic instead of in front of this tile's barriers (ncu: a quarter of the tail epilogue's samples sat there).
    using RP = RowPartials<EPI == EPI_TAIL ? 8 : 16>;  // the tail merges <= 8 partials of u (launch_tail), the modulation <= 16 of x
    constexpr int NB = BN / (32 * epi_warps(EPI)) > 0 ? BN / (32 * epi_warps(EPI)) : 1;  // bias values per thread
    RP rp_n;
    float2 st2_n = make_float2(0.f, 0.f);
    float bv_n[NB] = {}, gv_n[NB] = {}, btv_n[NB] = {};
    // fetch() only LOADS (addresses clamped into range instead of predicated): the masking of out-of-range columns and
    // the (1 + bias) of the scale columns happen where the values are consumed, a tile later -- an add or a select right
    // behind its load made the warp sit out the whole L2 round trip inside fetch() (ncu source view, round 2: 172 + 33
    // samples on the two FADDs behind the modulation epilogue's bias loads, 95 in the tail's).  The constants of a column
    // tile are skipped when the next tile is in the same column (fixed-column grids, launch_epi).
    auto fetch = [&](int t, int have_n) {
      const int t_n = t % num_n;
      const int t_m0 = (p.reverse_m ? num_m - 1 - t / num_n : t / num_n) * (BM * CG) + static_cast<int>(rank) * BM + q * 32;
      const bool t_mod = EPI == EPI_ADALN && t_n < p.n_mod_tiles;
      if ((t_mod || EPI == EPI_TAIL) && t_m0 + lane < p.M) {
        if (EPI == EPI_TAIL || p.stats_parts > 0) load_partials(p.part_in, p.stats_parts, p.M, t_m0 + lane, rp_n);
        else st2_n = *reinterpret_cast<const float2*>(p.rowstats + 2 * static_cast<int64_t>(t_m0 + lane));
      }
      if (t_n == have_n) return;
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int j = tid_e + i * 32 * epi_warps(EPI);
        const int col = t_n * BN + j < p.N ? t_n * BN + j : p.N - 1;
        if (p.bias != nullptr) bv_n[i] = __ldg(p.bias + col);
        if (EPI == EPI_TAIL) {
          gv_n[i] = __ldg(p.gamma + col);
          btv_n[i] = __ldg(p.beta + col);
        }
      }
    };
    int smem_n = -1;  // column tile whose bias (gamma, beta) tile is in shared memory
    if (group < num_tiles) fetch(group, -1);
    for (int tile = group; tile < num_tiles; tile += num_groups) {
      const int m_tile = p.reverse_m ? num_m - 1 - tile / num_n : tile / num_n;
      const int m_idx = m_tile * (BM * CG) + static_cast<int>(rank) * BM, n_idx = (tile % num_n) * BN;
      float* bias_s = bias_all;
      const int tile_n = tile % num_n;
      const int m0 = m_idx + q * 32;
      const bool mod_tile = EPI == EPI_ADALN && tile_n < p.n_mod_tiles;
      // ---- everything this tile's epilogue needs from memory is requested FIRST; the barriers below then wait under
      // those round trips instead of in front of them
      constexpr bool X_STAGED = EPI == EPI_ADALN && !WIDE_EPI;  // x chunks through the C staging buffers (TMA)
      constexpr int KPW = WIDE_EPI ? 1 : 2;  // 64-feature groups of a modulation tile per warp (set)
      const int cpar0 = cpar;                // staging buffer of this tile's first chunk
      if (X_STAGED && mod_tile && m0 < p.M) {
        if (lane == 0) {
          tma_store_wait_read<0>();  // both staging buffers: their last stores have read them
#pragma unroll
          for (int kk = 0; kk < KPW; ++kk) {
            const uint32_t cb = cbuf + static_cast<uint32_t>(cpar0 ^ kk) * C_BUF_BYTES;
            mbar_expect_tx(xfull(q, cpar0 ^ kk), C_BUF_BYTES);
            tma_load_2d(&tmap_c3, xfull(q, cpar0 ^ kk), cb, tile_n * 128 + 64 * kk, m0);
          }
        }
        __syncwarp();
      }
      const RP rp = rp_n;  // this tile's values, requested a tile ago
      const float2 st2 = st2_n;
      float bv[NB], gv[NB], btv[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        bv[i] = bv_n[i];
        gv[i] = EPI == EPI_TAIL ? gv_n[i] : 0.f;
        btv[i] = EPI == EPI_TAIL ? btv_n[i] : 0.f;
      }
      if (tile + num_groups < num_tiles) fetch(tile + num_groups, tile_n);
      if (EPI == EPI_TAIL && prev_tb >= 0) {
        // the staged buffer of the previous tile's last chunk goes back to the producer (its store has had the time of
        // the loads above to read it); waiting until my next chunk would hold up the other half's next chunk
        if (lane == 0) { tma_store_wait_read<0>(); mbar_arrive(tail_empty(prev_tb)); }
        prev_tb = -1;
      }
      if (tile_n != smem_n) {  // same for every epilogue warp: the two barriers stay matched
        smem_n = tile_n;
        NOVA_TT(tt_bar, epi_bar_sync<32 * epi_warps(EPI)>());  // every epilogue warp has finished reading the previous tile's bias
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          const int j = tid_e + i * 32 * epi_warps(EPI);
          if (j < BN) {
            const bool in = tile_n * BN + j < p.N;
            // modulation tiles: the scale half carries (1 + bias), so the epilogue forms 1 + scale with one addition
            bias_s[j] = ((p.bias != nullptr && in) ? bv[i] : 0.f) + ((mod_tile && j < 128) ? 1.0f : 0.f);
            if (EPI == EPI_TAIL) {
              reinterpret_cast<float*>(smem + P::OFF_GAMMA)[j] = in ? gv[i] : 0.f;
              reinterpret_cast<bf16*>(smem + P::OFF_BETA)[j] = __float2bfloat16_rn(in ? btv[i] : 0.f);  // column pairs: (beta_e, beta_e+1)
            }
          }
        }
        NOVA_TT(tt_bar, epi_bar_sync<32 * epi_warps(EPI)>());  // bias tile visible to the epilogue warps
      }
      // AdaLN modulation tile: this thread's row statistics (and, when x is not staged, its 128 features of x) do not
      // depend on the MMA, so they are requested BEFORE waiting for the accumulator and land while the tile is computed.
      uint4 xv[(EPI == EPI_ADALN && !X_STAGED) ? 8 * KPW : 1];  // my features of x: [64 half KPW, .. + 64 KPW) of the tile's 128
      float mean = 0.f, rstd = 0.f;
      if (mod_tile) {
        const int row = m0 + lane;
        const bool valid = row < p.M;
        if (!X_STAGED) {
          const bf16* xrow = p.x + static_cast<int64_t>(valid ? row : 0) * p.ldx + tile_n * 128 + half * 64 * KPW;
#pragma unroll
          for (int c = 0; c < ((EPI == EPI_ADALN && !X_STAGED) ? 8 * KPW : 1); ++c)
            xv[c] = p.x != nullptr ? ld_global_nc_v4(xrow + 8 * c) : make_uint4(0u, 0u, 0u, 0u);
        }
        if (valid) {
          if (p.stats_parts > 0) {  // x was produced by an EPI_TAIL epilogue: merge its per-tile partials
            finish_partials(rp, p.stats_parts, static_cast<float>(p.n_mod_tiles * 128 / p.stats_parts), 1e-6f, mean, rstd);
          } else {
            mean = st2.x;
            rstd = st2.y;
          }
        }
      }
      NOVA_TT(tt_acc, mbar_wait(tfull_bar(buf), buf_phase, dbg, 0x400u | buf));
      if (tile == group && threadIdx.x == EW0 * 32) NOVA_TL_STAMP(2);  // first accumulator complete
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      if (mod_tile) {
        // ---- TMEM columns [0,128) = scale, [128,256) = shift of features f0..f0+127
        const int f0 = tile_n * 128;
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) {  // 64 features -> one staging buffer -> one TMA store
          const int k = half * KPW + kk;
          if (m0 < p.M) {
            if (X_STAGED) {  // my 32 x 64 chunk of x has landed in this staging buffer; h overwrites it in place
              mbar_wait(xfull(q, cpar), (xphase >> cpar) & 1u, dbg, 0x900u | (q << 1) | cpar);
              xphase ^= 1u << cpar;
            } else {
              if (lane == 0) {  // the store that last used this staging buffer has read it
                if (WIDE_EPI) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
              }
              __syncwarp();
            }
            const uint32_t cb = cbuf + static_cast<uint32_t>(WIDE_EPI ? 0 : cpar) * C_BUF_BYTES;
            const uint32_t dst = cb + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
            for (int sc = 0; sc < 2; ++sc) {  // 32-feature sub-chunks keep the register footprint bounded
              const int fo = 64 * k + 32 * sc;  // feature offset inside the tile
              uint32_t rs[32], rh[32];
              tmem_ld_32x32(t_row + fo, rs);
              tmem_ld_32x32(t_row + 128 + fo, rh);
              tmem_ld_wait();
              const float* bsc = bias_s + fo;        // 1 + bias of the scale columns
              const float* bsh = bias_s + 128 + fo;
              const f32x2 nmean2 = pk2(-mean, -mean), rstd2 = pk2(rstd, rstd);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint4 xq;
                if (X_STAGED) {
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(xq.x), "=r"(xq.y), "=r"(xq.z), "=r"(xq.w)
                               : "r"(dst + (static_cast<uint32_t>((sc * 4 + c) ^ (lane & 7)) << 4)));
                } else {
                  xq = xv[(EPI == EPI_ADALN && !X_STAGED) ? (kk * 2 + sc) * 4 + c : 0];
                }
                const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w};
                uint32_t w[4];
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {  // column pairs (e, e + 1): 5 packed fp32 instructions per pair
                  const int e = c * 8 + 2 * hh;
                  const f32x2 s2 = add2(pk2u(rs[e], rs[e + 1]), lds_f32x2(bsc + e));
                  const f32x2 h2 = add2(pk2u(rh[e], rh[e + 1]), lds_f32x2(bsh + e));
                  const f32x2 xn2 = mul2(add2(bf16x2_to_f32x2(xw[hh]), nmean2), rstd2);
                  w[hh] = pack_bf16x2(fma2(xn2, s2, h2));
                }
                st_shared_v4(dst + (static_cast<uint32_t>((sc * 4 + c) ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_c, cb, f0 + 64 * k, m0);
              tma_store_commit();
            }
            cpar ^= 1;
          }
        }
      } else if (EPI == EPI_TAIL) {
        // ---- gate tile with the block tail fused in: x <- x + (LN(u) * gamma + beta) * g   (diffusion_mlp.py:53)
        // warps 4-7 take columns [0, 128) of the tile (chunks 0, 1; staged buffer 0), warps 8-11 columns [128, 256)
        const int row = m0 + lane;
        const bool valid = row < p.M;
        float mu = 0.f, ru = 0.f;
        if (valid) finish_partials(rp, p.stats_parts, static_cast<float>(p.N / p.stats_parts), 1e-5f, mu, ru);
        // statistics of the new x over my half tile, about its first value (packed: even / odd columns)
        float c0 = 0.f;
        f32x2 nc02 = pk2(0.f, 0.f), s1v = pk2(0.f, 0.f), s2v = pk2(0.f, 0.f);
        const f32x2 nmu2 = pk2(-mu, -mu), ru2 = pk2(ru, ru);
        const float* gam_s = reinterpret_cast<const float*>(smem + P::OFF_GAMMA);
        const uint32_t* bet_s = reinterpret_cast<const uint32_t*>(smem + P::OFF_BETA);
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
          const int cc = half * 2 + k;
          const int n0 = n_idx + cc * C_CHUNK;
          // chunk index in the producer's order (0, 2, 1, 3 per tile: halves alternate) -> buffer and barrier phase
          const int g = tile_it * (BN / C_CHUNK) + ((half ^ tile_it) & 1) + 2 * k;
          const int tb = g % P::T_BUFS;
          // the buffer of my previous chunk goes back to the producer once my store has read it
          if (prev_tb >= 0) {
            NOVA_TT(tt_store, { if (lane == 0) { tma_store_wait_read<0>(); mbar_arrive(tail_empty(prev_tb)); } });
            prev_tb = -1;
          }
          NOVA_TT(tt_full, mbar_wait(tail_full(tb), static_cast<uint32_t>(g / P::T_BUFS) & 1u, dbg, 0x800u | tb));
          prev_tb = tb;
          if (m0 >= p.M || n0 >= p.N) continue;  // warp-uniform: nothing of this sub-tile is in bounds
          // my row of the staged chunk: u at ub, x (overwritten in place by the new x) at xb; 128 B rows, 128 B swizzle
          const uint32_t ub = base + P::OFF_TAIL + static_cast<uint32_t>(tb) * P::T_BUF_BYTES +
                              static_cast<uint32_t>(q * 32 + lane) * 128u;
          const uint32_t xb = ub + P::T_CHUNK_BYTES;
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {  // 32 accumulator columns at a time keeps the register footprint bounded
            uint32_t ra[32];
            tmem_ld_32x32(t_row + cc * C_CHUNK + hc * 32, ra);
            tmem_ld_wait();
            const int col0 = cc * C_CHUNK + hc * 32;  // first tile column of these 32
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int c = hc * 4 + c4;
              const uint32_t off = static_cast<uint32_t>(c ^ (lane & 7)) << 4;
              uint32_t uw[4], xw[4];
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(uw[0]), "=r"(uw[1]), "=r"(uw[2]), "=r"(uw[3]) : "r"(ub + off));
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(xw[0]), "=r"(xw[1]), "=r"(xw[2]), "=r"(xw[3]) : "r"(xb + off));
              const uint4 bq = *reinterpret_cast<const uint4*>(bet_s + (col0 + c4 * 8) / 2);  // beta of 8 columns
              const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
              uint32_t w[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                // column pairs (e, e + 1), 8 packed fp32 instructions per pair; the gate stays in fp32 (it is no longer a
                // stored bf16 tensor) and the statistics take the values before rounding
                const int e = c4 * 8 + 2 * h;
                const f32x2 g2 = add2(pk2u(ra[e], ra[e + 1]), lds_f32x2(bias_s + col0 + e));
                const f32x2 t2 = mul2(add2(bf16x2_to_f32x2(uw[h]), nmu2), ru2);
                const f32x2 l2 = fma2(t2, lds_f32x2(gam_s + col0 + e), bf16x2_to_f32x2(bw[h]));
                const f32x2 y2 = fma2(l2, g2, bf16x2_to_f32x2(xw[h]));
                w[h] = pack_bf16x2(y2);
                if (k == 0 && c == 0 && h == 0) {
                  float y1;
                  upk2(y2, c0, y1);
                  nc02 = pk2(-c0, -c0);
                }
                const f32x2 d2 = add2(y2, nc02);
                s1v = add2(s1v, d2);
                s2v = fma2(d2, d2, s2v);
              }
              st_shared_v4(xb + off, w[0], w[1], w[2], w[3]);  // over the x it was computed from
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {  // my 32 x 64 slice of the buffer's x chunk
            tma_store_2d(&tmap_c, base + P::OFF_TAIL + static_cast<uint32_t>(tb) * P::T_BUF_BYTES + P::T_CHUNK_BYTES +
                                      static_cast<uint32_t>(q) * C_BUF_BYTES, n0, m0);
            tma_store_commit();
          }
        }
        float s1, s2;
        {
          float a0, a1, b0, b1;
          upk2(s1v, a0, a1);
          upk2(s2v, b0, b1);
          s1 = a0 + a1;
          s2 = b0 + b1;
        }
        if (valid && p.part_out != nullptr) {  // one partial per 128-column half tile
          const float nt = static_cast<float>(BN / 2);
          p.part_out[static_cast<int64_t>(tile_n * 2 + half) * p.M + row] = make_float2(c0 + s1 / nt, s2 - s1 * s1 / nt);
        }
      } else {
        // ---- plain tile: bias (+SiLU); for EPI_ADALN these are the gate tiles, written through tmap_c2
        const CUtensorMap* out_map = EPI == EPI_ADALN ? &tmap_c2 : &tmap_c;
        const int out_n = EPI == EPI_ADALN ? (tile_n - p.n_mod_tiles) * BN : n_idx;
        const int out_cols = EPI == EPI_ADALN ? p.N - p.n_mod_tiles * BN : p.N;
        const bool want_parts = EPI == EPI_BIAS && p.part_out != nullptr;  // fc2: partial LayerNorm statistics of u
        // EPI_BIAS_SILU_DUAL (training forward): the pre-activation a chunk's SiLU is taken of goes out as well (through
        // tmap_c2; the backward pass needs it), so a chunk fills BOTH staging buffers: 0 = pre-activation, 1 = activation
        constexpr bool SILU = EPI == EPI_BIAS_SILU || EPI == EPI_BIAS_SILU_DUAL;
        constexpr bool DUAL = EPI == EPI_BIAS_SILU_DUAL;
        float c0 = 0.f;
        f32x2 nc02 = pk2(0.f, 0.f), s1v = pk2(0.f, 0.f), s2v = pk2(0.f, 0.f);
        constexpr int CPW = (BN / C_CHUNK) / (WIDE_EPI ? 2 : 1);  // chunks per warp set
#pragma unroll 1
        for (int cc = half * CPW; cc < (half + 1) * CPW; ++cc) {
          const int n0 = out_n + cc * C_CHUNK;
          if (m0 >= p.M || n0 >= out_cols) break;  // warp-uniform: nothing of this sub-tile is in bounds
          uint32_t ra[32], rb[32];
          tmem_ld_32x32(t_row + cc * C_CHUNK, ra);
          tmem_ld_32x32(t_row + cc * C_CHUNK + 32, rb);
          if (lane == 0) {  // the store that last used this staging buffer has read it
            if (WIDE_EPI || DUAL) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
          }
          __syncwarp();
          tmem_ld_wait();
          const uint32_t dst = cbuf + static_cast<uint32_t>((WIDE_EPI || DUAL) ? 0 : cpar) * C_BUF_BYTES + static_cast<uint32_t>(lane) * 128u;
          const float* bs = bias_s + cc * C_CHUNK;
#pragma unroll
          for (int c = 0; c < 8; ++c) {  // 8 x 16 B chunks of this thread's 128 B row, XOR-swizzled
            uint32_t w[4];
            uint32_t wp[DUAL ? 4 : 1];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int e = c * 8 + 2 * h;  // compile-time after unrolling: picks ra or rb statically
              f32x2 x2 = add2(e < 32 ? pk2u(ra[e & 31], ra[(e + 1) & 31]) : pk2u(rb[e & 31], rb[(e + 1) & 31]), lds_f32x2(bs + e));
              if (DUAL) wp[h] = pack_bf16x2(x2);
              if (SILU) {  // silu(x) = h + h tanh(h), h = x / 2 (common.cuh), on the pair
                const f32x2 h2 = mul2(x2, pk2(0.5f, 0.5f));
                float h0, h1, t0, t1;
                upk2(h2, h0, h1);
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                x2 = fma2(h2, pk2(t0, t1), h2);
              }
              w[h] = pack_bf16x2(x2);
              if (want_parts) {  // statistics of what is stored (rounded), about the tile's first value
                const f32x2 y2 = bf16x2_to_f32x2(w[h]);
                if (cc == 0 && c == 0 && h == 0) {
                  c0 = bf16_lo(w[h]);
                  nc02 = pk2(-c0, -c0);
                }
                const f32x2 d2 = add2(y2, nc02);
                s1v = add2(s1v, d2);
                s2v = fma2(d2, d2, s2v);
              }
            }
            if (DUAL) {
              st_shared_v4(dst + (static_cast<uint32_t>(c ^ (lane & 7)) << 4), wp[0], wp[1], wp[2], wp[3]);
              st_shared_v4(dst + C_BUF_BYTES + (static_cast<uint32_t>(c ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
            } else {
              st_shared_v4(dst + (static_cast<uint32_t>(c ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
            }
          }
          fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            if (DUAL) {
              tma_store_2d(&tmap_c2, cbuf, n0, m0);               // pre-activation
              tma_store_2d(&tmap_c, cbuf + C_BUF_BYTES, n0, m0);  // activation
            } else {
              tma_store_2d(out_map, cbuf + static_cast<uint32_t>(WIDE_EPI ? 0 : cpar) * C_BUF_BYTES, n0, m0);
            }
            tma_store_commit();
          }
          cpar ^= 1;
        }
        if (want_parts && m0 + lane < p.M) {
          float a0, a1, b0, b1;
          upk2(s1v, a0, a1);
          upk2(s2v, b0, b1);
          const float s1 = a0 + a1, s2 = b0 + b1;
          const float nt = static_cast<float>(BN);
          p.part_out[static_cast<int64_t>(tile_n) * p.M + (m0 + lane)] = make_float2(c0 + s1 / nt, s2 - s1 * s1 / nt);
        }
      }
      tcgen05_fence_before();
      // all tcgen05.ld of this accumulator have completed: hand the buffer back to the MMA issuer
      if (CG == 1) mbar_arrive(tempty_bar(buf)); else mbar_arrive_cluster(tempty_bar(buf), 0u);
      if (++buf == 2) { buf = 0; buf_phase ^= 1u; }
      ++tile_it;
    }
    if (lane == 0) tma_store_wait_read<0>();  // staging memory must outlive the last stores' reads
    if (threadIdx.x == EW0 * 32) NOVA_TL_STAMP(3);  // epilogue done
#ifdef NOVA_TAIL_TIMELINE
    if (dbg && blockIdx.x == 0 && threadIdx.x == EW0 * 32) {
      dbg[0] = static_cast<uint32_t>(clock64() - tt_begin);  // whole epilogue loop
      dbg[1] = static_cast<uint32_t>(tt_full);                // waiting for the staged {u, x} chunks
      dbg[2] = static_cast<uint32_t>(tt_acc);                 // waiting for accumulators
      // waiting for my staging buffer's last store (low 16 bits) | at the epilogue warps' barriers (high 16), units of 64 cycles
      dbg[3] = static_cast<uint32_t>((tt_store >> 6) & 0xffff) | (static_cast<uint32_t>((tt_bar >> 6) & 0xffff) << 16);
      __threadfence_system();
    }
#endif
  }
  __syncwarp();  // lanes of the single-lane roles reconverge before the CTA-wide barrier
  tcgen05_fence_before();
  // CG == 2: neither CTA may exit (or free TMEM) while the pair's MMAs can still touch its smem / TMEM
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side (declarations: gemm_api.cuh)
template <int EPI, int CG, int BN = BN_FULL>
int launch_epi(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc,
               int M, int N, int K, cudaStream_t stream, const AdaLNArgs* ada = nullptr, bool reverse_m = false,
               float2* part_out = nullptr, const TailArgs* tail = nullptr, int batch_m_rows = 0, int batch_w_rows = 0,
               int mn_rows = 0, bf16* pre_out = nullptr, int64_t ldpre = 0) {
  using P = Plan<CG, BN, EPI == EPI_TAIL>;
  static std::atomic<unsigned long long> attr_done{0ull};  // one bit per device
  NOVA_PROPAGATE(ensure_smem_attr(reinterpret_cast<const void*>(gemm_kernel<EPI, CG, BN>), P::SMEM_BYTES, &attr_done));
  CUtensorMap ta, tb, tc_, tc2, tc3;
  if (mn_rows > 0) {
    // MN-major batched form: A = source [mn_rows, batch_m_rows] (row stride lda), W = source [mn_rows, N] (row stride ldw);
    // boxes of {64 columns, 64 source rows}
    NOVA_PROPAGATE(make_tmap_kmajor(&ta, A, mn_rows, batch_m_rows, lda, 64));
    NOVA_PROPAGATE(make_tmap_kmajor(&tb, W, mn_rows, N, ldw, 64));
  } else {
    NOVA_PROPAGATE(make_tmap_kmajor(&ta, A, M, K, lda, BM));
    // batched: W holds one [batch_w_rows, K] slab per batch (M / batch_m_rows of them)
    NOVA_PROPAGATE(make_tmap_kmajor(&tb, W, batch_m_rows > 0 ? static_cast<int64_t>(M / batch_m_rows) * batch_w_rows : N, K, ldw,
                                    P::B_ROWS));
  }
  EpiParams p{};
  p.batch_m_rows = batch_m_rows; p.batch_w_rows = batch_w_rows; p.mn_major = mn_rows > 0 ? 1 : 0;
  p.bias = bias; p.M = M; p.N = N; p.K = K; p.reverse_m = reverse_m ? 1 : 0;
  p.part_out = part_out;
  if (EPI == EPI_ADALN) {
    // C = h [M, features] fed by the mod tiles; gate [M, N - 2 features] fed by the remaining tiles
    NOVA_PROPAGATE(make_tmap_kmajor(&tc_, C, M, ada->features, ldc, 32));
    const int gate_cols = N - 2 * ada->features;
    if (gate_cols > 0) NOVA_PROPAGATE(make_tmap_kmajor(&tc2, ada->gate, M, gate_cols, ada->ldg, 32));
    else tc2 = tc_;
    NOVA_PROPAGATE(make_tmap_kmajor(&tc3, ada->x, M, ada->features, ada->ldx, 32));  // x chunks into the C staging buffers
    p.x = ada->x; p.ldx = ada->ldx; p.rowstats = ada->rowstats; p.n_mod_tiles = 2 * ada->features / BN;
    p.part_in = ada->parts; p.stats_parts = ada->n_parts;
  } else if (EPI == EPI_TAIL) {
    NOVA_PROPAGATE(make_tmap_kmajor(&tc_, tail->x, M, N, tail->ldx, 32));    // new x: stores of 32 x 64 sub-tiles
    NOVA_PROPAGATE(make_tmap_kmajor(&tc2, tail->u, M, N, tail->ldu, BM));    // staged u chunks: 128 x 64
    NOVA_PROPAGATE(make_tmap_kmajor(&tc3, tail->x, M, N, tail->ldx, BM));    // staged x chunks
    p.part_in = tail->u_parts; p.stats_parts = tail->n_parts; p.part_out = tail->x_parts;
    p.gamma = tail->gamma; p.beta = tail->beta;
  } else {
    NOVA_PROPAGATE(make_tmap_kmajor(&tc_, C, M, N, ldc, 32));
    tc2 = tc_;
    tc3 = tc_;
    if (EPI == EPI_BIAS_SILU_DUAL) NOVA_PROPAGATE(make_tmap_kmajor(&tc2, pre_out, M, N, ldpre, 32));  // the pre-activation
  }
  const int tiles = static_cast<int>(ceil_div(M, BM * CG) * ceil_div(N, BN));
  int groups = tiles < num_sms() / CG ? tiles : num_sms() / CG;
  // Fixed-column grid (NOVA_B200_FIXED_N=1; off by default): with a group count that is a multiple of the column tiles,
  // group g works on column tile g % num_n for the whole launch (tile = g + i * groups), so the epilogue keeps its bias /
  // gamma / beta tile in shared memory and drops the two epilogue-wide barriers per tile around it (ncu source view of the
  // tail GEMM, round 2: 350 of the epilogue warps' ~3000 samples sat behind those two barriers).  Taken only when it
  // costs no extra wave: 768 tiles (cfg2's fc and gate GEMMs) are 11 waves on 74 CTA pairs and on 72.  Measured at cfg2,
  // same box, twice: 66.4 / 67.0 ms per pass with, 66.4 / 66.9 without -- the power-capped step does not see the stall.
  const int num_n_tiles = static_cast<int>(ceil_div(N, BN));
  if (fixed_column_grid() && groups > num_n_tiles && groups % num_n_tiles != 0) {
    const int g2 = groups / num_n_tiles * num_n_tiles;
    if (ceil_div(tiles, g2) == ceil_div(tiles, groups)) groups = g2;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(groups * CG));
  cfg.blockDim = dim3(num_threads(EPI));
  cfg.dynamicSmemBytes = P::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // Programmatic dependent launch is per epilogue kind (pdl_epi_mask): the tail GEMM stays an ordinary launch.  With
  // mod -> fc1 -> fc2 -> tail ALL dependent launches the step has no full kernel boundary left, and that chain is not
  // safe: measured at 14 336 / 16 384 rows, 0.1-0.7 % of the rows came out one bf16 ulp off and runs stopped being
  // reproducible; taking ANY one of the four kinds out of the chain restores bit-identical results (a gpu-scope fence
  // after griddepcontrol.wait -- i.e. an L1 invalidate -- does not).  The unfused flow never had the problem because its
  // resid kernel is an ordinary launch.
  cfg.numAttrs = (pdl_enabled() && ((pdl_epi_mask() >> (EPI == EPI_BIAS_SILU_DUAL ? EPI_BIAS_SILU : EPI)) & 1)) ? 2 : 1;
  NOVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<EPI, CG, BN>, ta, tb, tc_, tc2, tc3, p, debug_word()));
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

// cta_group selection: 0 = automatic (CTA pairs once there are at least 2 x 128 rows), 1, 2
// The instantiations are spread over three translation units so that they compile in parallel:
//   gemm_bias.cu (NOVA_GEMM_TU == 0): EPI_BIAS kernels + the `launch` dispatcher,
//   gemm_silu.cu (NOVA_GEMM_TU == 1): EPI_BIAS_SILU kernels,   gemm_adaln.cu (NOVA_GEMM_TU == 2): EPI_ADALN kernels.
template <int EPI>
int launch_plain(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc, int M,
                 int N, int K, cudaStream_t stream, int cta_group, int bn, bool reverse_m, float2* part_out = nullptr,
                 int batch_m_rows = 0, int batch_w_rows = 0, int mn_rows = 0) {
#define NOVA_GEMM_CASE(G, B) \
  if (cta_group == G && bn == B) \
    return launch_epi<EPI, G, B>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, nullptr, reverse_m, part_out, nullptr, \
                                 batch_m_rows, batch_w_rows, mn_rows);
  NOVA_GEMM_CASE(2, 256) NOVA_GEMM_CASE(1, 256) NOVA_GEMM_CASE(2, 128) NOVA_GEMM_CASE(1, 128) NOVA_GEMM_CASE(2, 64)
  NOVA_GEMM_CASE(1, 64)
#undef NOVA_GEMM_CASE
  set_error("tcgen05 gemm: unsupported cta_group %d / tile columns %d", cta_group, bn);
  return NOVA_ERR_INVALID;
}
int launch_bias(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc, int M,
                int N, int K, cudaStream_t stream, int cta_group, int bn, bool reverse_m, float2* part_out,
                int batch_m_rows = 0, int batch_w_rows = 0, int mn_rows = 0);
int launch_silu(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc, int M,
                int N, int K, cudaStream_t stream, int cta_group, int bn, bool reverse_m);

#if NOVA_GEMM_TU == 1
int launch_silu(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc, int M,
                int N, int K, cudaStream_t stream, int cta_group, int bn, bool reverse_m) {
  return launch_plain<EPI_BIAS_SILU>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, cta_group, bn, reverse_m);
}

// pre [M, N] = A W^T + bias and act [M, N] = silu(that) from ONE launch (256-column tiles): the training forward keeps the
// pre-activation for the backward pass without a separate SiLU kernel re-reading and re-writing [M, N]
int launch_silu_dual(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* pre, int64_t ldpre,
                     bf16* act, int64_t ldact, int M, int N, int K, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return NOVA_OK;
  NOVA_REQUIRE(K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldpre % 8 == 0 && ldact % 8 == 0,
               "tcgen05 dual gemm: K and leading dimensions must be multiples of 8");
  NOVA_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(pre) & 15) == 0 && (reinterpret_cast<uintptr_t>(act) & 15) == 0 && pre && act,
               "tcgen05 dual gemm: operands must be 16-byte aligned");
  if (default_cta_group(M) == 2)
    return launch_epi<EPI_BIAS_SILU_DUAL, 2, BN_FULL>(A, lda, W, ldw, bias, act, ldact, M, N, K, stream, nullptr, false, nullptr,
                                                      nullptr, 0, 0, 0, pre, ldpre);
  return launch_epi<EPI_BIAS_SILU_DUAL, 1, BN_FULL>(A, lda, W, ldw, bias, act, ldact, M, N, K, stream, nullptr, false, nullptr,
                                                    nullptr, 0, 0, 0, pre, ldpre);
}
#endif

#if NOVA_GEMM_TU == 0
int launch_bias(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc, int M,
                int N, int K, cudaStream_t stream, int cta_group, int bn, bool reverse_m, float2* part_out,
                int batch_m_rows, int batch_w_rows, int mn_rows) {
  return launch_plain<EPI_BIAS>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, cta_group, bn, reverse_m, part_out,
                                batch_m_rows, batch_w_rows, mn_rows);
}

// The same S products WITHOUT transposed copies (MN-major operands): C[b] [n_rows, N] = Y[b]^T X[b], where Y is the
// source matrix [src_rows, n_rows] (row stride ldy), X the source matrix [src_rows, N] (row stride ldx), both row-major,
// and batch b reduces over source rows [b k_len, (b + 1) k_len) (rows >= src_rows read as zero).  Outputs stacked.
int launch_batched_mn(const bf16* Y, int64_t ldy, const bf16* X, int64_t ldx, bf16* C, int64_t ldc, int src_rows, int n_rows,
                      int N, int k_len, int batches, cudaStream_t stream) {
  if (n_rows <= 0 || N <= 0 || batches <= 0) return NOVA_OK;
  NOVA_REQUIRE(n_rows % (2 * BM) == 0, "tcgen05 MN-major gemm: output rows per batch must be a multiple of %d", 2 * BM);
  NOVA_REQUIRE(k_len > 0 && k_len % BK == 0 && ldy % 8 == 0 && ldx % 8 == 0 && ldc % 8 == 0 && N % 8 == 0,
               "tcgen05 MN-major gemm: reduction length must be a multiple of %d, leading dimensions of 8", BK);
  NOVA_REQUIRE((reinterpret_cast<uintptr_t>(Y) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "tcgen05 MN-major gemm: operands must be 16-byte aligned");
  return launch_bias(Y, ldy, X, ldx, nullptr, C, ldc, batches * n_rows, N, k_len, stream, 2, BN_FULL, false, nullptr, n_rows, N,
                     src_rows);
}

// S = M / batch_m_rows GEMMs in one launch: C[b] [batch_m_rows, N] = A[b] [batch_m_rows, K] W[b] [batch_w_rows = N, K]^T,
// operands and outputs stacked along their rows (CTA pairs, 256-column tiles).
int launch_batched(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, bf16* C, int64_t ldc, int M, int N, int K,
                   int batch_m_rows, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return NOVA_OK;
  NOVA_REQUIRE(batch_m_rows > 0 && batch_m_rows % (2 * BM) == 0 && M % batch_m_rows == 0,
               "tcgen05 batched gemm: rows per batch must be a multiple of %d and divide M", 2 * BM);
  NOVA_REQUIRE(K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "tcgen05 batched gemm: K and leading dimensions must be multiples of 8");
  NOVA_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "tcgen05 batched gemm: operands must be 16-byte aligned");
  return launch_bias(A, lda, W, ldw, nullptr, C, ldc, M, N, K, stream, 2, BN_FULL, false, nullptr, batch_m_rows, N);
}

int launch(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* C, int64_t ldc, int M, int N,
           int K, int epi, cudaStream_t stream, int cta_group, bool reverse_m, float2* part_out) {
  if (M <= 0 || N <= 0) return NOVA_OK;
  NOVA_REQUIRE(part_out == nullptr || (epi == EPI_BIAS && N % BN_FULL == 0),
               "tcgen05 gemm: partial statistics need the bias epilogue and N a multiple of %d", BN_FULL);
  NOVA_REQUIRE(K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "tcgen05 gemm: K, lda, ldw must be multiples of 8");
  NOVA_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 8 == 0,
               "tcgen05 gemm: operands must be 16-byte aligned");
  NOVA_REQUIRE(epi == EPI_BIAS || epi == EPI_BIAS_SILU, "tcgen05 gemm: unsupported epilogue %d", epi);
  if (cta_group == 0) cta_group = default_cta_group(M);
  // Tile columns: 256 unless that leaves most SMs without a tile (small M: the set-by-set pattern); then the same
  // work is cut into 128- or 64-column tiles -- more CTAs busy and a 2-4x shorter MMA chain per launch.
  const int units = num_sms() / cta_group;
  const int64_t row_blocks = ceil_div(M, BM * cta_group);
  int bn = BN_FULL;
  if (part_out != nullptr) bn = BN_FULL;  // one partial per 256-column tile
  else if (tile_columns_override() > 0) bn = tile_columns_override();
  else if (row_blocks * ceil_div(N, 256) * 2 <= units) bn = row_blocks * ceil_div(N, 128) * 2 <= units ? 64 : 128;
  return epi == EPI_BIAS ? launch_bias(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, cta_group, bn, reverse_m, part_out)
                         : launch_silu(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, cta_group, bn, reverse_m);
}
#endif

#if NOVA_GEMM_TU == 2
// AdaLN statistics GEMM with the modulation fused into the epilogue:
//   W [2 features + gate_cols, K] packed per 128 features as [scale | shift], then gate rows;
//   h [M, features] = LN(x)(1 + scale) + shift,  gate [M, gate_cols] = a W_gate^T + b_gate.
int launch_adaln(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, bf16* h_out, int64_t ldh,
                 const AdaLNArgs& ada, int M, int N, int K, cudaStream_t stream, int cta_group, bool reverse_m) {
  if (M <= 0 || N <= 0) return NOVA_OK;
  NOVA_REQUIRE(K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldh % 8 == 0 && ada.ldx % 8 == 0,
               "tcgen05 adaln gemm: K and leading dimensions must be multiples of 8");
  NOVA_REQUIRE(ada.features % 128 == 0 && (2 * ada.features) % BN_FULL == 0 && N >= 2 * ada.features &&
                   N % BN_FULL == 0,
               "tcgen05 adaln gemm: features must be a multiple of 128 and N a multiple of %d", BN_FULL);
  NOVA_REQUIRE(ada.x && ada.rowstats && (N == 2 * ada.features || ada.gate), "tcgen05 adaln gemm: null operand");
  if (cta_group == 0) cta_group = default_cta_group(M);
  if (cta_group == 2)
    return launch_epi<EPI_ADALN, 2>(A, lda, W, ldw, bias, h_out, ldh, M, N, K, stream, &ada, reverse_m);
  return launch_epi<EPI_ADALN, 1>(A, lda, W, ldw, bias, h_out, ldh, M, N, K, stream, &ada, reverse_m);
}

#endif  // NOVA_GEMM_TU == 2

#if NOVA_GEMM_TU == 3
int launch_tail(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, const TailArgs& tail, int M, int N,
                int K, cudaStream_t stream, bool reverse_m) {
  if (M <= 0 || N <= 0) return NOVA_OK;
  NOVA_REQUIRE(K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && tail.ldu % 8 == 0 && tail.ldx % 8 == 0,
               "tcgen05 tail gemm: K and leading dimensions must be multiples of 8");
  NOVA_REQUIRE(N % BN_FULL == 0 && tail.n_parts == N / BN_FULL && tail.n_parts <= 8,
               "tcgen05 tail gemm: N must be a multiple of %d (<= 2048) with one u partial per tile", BN_FULL);
  NOVA_REQUIRE(tail.u && tail.x && tail.gamma && tail.beta && tail.u_parts && tail.x_parts, "tcgen05 tail gemm: null operand");
  return launch_epi<EPI_TAIL, 2>(A, lda, W, ldw, bias, nullptr, 0, M, N, K, stream, nullptr, reverse_m, nullptr, &tail);
}
#endif  // NOVA_GEMM_TU == 3

}  // namespace tc
}  // namespace nova
