// Earth mover's distance between equal-size point clouds on the GPU (the last geometry op).
//
// The reference solves the assignment problem on the CPU: dist = cdist(a, b), scipy's linear_sum_assignment
// (Hungarian / Jonker-Volgenant, O(N^3)), EMD = mean of the matched distances
//   emd_approx              train_newloss.py:352-377   (inputs clamped to [-2, 2], distances to >= 1e-8)
//   earth_mover_distance    demo.py:57-74
//   compute_emd             test_optimize.py:395-414   (result clamped to [0, 10])
// The minimum of the assignment problem is unique even when the matching is not, so the quantity to reproduce is the
// optimal COST.  A Hungarian augmenting path is inherently sequential; what maps to a GPU is Bertsekas' auction
// algorithm with epsilon scaling: unassigned points of `a` ("persons") bid for points of `b` ("objects"),
//   v_ij = -|a_i - b_j| - price_j ;  j1 = argmax_j v_ij ;  bid = price_j1 + (v_i,j1 - second best v) + eps
// every object goes to its highest bidder and its price rises; a phase ends when everybody is assigned, and the result
// of the last phase is within N * eps of the optimum (mean distance within eps).  eps runs from max-distance / 2 down to
// `eps_final` by factors of 5, prices kept between phases.
//
// One CTA per cloud pair.  Both clouds, prices, owners and the bid table live in shared memory (52 B per point: up to
// 4096 points; the objects as three coordinate planes); a bid is one warp: its lanes scan the objects two at a time on
// packed fp32 pairs (distances recomputed from the coordinates -- cheaper than a 16 MB matrix per pair through L2), a shuffle tree merges (best, second best), lane 0 posts the bid with a 64-bit
// atomicMax of (bid bits, person) -- unique keys, so the winner does not depend on the order of arrival.  The list of
// unassigned persons is rebuilt by an ordered block scan.  Everything is deterministic run to run; every loop is
// bounded (a cloud that does not converge within the round budget is reported in `status`, never spun on).
// Distances are exact fp32 differences (the Chamfer kernel's arithmetic); prices are fp32 and re-based to min = 0 at
// every phase so that eps_final = 1e-5 stays well above their rounding step.
#include <atomic>
#include <cfloat>
#include <cstdlib>

#include "common.cuh"

namespace nova {
namespace emd {

constexpr int MAX_THREADS = 1024;  // the kernel runs with 1024 or 512 threads per CTA (nova_emd picks), same results
constexpr int MAX_POINTS = 4096;

__device__ __forceinline__ float dist(const float4 a, const float4 b) {
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
}
// ---- packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): a lane evaluates objects (j, j + 1) with one set of instructions
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
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
// The bidding loop's square root as MUFU.SQRT (<= 2 ulp) instead of the correctly rounded one (MUFU.RSQ + a Newton step +
// fix-up: ~8 of the ~25 instructions of a (person, object) evaluation).  Only the VALUES that steer the auction see it --
// the matching stays eps-optimal for costs that differ from the exact ones by 2 ulp (1e-7 relative, against eps_final =
// 1e-5) -- the reported mean is always summed from exact distances.
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// monotone map float -> uint32 (bids are finite and may be negative after re-basing)
__device__ __forceinline__ uint32_t ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct Best {
  float v1, v2;  // best and second-best value
  int j1;        // object of the best value (lowest index on ties)
};
__device__ __forceinline__ void offer(Best& b, float v, int j) {
  if (v > b.v1 || (v == b.v1 && j < b.j1)) {
    b.v2 = b.v1;
    b.v1 = v;
    b.j1 = j;
  } else if (v > b.v2) {
    b.v2 = v;
  }
}
// the same for a lane's own scan, where j only grows (an equal later value never displaces the earlier index): no branch
__device__ __forceinline__ void offer_ascending(Best& b, float v, int j) {
  const bool better = v > b.v1;
  b.v2 = fmaxf(b.v2, fminf(v, b.v1));
  b.j1 = better ? j : b.j1;
  b.v1 = fmaxf(b.v1, v);
}
__device__ __forceinline__ Best merge(const Best& a, float ov1, float ov2, int oj1) {
  Best r;
  if (ov1 > a.v1 || (ov1 == a.v1 && oj1 < a.j1)) {
    r.v1 = ov1; r.j1 = oj1;
    r.v2 = fmaxf(a.v1, ov2);
  } else {
    r.v1 = a.v1; r.j1 = a.j1;
    r.v2 = fmaxf(a.v2, ov1);
  }
  return r;
}

// block-wide exclusive scan of one int per thread (<= 32 warps), returns the total in `total`
template <int NWARPS>
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < NWARPS ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < NWARPS) warp_sums[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  total = warp_sums[NWARPS - 1];
  const int base = warp == 0 ? 0 : warp_sums[warp - 1];
  __syncthreads();  // warp_sums may be reused by the caller
  return base + incl - v;
}

// THREADS = 1024: one CTA per SM, every warp of the SM on one pair (few pairs).  THREADS = 512: two CTAs per SM at
// 2048 points (106 KB of shared memory each), so that 256 pairs are ONE wave on 148 SMs instead of 148 + 108 -- a pair's
// late rounds (a handful of bidders, barrier-paced) leave most of an SM idle, which the second resident pair fills.
// Bids, prices and the compaction order do not depend on the thread count; the closing sum is taken in the 1024-thread
// order by both, so the two variants return the same bits.
template <int THREADS, bool FAST_SQRT>
static __global__ void __launch_bounds__(THREADS, MAX_THREADS / THREADS)
auction_kernel(const float* __restrict__ A, const float* __restrict__ Bp, int N, float eps_final, int max_rounds,
               float* __restrict__ emd_out, int32_t* __restrict__ assign_out, int32_t* __restrict__ status_out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float4* sa = reinterpret_cast<float4*>(smem_raw);                                  // persons
  const int Np = (N + 1) & ~1;                                                       // objects: three coordinate planes of
  float* bx = reinterpret_cast<float*>(sa + N);                                      // Np floats (a lane reads objects j, j + 1
  float* by = bx + Np;                                                               // of a plane with one LDS.64)
  float* bz = by + Np;
  unsigned long long* bidkey = reinterpret_cast<unsigned long long*>(bz + Np);       // per object: (ordered bid, ~person)
  float* price = reinterpret_cast<float*>(bidkey + N);
  int* owner = reinterpret_cast<int*>(price + N);      // person holding object j, -1 if none
  int* mine = owner + N;                               // object held by person i, -1 if none
  int* todo = mine + N;                                // compact list of unassigned persons
  constexpr int NWARPS = THREADS / 32;
  __shared__ int warp_sums[NWARPS];
  __shared__ float red[MAX_THREADS / 32];
  __shared__ int n_todo;

  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pa = A + (size_t)pair * N * 3;
  const float* pb = Bp + (size_t)pair * N * 3;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int j = tid; j < N; j += THREADS) {
    const float4 a = make_float4(pa[j * 3], pa[j * 3 + 1], pa[j * 3 + 2], 0.f);
    const float4 b = make_float4(pb[j * 3], pb[j * 3 + 1], pb[j * 3 + 2], 0.f);
    sa[j] = a;
    bx[j] = b.x;
    by[j] = b.y;
    bz[j] = b.z;
    price[j] = 0.f;
    lo[0] = fminf(lo[0], fminf(a.x, b.x)); hi[0] = fmaxf(hi[0], fmaxf(a.x, b.x));
    lo[1] = fminf(lo[1], fminf(a.y, b.y)); hi[1] = fmaxf(hi[1], fmaxf(a.y, b.y));
    lo[2] = fminf(lo[2], fminf(a.z, b.z)); hi[2] = fmaxf(hi[2], fmaxf(a.z, b.z));
  }
  if (tid == 0 && Np != N) bx[N] = by[N] = bz[N] = 0.f;  // the pad object of an odd cloud is evaluated, never offered
  // upper bound of any distance: the diagonal of the joint bounding box
  float diag2 = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float l = lo[k], h = hi[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
      h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
    }
    if (lane == 0) { red[warp] = l; }
    __syncthreads();
    float bl = FLT_MAX;
    for (int w = 0; w < NWARPS; ++w) bl = fminf(bl, red[w]);
    __syncthreads();
    if (lane == 0) { red[warp] = h; }
    __syncthreads();
    float bh = -FLT_MAX;
    for (int w = 0; w < NWARPS; ++w) bh = fmaxf(bh, red[w]);
    __syncthreads();
    const float e = fmaxf(bh - bl, 0.f);
    diag2 = fmaf(e, e, diag2);
  }
  const float cmax = sqrtf(diag2);

  int rounds = 0;
  bool failed = false;
  float eps = fmaxf(cmax * 0.5f, eps_final);
  for (;;) {  // ---------------------------------------------------------------- epsilon phases
    // re-base the prices (only differences matter) and start the phase with nobody assigned
    float pmin = FLT_MAX;
    for (int j = tid; j < N; j += THREADS) pmin = fminf(pmin, price[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    if (lane == 0) red[warp] = pmin;
    __syncthreads();
    pmin = FLT_MAX;
    for (int w = 0; w < NWARPS; ++w) pmin = fminf(pmin, red[w]);
    __syncthreads();
    for (int j = tid; j < N; j += THREADS) {
      price[j] -= pmin;
      owner[j] = -1;
      mine[j] = -1;
      todo[j] = j;
      bidkey[j] = 0ull;
    }
    if (tid == 0) n_todo = N;
    __syncthreads();
    while (n_todo > 0 && rounds < max_rounds) {  // ------------------------------ bidding rounds (Jacobi)
      ++rounds;
      const int cnt = n_todo;
      for (int u = warp; u < cnt; u += NWARPS) {  // one warp per bidding person
        const int i = todo[u];
        const float4 a = sa[i];
        Best b{-FLT_MAX, -FLT_MAX, 0x7fffffff};
        // lane l scans objects (2l, 2l + 1), (2l + 64, 2l + 65), ...: (b - a)^2 on packed pairs, the same bits as dist()
        const f32x2 nax = pk2(-a.x, -a.x), nay = pk2(-a.y, -a.y), naz = pk2(-a.z, -a.z);
#pragma unroll 2
        for (int j = 2 * lane; j < N; j += 64) {
          const f32x2 dx = add2(*reinterpret_cast<const f32x2*>(bx + j), nax);
          const f32x2 dy = add2(*reinterpret_cast<const f32x2*>(by + j), nay);
          const f32x2 dz = add2(*reinterpret_cast<const f32x2*>(bz + j), naz);
          float d0, d1;
          upk2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), d0, d1);
          const float2 pr = *reinterpret_cast<const float2*>(price + j);
          d0 = FAST_SQRT ? sqrt_fast(d0) : sqrtf(d0);
          d1 = FAST_SQRT ? sqrt_fast(d1) : sqrtf(d1);
          offer_ascending(b, -d0 - pr.x, j);
          if (j + 1 < N) offer_ascending(b, -d1 - pr.y, j + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov1 = __shfl_xor_sync(0xffffffffu, b.v1, o), ov2 = __shfl_xor_sync(0xffffffffu, b.v2, o);
          const int oj1 = __shfl_xor_sync(0xffffffffu, b.j1, o);
          b = merge(b, ov1, ov2, oj1);
        }
        if (lane == 0) {
          const float gap = N > 1 ? b.v1 - b.v2 : 0.f;
          float bid = price[b.j1] + gap + eps;
          bid = fmaxf(bid, nextafterf(price[b.j1], FLT_MAX));  // a bid always raises the price
          const unsigned long long key = (static_cast<unsigned long long>(ordered(bid)) << 32) |
                                         static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(i));
          atomicMax(&bidkey[b.j1], key);  // highest bid wins, lowest person index on equal bids
        }
      }
      __syncthreads();
      for (int j = tid; j < N; j += THREADS) {  // every object with a bid changes hands
        const unsigned long long key = bidkey[j];
        if (key != 0ull) {
          const int i = static_cast<int>(0xffffffffu - static_cast<uint32_t>(key & 0xffffffffull));
          const int old = owner[j];
          if (old >= 0) mine[old] = -1;  // `old` bid for nothing this round (it was assigned): no conflict with mine[i]
          owner[j] = i;
          price[j] = unordered(static_cast<uint32_t>(key >> 32));
          bidkey[j] = 0ull;
        }
      }
      __syncthreads();
      for (int j = tid; j < N; j += THREADS) {  // after the evictions: winners take their objects
        const int i = owner[j];
        if (i >= 0) mine[i] = j;
      }
      __syncthreads();
      // ordered compaction of the persons without an object
      const int per = (N + THREADS - 1) / THREADS;
      const int i0 = tid * per, i1 = min(N, i0 + per);
      int c = 0;
      for (int i = i0; i < i1; ++i) c += mine[i] < 0 ? 1 : 0;
      int total;
      int pos = block_excl_scan<NWARPS>(c, warp_sums, total);
      for (int i = i0; i < i1; ++i)
        if (mine[i] < 0) todo[pos++] = i;
      if (tid == 0) n_todo = total;
      __syncthreads();
    }
    if (n_todo > 0) { failed = true; break; }
    if (eps <= eps_final) break;
    eps = fmaxf(eps * 0.2f, eps_final);
  }
  __syncthreads();
  // mean matched distance (fixed-order tree) and the assignment; persons left without an object (budget exhausted)
  // take the objects nobody owns, in index order, so that the output is always a permutation
  if (failed) {
    const int per = (N + THREADS - 1) / THREADS;
    const int i0 = tid * per, i1 = min(N, i0 + per);
    int c = 0;
    for (int j = i0; j < i1; ++j) c += owner[j] < 0 ? 1 : 0;
    int total;
    int pos = block_excl_scan<NWARPS>(c, warp_sums, total);
    for (int j = i0; j < i1; ++j)
      if (owner[j] < 0) {
        const int i = todo[pos++];
        mine[i] = j;
      }
    __syncthreads();
  }
  // person i belongs to virtual thread i % 1024; a real thread sums its MAX_THREADS / THREADS virtual threads separately
#pragma unroll
  for (int v = 0; v < MAX_THREADS / THREADS; ++v) {
    float acc = 0.f;
    for (int i = tid + v * THREADS; i < N; i += MAX_THREADS) {
      const int j = mine[i];
      acc += dist(sa[i], make_float4(bx[j], by[j], bz[j], 0.f));
      if (assign_out != nullptr) assign_out[(size_t)pair * N + i] = j;
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp + v * NWARPS] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < MAX_THREADS / 32; ++w) s += red[w];
    emd_out[pair] = s / static_cast<float>(N);
    if (status_out != nullptr) status_out[pair] = failed ? -rounds : rounds;
  }
}

}  // namespace emd
}  // namespace nova

using namespace nova;

extern "C" int nova_emd(const float* a, const float* b, int64_t B, int64_t N, float eps_final, int32_t max_rounds,
                        float* emd_out, int32_t* assign_out, int32_t* status_out, void* stream) {
  NOVA_REQUIRE(B >= 0 && N >= 0, "nova_emd: bad sizes B=%lld N=%lld", (long long)B, (long long)N);
  if (B == 0) return NOVA_OK;
  NOVA_REQUIRE(N >= 1, "nova_emd: empty clouds (the reference's linear_sum_assignment mean is undefined there)");
  NOVA_REQUIRE(N <= emd::MAX_POINTS, "nova_emd: at most %d points per cloud (got %lld)", emd::MAX_POINTS, (long long)N);
  NOVA_REQUIRE(a && b && emd_out, "nova_emd: null argument");
  NOVA_REQUIRE(eps_final > 0.f && max_rounds > 0, "nova_emd: eps_final and max_rounds must be positive");
  const size_t smem = static_cast<size_t>(N) * (16 + 8 + 4 + 4 + 4 + 4) + static_cast<size_t>((N + 1) & ~1ll) * 12 + 8;
  static std::atomic<unsigned long long> attr_done{0ull};
  int dev = 0;
  NOVA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !((attr_done.load() >> dev) & 1ull)) {
    const int cap = 227 * 1024 - 1024;
    NOVA_CHECK_CUDA(cudaFuncSetAttribute(emd::auction_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    NOVA_CHECK_CUDA(cudaFuncSetAttribute(emd::auction_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    NOVA_CHECK_CUDA(cudaFuncSetAttribute(emd::auction_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    NOVA_CHECK_CUDA(cudaFuncSetAttribute(emd::auction_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    attr_done.fetch_or(1ull << dev);
  }
  // more pairs than SMs and two CTAs' clouds fit one SM: 512 threads per pair, two pairs per SM (NOVA_B200_EMD_THREADS
  // = 512 | 1024 forces either; the results are the same bits).  Measured at 256 pairs of 2048 points: 397 -> 288 ms;
  // at 32 pairs the 1024-thread form is the faster one (246 vs 263 ms).
  int sms = 0;
  NOVA_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const char* env = std::getenv("NOVA_B200_EMD_THREADS");
  const int forced = env ? std::atoi(env) : 0;
  // The approximate square root steers the bids only (sqrt_fast); on by default: 288 -> 227 ms at 256 pairs, the mean
  // matched distance moves by <= 1.2e-7 (NOVA_B200_EMD_FAST_SQRT=0 restores the correctly rounded one).
  const char* fenv = std::getenv("NOVA_B200_EMD_FAST_SQRT");
  const bool fast = fenv == nullptr || std::atoi(fenv) != 0;
  bool half = forced == 512;
  if (forced != 512 && forced != 1024 && B > sms) {
    int resident = 0;
    NOVA_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &resident, fast ? emd::auction_kernel<512, true> : emd::auction_kernel<512, false>, 512, smem));
    half = resident >= 2;
  }
  auto* kern = half ? (fast ? emd::auction_kernel<512, true> : emd::auction_kernel<512, false>)
                    : (fast ? emd::auction_kernel<1024, true> : emd::auction_kernel<1024, false>);
  kern<<<static_cast<unsigned>(B), half ? 512 : 1024, smem, static_cast<cudaStream_t>(stream)>>>(
      a, b, static_cast<int>(N), eps_final, max_rounds, emd_out, assign_out, status_out);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}
