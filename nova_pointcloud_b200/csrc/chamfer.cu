// Chamfer nearest-neighbour primitive for sm_100a.
//
//   d1[b,i] = min_j |a_bi - b_bj|,  d2[b,j] = min_i |a_bi - b_bj|   (Euclidean, not squared)
//
// serves the three reductions the reference applies:
//   A  demo.py:38-55 (scipy cdist + min)      B  train_newloss.py:316-349      C  test_optimize.py:354-383
//
// Brute force, exact difference form sum((x-y)^2) in fp32 (the reference's torch.cdist uses
// the |x|^2+|y|^2-2xy matrix form above 25 points and loses ~1e-5; scipy's float64 is the pin).
//
// Mapping: blockIdx = (query chunk, cloud, direction).  A CTA owns QPB = 32*Q consecutive
// query points (Q per lane, in registers); the target cloud streams through shared memory in
// tiles of TILE points stored as float4 (x,y,z,0) so one LDS.128 broadcast feeds all lanes.
// The 8 warps split every tile (warp w takes points w, w+8, ...) to keep small batches busy;
// their partial minima meet in shared memory at the end.  Loads of the xyz stream are
// coalesced float reads; bandwidth is irrelevant here (16.8 MB for 1.07 G pair evaluations),
// the limiter is fp32 issue rate: 3 FADD + 3 FFMA/FMUL + compare/select per pair.
#include <cstdlib>

#include "common.cuh"

namespace nova {
namespace chamfer {

constexpr int Q = 4;          // queries per lane
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int QPB = 32 * Q;   // queries per CTA
constexpr int TILE = 1024;    // target points per shared-memory tile

// IDX = false drops the arg-min bookkeeping from the inner loop (FMNMX instead of compare + two selects):
// the kernel is issue-bound (87 % issue-slot utilisation in the round-1 ncu capture), so that is ~30 % fewer
// instructions per pair for the distance-only reductions (Chamfer A/B/C).
template <bool IDX>
__global__ void __launch_bounds__(THREADS)
nn_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t N, int64_t M, float* __restrict__ d1,
          float* __restrict__ d2, int32_t* __restrict__ idx1, int32_t* __restrict__ idx2) {
  __shared__ float4 tile[TILE];
  __shared__ float red_d[WARPS][QPB];
  __shared__ int red_i[WARPS][QPB];

  const int dir = blockIdx.z;
  const int64_t cloud = blockIdx.y;
  const float* qry = (dir == 0 ? a + cloud * N * 3 : b + cloud * M * 3);
  const float* tgt = (dir == 0 ? b + cloud * M * 3 : a + cloud * N * 3);
  const int64_t nq = dir == 0 ? N : M, nt = dir == 0 ? M : N;
  float* dout = (dir == 0 ? d1 + cloud * N : d2 + cloud * M);
  int32_t* iout = dir == 0 ? (idx1 ? idx1 + cloud * N : nullptr) : (idx2 ? idx2 + cloud * M : nullptr);

  const int64_t q0 = (int64_t)blockIdx.x * QPB;
  if (q0 >= nq) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float qx[Q], qy[Q], qz[Q], best[Q];
  int bidx[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    int64_t qi = q0 + k * 32 + lane;
    if (qi >= nq) qi = nq - 1;  // clamp: duplicates are computed and dropped at the store
    qx[k] = qry[qi * 3 + 0];
    qy[k] = qry[qi * 3 + 1];
    qz[k] = qry[qi * 3 + 2];
    best[k] = 3.4e38f;
    bidx[k] = 0;
  }

  for (int64_t t0 = 0; t0 < nt; t0 += TILE) {
    const int cnt = static_cast<int>(nt - t0 < TILE ? nt - t0 : TILE);
    __syncthreads();
    // coalesced read of cnt*3 consecutive floats, scattered into float4 slots
    float* tf = reinterpret_cast<float*>(tile);
    for (int i = threadIdx.x; i < cnt * 3; i += THREADS) {
      const int pt = i / 3, c = i - pt * 3;
      tf[pt * 4 + c] = tgt[t0 * 3 + i];
    }
    __syncthreads();
#pragma unroll 4
    for (int j = warp; j < cnt; j += WARPS) {
      const float4 t = tile[j];
      const int gj = static_cast<int>(t0) + j;
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        const float dx = qx[k] - t.x, dy = qy[k] - t.y, dz = qz[k] - t.z;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (IDX) {
          if (d < best[k]) {  // strict: first (lowest) index wins within a warp's stride
            best[k] = d;
            bidx[k] = gj;
          }
        } else {
          best[k] = fminf(best[k], d);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    red_d[warp][k * 32 + lane] = best[k];
    red_i[warp][k * 32 + lane] = bidx[k];
  }
  __syncthreads();
  if (threadIdx.x < QPB) {
    float m = red_d[0][threadIdx.x];
    int mi = red_i[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
      const float v = red_d[w][threadIdx.x];
      const int vi = red_i[w][threadIdx.x];
      if (v < m || (v == m && vi < mi)) {  // ties -> lowest index, like argmin
        m = v;
        mi = vi;
      }
    }
    const int64_t qi = q0 + threadIdx.x;
    if (qi < nq) {
      dout[qi] = sqrtf(m);
      if (iout) iout[qi] = mi;
    }
  }
}


// ------------------------------------------------------------------ one sweep, both directions (distance only)
// Every |a_i - b_j|^2 is evaluated ONCE and feeds both minima: the running minimum of query i (a register) and
// the minimum of target j over the warp's 32 * SQ queries.  The second one is a reduction ACROSS lanes; doing it
// per target would cost a 5-step shuffle tree per 4 pair evaluations, so targets are taken 32 at a time and the
// 32 x 32 (target x lane) partial minima are reduced with one butterfly transpose: 31 shuffles per 32 targets,
// after which lane l holds the warp's minimum for target l.  Warps and CTAs then meet through atomicMin on the
// bit pattern of d^2 (non-negative floats order like unsigned integers): shared memory first, global once per tile.
#ifndef NOVA_CHAMFER_SQ
#define NOVA_CHAMFER_SQ 4
#endif
#ifndef NOVA_CHAMFER_SWARPS
#define NOVA_CHAMFER_SWARPS 4
#endif
constexpr int SQ = NOVA_CHAMFER_SQ;          // queries per lane
constexpr int SWARPS = NOVA_CHAMFER_SWARPS;  // each warp owns its own 32 * SQ queries and sweeps every target
constexpr int STHREADS = SWARPS * 32;
constexpr int SQPB = SWARPS * 32 * SQ;  // 512 queries per CTA
constexpr int STILE = 1024;       // targets per shared-memory tile (multiple of 32)

__global__ void __launch_bounds__(STHREADS)
nn_sym_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t N, int64_t M, float* __restrict__ d1,
              unsigned int* __restrict__ d2sq_bits) {
  __shared__ float4 tile[STILE];
  __shared__ unsigned int tmin[STILE];
  const int64_t cloud = blockIdx.y;
  const float* qry = a + cloud * N * 3;
  const float* tgt = b + cloud * M * 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = (int64_t)blockIdx.x * SQPB + warp * (32 * SQ);
  float qx[SQ], qy[SQ], qz[SQ], best[SQ];
#pragma unroll
  for (int k = 0; k < SQ; ++k) {
    int64_t qi = q0 + k * 32 + lane;
    if (qi >= N) qi = N - 1;  // duplicates of a valid query change neither minimum
    qx[k] = qry[qi * 3 + 0];
    qy[k] = qry[qi * 3 + 1];
    qz[k] = qry[qi * 3 + 2];
    best[k] = 3.4e38f;
  }
  for (int64_t t0 = 0; t0 < M; t0 += STILE) {
    const int cnt = static_cast<int>(M - t0 < STILE ? M - t0 : STILE);
    const int cnt32 = (cnt + 31) & ~31;
    __syncthreads();
    float* tf = reinterpret_cast<float*>(tile);
    for (int i = threadIdx.x; i < cnt32 * 4; i += STHREADS) {
      const int pt = i >> 2, c = i & 3;
      tf[i] = (pt < cnt && c < 3) ? tgt[(t0 + pt) * 3 + c] : (pt < cnt ? 0.f : 1e18f);  // padding: a point far away
    }
    for (int i = threadIdx.x; i < cnt32; i += STHREADS) tmin[i] = 0x7F7F7F7Fu;
    __syncthreads();
    for (int g = 0; g < cnt32; g += 32) {
      float pm[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float4 p = tile[g + t];
        float m = 3.4e38f;
#pragma unroll
        for (int k = 0; k < SQ; ++k) {
          const float dx = qx[k] - p.x, dy = qy[k] - p.y, dz = qz[k] - p.z;
          const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          best[k] = fminf(best[k], d);
          m = fminf(m, d);
        }
        pm[t] = m;
      }
      // butterfly transpose-reduce: lane l ends with min over the 32 lanes of pm[l]
#pragma unroll
      for (int r = 16; r >= 1; r >>= 1) {
        const bool upper = (lane & r) != 0;
#pragma unroll
        for (int i = 0; i < r; ++i) {
          const float keep = upper ? pm[i + r] : pm[i];
          const float give = upper ? pm[i] : pm[i + r];
          pm[i] = fminf(keep, __shfl_xor_sync(0xffffffffu, give, r));
        }
      }
      atomicMin(&tmin[g + lane], __float_as_uint(pm[0]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += STHREADS) atomicMin(&d2sq_bits[cloud * M + t0 + i], tmin[i]);
  }
#pragma unroll
  for (int k = 0; k < SQ; ++k) {
    const int64_t qi = q0 + k * 32 + lane;
    if (qi < N) d1[cloud * N + qi] = sqrtf(best[k]);
  }
}

// ---- the same sweep on packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2) and three-input minima (FMNMX3).
// A 64-bit register holds one coordinate of TWO consecutive targets, so the six arithmetic instructions of a pair
// evaluation (3 subtractions, 1 multiply, 2 FMAs) serve two pairs, the query's running minimum takes both with one
// FMNMX3, and a target's minimum over the lane's SQ queries takes SQ / 2 of them: 4.5 issued instructions per pair
// evaluation instead of 8 (+ the shuffle tree).  The kernel is issue-bound, so that is the speed-up to expect.  Each
// lane computes (t - q) instead of (q - t): the squares are the same bits, results are identical to nn_sym_kernel.
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
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int SQ, int SWARPS>
__global__ void __launch_bounds__(SWARPS * 32)
nn_sym2_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t N, int64_t M, float* __restrict__ d1,
               unsigned int* __restrict__ d2sq_bits) {
  // coordinate planes of the target tile: element i of a plane = that coordinate of targets (2i, 2i + 1)
  __shared__ __align__(16) float tx[STILE], ty[STILE], tz[STILE];
  __shared__ unsigned int tmin[STILE];
  constexpr int STHREADS = SWARPS * 32, SQPB = SWARPS * 32 * SQ;
  const int64_t cloud = blockIdx.y;
  const float* qry = a + cloud * N * 3;
  const float* tgt = b + cloud * M * 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = (int64_t)blockIdx.x * SQPB + warp * (32 * SQ);
  f32x2 nqx[SQ], nqy[SQ], nqz[SQ];  // (-q, -q)
  float best[SQ];
#pragma unroll
  for (int k = 0; k < SQ; ++k) {
    int64_t qi = q0 + k * 32 + lane;
    if (qi >= N) qi = N - 1;  // duplicates of a valid query change neither minimum
    const float x = qry[qi * 3 + 0], y = qry[qi * 3 + 1], z = qry[qi * 3 + 2];
    nqx[k] = pk2(-x, -x);
    nqy[k] = pk2(-y, -y);
    nqz[k] = pk2(-z, -z);
    best[k] = 3.4e38f;
  }
  for (int64_t t0 = 0; t0 < M; t0 += STILE) {
    const int cnt = static_cast<int>(M - t0 < STILE ? M - t0 : STILE);
    const int cnt32 = (cnt + 31) & ~31;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt32 * 3; i += STHREADS) {  // coalesced read of the xyz stream
      const int pt = i / 3, c = i - pt * 3;
      const float v = pt < cnt ? tgt[t0 * 3 + i] : 1e18f;  // padding: a point far away
      (c == 0 ? tx : c == 1 ? ty : tz)[pt] = v;
    }
    for (int i = threadIdx.x; i < cnt32; i += STHREADS) tmin[i] = 0x7F7F7F7Fu;
    __syncthreads();
    for (int g = 0; g < cnt32; g += 32) {
      float pm[32];
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        const f32x2 px = *reinterpret_cast<const f32x2*>(tx + g + t);
        const f32x2 py = *reinterpret_cast<const f32x2*>(ty + g + t);
        const f32x2 pz = *reinterpret_cast<const f32x2*>(tz + g + t);
        float lo[SQ], hi[SQ];
#pragma unroll
        for (int k = 0; k < SQ; ++k) {
          const f32x2 dx = add2(px, nqx[k]), dy = add2(py, nqy[k]), dz = add2(pz, nqz[k]);
          const f32x2 d = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
          upk2(d, lo[k], hi[k]);
          best[k] = min3(best[k], lo[k], hi[k]);
        }
        float m0 = lo[0], m1 = hi[0];
#pragma unroll
        for (int k = 1; k + 1 < SQ; k += 2) {
          m0 = min3(m0, lo[k], lo[k + 1]);
          m1 = min3(m1, hi[k], hi[k + 1]);
        }
        if (SQ % 2 == 0) {
          m0 = fminf(m0, lo[SQ - 1]);
          m1 = fminf(m1, hi[SQ - 1]);
        }
        pm[t] = m0;
        pm[t + 1] = m1;
      }
      // butterfly transpose-reduce: lane l ends with min over the 32 lanes of pm[l]
#pragma unroll
      for (int r = 16; r >= 1; r >>= 1) {
        const bool upper = (lane & r) != 0;
#pragma unroll
        for (int i = 0; i < r; ++i) {
          const float keep = upper ? pm[i + r] : pm[i];
          const float give = upper ? pm[i] : pm[i + r];
          pm[i] = fminf(keep, __shfl_xor_sync(0xffffffffu, give, r));
        }
      }
      atomicMin(&tmin[g + lane], __float_as_uint(pm[0]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += STHREADS) atomicMin(&d2sq_bits[cloud * M + t0 + i], tmin[i]);
  }
#pragma unroll
  for (int k = 0; k < SQ; ++k) {
    const int64_t qi = q0 + k * 32 + lane;
    if (qi < N) d1[cloud * N + qi] = sqrtf(best[k]);
  }
}

// ---- the two-sweep kernel WITH indices on packed pairs: a 64-bit register holds one coordinate of TWO queries of the lane
// and the tile stores every target coordinate twice ((x, x, y, y) + (z, z): one LDS.128 + one LDS.64, no register moves
// to build the broadcast), so the six arithmetic instructions serve two pair evaluations; compare + two selects per
// pair as before.  (t - q)^2 instead of (q - t)^2: the same bits, the same arg-min.
__global__ void __launch_bounds__(THREADS)
nn_idx2_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t N, int64_t M, float* __restrict__ d1,
               float* __restrict__ d2, int32_t* __restrict__ idx1, int32_t* __restrict__ idx2) {
  static_assert(Q % 2 == 0, "queries are taken in pairs");
  __shared__ ulonglong2 txy[TILE];
  __shared__ f32x2 tzz[TILE];
  __shared__ float red_d[WARPS][QPB];
  __shared__ int red_i[WARPS][QPB];

  const int dir = blockIdx.z;
  const int64_t cloud = blockIdx.y;
  const float* qry = (dir == 0 ? a + cloud * N * 3 : b + cloud * M * 3);
  const float* tgt = (dir == 0 ? b + cloud * M * 3 : a + cloud * N * 3);
  const int64_t nq = dir == 0 ? N : M, nt = dir == 0 ? M : N;
  float* dout = (dir == 0 ? d1 + cloud * N : d2 + cloud * M);
  int32_t* iout = dir == 0 ? (idx1 ? idx1 + cloud * N : nullptr) : (idx2 ? idx2 + cloud * M : nullptr);

  const int64_t q0 = (int64_t)blockIdx.x * QPB;
  if (q0 >= nq) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  f32x2 nqx[Q / 2], nqy[Q / 2], nqz[Q / 2];  // (-q_2k, -q_2k+1)
  float best[Q];
  int bidx[Q];
#pragma unroll
  for (int k = 0; k < Q; k += 2) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int64_t qi = q0 + (k + h) * 32 + lane;
      if (qi >= nq) qi = nq - 1;  // clamp: duplicates are computed and dropped at the store
      c[h][0] = qry[qi * 3 + 0];
      c[h][1] = qry[qi * 3 + 1];
      c[h][2] = qry[qi * 3 + 2];
      best[k + h] = 3.4e38f;
      bidx[k + h] = 0;
    }
    nqx[k / 2] = pk2(-c[0][0], -c[1][0]);
    nqy[k / 2] = pk2(-c[0][1], -c[1][1]);
    nqz[k / 2] = pk2(-c[0][2], -c[1][2]);
  }

  for (int64_t t0 = 0; t0 < nt; t0 += TILE) {
    const int cnt = static_cast<int>(nt - t0 < TILE ? nt - t0 : TILE);
    __syncthreads();
    float* fxy = reinterpret_cast<float*>(txy);
    float* fz = reinterpret_cast<float*>(tzz);
    for (int i = threadIdx.x; i < cnt * 3; i += THREADS) {  // coalesced read of cnt*3 consecutive floats
      const int pt = i / 3, c = i - pt * 3;
      const float v = tgt[t0 * 3 + i];
      float* dst = c < 2 ? fxy + pt * 4 + 2 * c : fz + pt * 2;
      dst[0] = v;
      dst[1] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = warp; j < cnt; j += WARPS) {
      const ulonglong2 pxy = txy[j];
      const f32x2 pzz = tzz[j];
      const int gj = static_cast<int>(t0) + j;
#pragma unroll
      for (int k = 0; k < Q; k += 2) {
        const f32x2 dx = add2(pxy.x, nqx[k / 2]), dy = add2(pxy.y, nqy[k / 2]), dz = add2(pzz, nqz[k / 2]);
        float d[2];
        upk2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), d[0], d[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (d[h] < best[k + h]) {  // strict: first (lowest) index wins within a warp's stride
            best[k + h] = d[h];
            bidx[k + h] = gj;
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    red_d[warp][k * 32 + lane] = best[k];
    red_i[warp][k * 32 + lane] = bidx[k];
  }
  __syncthreads();
  if (threadIdx.x < QPB) {
    float m = red_d[0][threadIdx.x];
    int mi = red_i[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
      const float v = red_d[w][threadIdx.x];
      const int vi = red_i[w][threadIdx.x];
      if (v < m || (v == m && vi < mi)) {  // ties -> lowest index, like argmin
        m = v;
        mi = vi;
      }
    }
    const int64_t qi = q0 + threadIdx.x;
    if (qi < nq) {
      dout[qi] = sqrtf(m);
      if (iout) iout[qi] = mi;
    }
  }
}

// d2 <- sqrt(d2^2) in place (the buffer holds float bit patterns written by atomicMin)
__global__ void sqrt_inplace_kernel(float* __restrict__ d, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = sqrtf(d[i]);
}

}  // namespace chamfer
}  // namespace nova

static int chamfer_packed() {  // NOVA_B200_CHAMFER_PACKED=0 keeps the scalar one-sweep kernel, =8 takes 8 queries per lane (A/B runs)
  const char* e = std::getenv("NOVA_B200_CHAMFER_PACKED");
  return e == nullptr ? 1 : std::atoi(e);
}
static bool chamfer_one_sweep() {  // NOVA_B200_CHAMFER_SWEEPS=2 keeps the two-sweep distance-only kernel (A/B runs)
  const char* e = std::getenv("NOVA_B200_CHAMFER_SWEEPS");
  return e == nullptr || std::atoi(e) != 2;
}

namespace nova {
namespace chamfer {
// cd[b] = mean_i d1[b, i] + mean_j d2[b, j] in double (chamfer_distance's two np.mean over float64 distances,
// demo.py:50-53): one CTA per pair, fixed-order reduction (deterministic), one launch instead of five ATen kernels
__global__ void __launch_bounds__(256) pair_mean_kernel(const float* __restrict__ d1, const float* __restrict__ d2, int64_t N,
                                                        int64_t M, double* __restrict__ cd) {
  __shared__ double red[2][8];
  const int64_t b = blockIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) s1 += static_cast<double>(d1[b * N + i]);
  for (int64_t j = threadIdx.x; j < M; j += blockDim.x) s2 += static_cast<double>(d2[b * M + j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < 8; ++w) {
      t1 += red[0][w];
      t2 += red[1][w];
    }
    cd[b] = t1 / static_cast<double>(N) + t2 / static_cast<double>(M);
  }
}
}  // namespace chamfer
}  // namespace nova

extern "C" int nova_chamfer_pair_mean(const float* d1, const float* d2, int64_t B, int64_t N, int64_t M, double* cd,
                                      void* stream) {
  using namespace nova;
  NOVA_REQUIRE(d1 && d2 && cd, "nova_chamfer_pair_mean: null pointer");
  NOVA_REQUIRE(B >= 0 && N > 0 && M > 0, "nova_chamfer_pair_mean: empty input (B=%lld N=%lld M=%lld)", (long long)B,
               (long long)N, (long long)M);
  if (B == 0) return NOVA_OK;
  chamfer::pair_mean_kernel<<<(unsigned)B, 256, 0, static_cast<cudaStream_t>(stream)>>>(d1, d2, N, M, cd);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

extern "C" int nova_chamfer_nn(const float* a, const float* b, int64_t B, int64_t N, int64_t M, float* d1, float* d2,
                               int32_t* idx1, int32_t* idx2, void* stream) {
  using namespace nova;
  NOVA_REQUIRE(a && b && d1 && d2, "nova_chamfer_nn: null pointer");
  NOVA_REQUIRE(B >= 0 && N > 0 && M > 0, "nova_chamfer_nn: empty point cloud (B=%lld N=%lld M=%lld); the reference's "
               "np.min over an empty axis raises as well", (long long)B, (long long)N, (long long)M);
  NOVA_REQUIRE(B <= 65535 && N < (1ll << 31) && M < (1ll << 31), "nova_chamfer_nn: batch > 65535 or cloud too large");
  if (B == 0) return NOVA_OK;
  const int64_t big = N > M ? N : M;
  dim3 grid((unsigned)ceil_div(big, chamfer::QPB), (unsigned)B, 2);
  if ((idx1 != nullptr || idx2 != nullptr) && chamfer_packed())
    chamfer::nn_idx2_kernel<<<grid, chamfer::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, b, N, M, d1, d2, idx1, idx2);
  else if (idx1 != nullptr || idx2 != nullptr)
    chamfer::nn_kernel<true><<<grid, chamfer::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, b, N, M, d1, d2, idx1,
                                                                                            idx2);
  else if (chamfer_one_sweep()) {
    // distance only: one sweep over the N x M pairs feeds both directions (3 launches: init, sweep, sqrt)
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NOVA_CHECK_CUDA(cudaMemsetAsync(d2, 0x7F, sizeof(float) * B * M, s));  // 0x7F7F7F7F = 3.39e38
    dim3 sgrid((unsigned)ceil_div(N, chamfer::SQPB), (unsigned)B);
    if (chamfer_packed() == 8) {  // 8 queries per lane: the shuffle tree of a 32-target group serves twice the pairs
      dim3 g8((unsigned)ceil_div(N, 4 * 32 * 8), (unsigned)B);
      chamfer::nn_sym2_kernel<8, 4><<<g8, 128, 0, s>>>(a, b, N, M, d1, reinterpret_cast<unsigned int*>(d2));
    } else if (chamfer_packed())
      chamfer::nn_sym2_kernel<chamfer::SQ, chamfer::SWARPS><<<sgrid, chamfer::STHREADS, 0, s>>>(a, b, N, M, d1, reinterpret_cast<unsigned int*>(d2));
    else
      chamfer::nn_sym_kernel<<<sgrid, chamfer::STHREADS, 0, s>>>(a, b, N, M, d1, reinterpret_cast<unsigned int*>(d2));
    NOVA_CHECK_LAUNCH();
    chamfer::sqrt_inplace_kernel<<<(unsigned)ceil_div(B * M, 256), 256, 0, s>>>(d2, B * M);
  } else
    chamfer::nn_kernel<false><<<grid, chamfer::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, b, N, M, d1, d2,
                                                                                             nullptr, nullptr);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}
