// Point-cloud neighbourhood ops that share the Chamfer tiling, sm_100a.
//
//   nova_knn            k smallest Euclidean distances (ascending) + indices of every query point
//   nova_local_density  mean distance to the k nearest neighbours, nearest one (self) dropped
//                       -- compute_local_density, diffnext/models/transformers/transformer_pointcloud_nova.py:81-89
//   nova_softmax_interp out_i = sum_j softmax_j(-|t_i - p_j|) p_j
//                       -- the weighted average of feature_aware_interpolation, same file :146-150
//
// Same arithmetic as chamfer.cu: exact difference form sum((x-y)^2) in fp32 (the reference's torch.cdist
// switches to the |x|^2+|y|^2-2xy matrix form above 25 points and loses ~1e-5 absolute; scipy float64 is
// the pin).  One thread owns one query; the target cloud streams through shared memory as float4 tiles, so
// every LDS.128 is a broadcast.  The k best d^2 live in registers as a sorted list; candidates reach it
// through a per-lane queue in shared memory so that the (divergent) insertion runs warp-wide, see knn_kernel.
// Ties keep the lowest index.
// Bound: fp32 issue rate (HBM traffic is 12 B per query + 12 B per target per CTA against N targets of work).
#include <cstdlib>

#include "common.cuh"

namespace nova {
namespace knn {

constexpr int THREADS = 128;  // queries per CTA
constexpr int TILE = 1024;    // target points per shared-memory tile

__device__ __forceinline__ void stage_tile(float4* tile, const float* __restrict__ tgt, int64_t t0, int cnt) {
  float* tf = reinterpret_cast<float*>(tile);
  for (int i = threadIdx.x; i < cnt * 3; i += THREADS) {  // coalesced read of cnt*3 consecutive floats
    const int pt = i / 3, c = i - pt * 3;
    tf[pt * 4 + c] = tgt[t0 * 3 + i];
  }
}

// ---- packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): the two queries of a thread share every arithmetic instruction.
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
// Packed form of the tile: every coordinate twice, (x, x, y, y) + (z, z), so that one LDS.128 + one LDS.64 put a
// target into 64-bit registers that pair with (query 0, query 1) -- no register moves to build the broadcast.
constexpr int TILE2 = 512;  // 24 B per target
__device__ __forceinline__ void stage_tile2(ulonglong2* txy, f32x2* tzz, const float* __restrict__ tgt, int64_t t0, int cnt) {
  float* fxy = reinterpret_cast<float*>(txy);
  float* fz = reinterpret_cast<float*>(tzz);
  for (int i = threadIdx.x; i < cnt * 3; i += THREADS) {
    const int pt = i / 3, c = i - pt * 3;
    const float v = tgt[t0 * 3 + i];
    float* dst = c < 2 ? fxy + pt * 4 + 2 * c : fz + pt * 2;
    dst[0] = v;
    dst[1] = v;
  }
}

// MODE 0: write the k distances (+ indices); MODE 1: density = mean(sqrt(best[1..k-1])) (k = k_neighbors + 1).
//
// QPT queries per thread share every LDS.128 (one query per thread is shared-memory-issue bound: one broadcast
// load per 32 pair evaluations).
//
// Deferred insertion.  A lane inserts ~k (1 + ln(N/k)) times over N targets, but a warp executes the insertion
// code whenever ANY of its lanes does: for k = 9, N = 2048 that is ~850 of the 2048 iterations at ~45
// instructions each -- three times the distance loop itself.  So a candidate that beats the lane's current
// k-th distance is only PUSHED (one or two STS) onto the lane's queue in shared memory; when some lane's
// queue is nearly full the whole warp drains, every lane inserting its own pending candidates, so the
// insertion code runs with most lanes active.  The threshold is stale between drains (a few extra pushes);
// candidates are drained in index order and the insertion compare is strict, so ties keep the lowest index.
// In MODE 1 no index travels with a distance and the sorted insertion is a branch-free min/max chain.
constexpr int QCAP = 8;  // queue entries per query
constexpr int GRP = 4;   // targets between two queue checks (tile padded to a multiple with NaN points)

// PACKED (QPT == 2 only): the distance arithmetic of the thread's two queries on packed fp32 pairs; (t - q)^2 instead of
// (q - t)^2, the same bits.
template <int KCAP, int MODE, int QPT, bool PACKED = false>
__global__ void __launch_bounds__(THREADS)
knn_kernel(const float* __restrict__ q, const float* __restrict__ t, int64_t Nq, int64_t Nt, int k,
           float* __restrict__ dist, int32_t* __restrict__ idx) {
  static_assert(!PACKED || QPT == 2, "the packed form pairs the two queries of a thread");
  constexpr int TILE_PTS = PACKED ? TILE2 : TILE;
  __shared__ __align__(16) float4 tile[PACKED ? TILE2 * 3 / 2 : TILE];  // PACKED: TILE2 x (x,x,y,y) then TILE2 x (z,z)
  ulonglong2* txy = reinterpret_cast<ulonglong2*>(tile);
  f32x2* tzz = reinterpret_cast<f32x2*>(tile + TILE2);
  __shared__ float qd[QPT][QCAP][THREADS];
  __shared__ int qj[MODE == 0 ? QPT : 1][QCAP][THREADS];
  const int64_t cloud = blockIdx.y;
  const float* qry = q + cloud * Nq * 3;
  const float* tgt = t + cloud * Nt * 3;
  const int tid = threadIdx.x;
  const int64_t q0 = (int64_t)blockIdx.x * (THREADS * QPT) + tid;  // query u of this thread: q0 + u * THREADS

  float qx[QPT], qy[QPT], qz[QPT];
  float best[QPT][KCAP];
  int bidx[QPT][KCAP];
  int pending[QPT];
#pragma unroll
  for (int u = 0; u < QPT; ++u) {
    const int64_t qi = q0 + u * THREADS;
    const int64_t ql = qi < Nq ? qi : Nq - 1;  // clamped lanes compute a duplicate and skip the store
    qx[u] = qry[ql * 3 + 0];
    qy[u] = qry[ql * 3 + 1];
    qz[u] = qry[ql * 3 + 2];
    pending[u] = 0;
#pragma unroll
    for (int i = 0; i < KCAP; ++i) {
      best[u][i] = 3.4e38f;
      bidx[u][i] = 0;
    }
  }

  auto drain = [&]() {
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
#pragma unroll 1
      for (int e = 0; e < QCAP; ++e) {
        if (e < pending[u]) {
          const float d = qd[u][e][tid];
          if (d < best[u][KCAP - 1]) {  // strict: an equal later index never displaces an earlier one
            if (MODE == 1) {
              float v = d;
#pragma unroll
              for (int i = 0; i < KCAP; ++i) {
                const float lo = fminf(best[u][i], v);
                v = fmaxf(best[u][i], v);
                best[u][i] = lo;
              }
            } else {
              best[u][KCAP - 1] = d;
              bidx[u][KCAP - 1] = qj[u][e][tid];
#pragma unroll
              for (int i = KCAP - 1; i > 0; --i) {
                if (best[u][i] < best[u][i - 1]) {
                  const float fd = best[u][i];
                  best[u][i] = best[u][i - 1];
                  best[u][i - 1] = fd;
                  const int fi = bidx[u][i];
                  bidx[u][i] = bidx[u][i - 1];
                  bidx[u][i - 1] = fi;
                }
              }
            }
          }
        }
      }
      pending[u] = 0;
    }
  };

  f32x2 nqx2 = 0, nqy2 = 0, nqz2 = 0;  // PACKED: (-q0, -q1) per coordinate
  if (PACKED) {
    nqx2 = pk2(-qx[0], -qx[QPT - 1]);
    nqy2 = pk2(-qy[0], -qy[QPT - 1]);
    nqz2 = pk2(-qz[0], -qz[QPT - 1]);
  }
  for (int64_t t0 = 0; t0 < Nt; t0 += TILE_PTS) {
    const int cnt = static_cast<int>(Nt - t0 < TILE_PTS ? Nt - t0 : TILE_PTS);
    const int cntp = (cnt + GRP - 1) / GRP * GRP;
    __syncthreads();
    if (PACKED) {
      stage_tile2(txy, tzz, tgt, t0, cnt);
      if (tid < cntp - cnt) txy[cnt + tid] = make_ulonglong2(pk2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000)), 0ull);  // NaN: never < anything
      if (tid < cntp - cnt) tzz[cnt + tid] = 0ull;
    } else {
      stage_tile(tile, tgt, t0, cnt);
      if (tid < cntp - cnt) tile[cnt + tid] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);  // NaN: never < anything
    }
    __syncthreads();
    for (int j0 = 0; j0 < cntp; j0 += GRP) {
#pragma unroll
      for (int g = 0; g < GRP; ++g) {
        float dd[QPT];
        if (PACKED) {
          const ulonglong2 pxy = txy[j0 + g];
          const f32x2 pzz = tzz[j0 + g];
          const f32x2 dx = add2(pxy.x, nqx2), dy = add2(pxy.y, nqy2), dz = add2(pzz, nqz2);
          upk2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), dd[0], dd[QPT - 1]);
        } else {
          const float4 p = tile[j0 + g];
#pragma unroll
          for (int u = 0; u < QPT; ++u) {
            const float dx = qx[u] - p.x, dy = qy[u] - p.y, dz = qz[u] - p.z;
            dd[u] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          }
        }
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
          const float d = dd[u];
          if (d < best[u][KCAP - 1]) {
            qd[u][pending[u]][tid] = d;
            if (MODE == 0) qj[u][pending[u]][tid] = static_cast<int>(t0) + j0 + g;
            ++pending[u];
          }
        }
      }
      bool nearly_full = false;
#pragma unroll
      for (int u = 0; u < QPT; ++u) nearly_full |= pending[u] > QCAP - GRP;
      if (__any_sync(0xffffffffu, nearly_full)) drain();
    }
  }
  drain();
#pragma unroll
  for (int u = 0; u < QPT; ++u) {
    const int64_t qi = q0 + u * THREADS;
    if (qi >= Nq) continue;
    if (MODE == 0) {
      float* dq = dist + (cloud * Nq + qi) * k;
      int32_t* iq = idx ? idx + (cloud * Nq + qi) * k : nullptr;
#pragma unroll
      for (int i = 0; i < KCAP; ++i)
        if (i < k) {
          dq[i] = sqrtf(best[u][i]);
          if (iq) iq[i] = bidx[u][i];
        }
    } else {
      float s = 0.f;  // ascending order, like mean() over topk(...)[..., 1:]
#pragma unroll
      for (int i = 1; i < KCAP; ++i)
        if (i < k) s += sqrtf(best[u][i]);
      dist[cloud * Nq + qi] = s / static_cast<float>(k - 1);
    }
  }
}

// out_i = sum_j w_ij p_j / sum_j w_ij with w_ij = exp(-(|t_i - p_j| - m_i)); m_i is the running minimum distance
// (online softmax: when a nearer source appears the accumulators are rescaled once).  One exp + one sqrt per
// pair: the kernel is MUFU-bound (2 special-function ops per pair against ~12 FMA-pipe instructions).
__global__ void __launch_bounds__(THREADS)
softmax_interp_kernel(const float* __restrict__ q, const float* __restrict__ t, int64_t Nq, int64_t Nt,
                      float* __restrict__ out) {
  __shared__ float4 tile[TILE];
  const int64_t cloud = blockIdx.y;
  const float* qry = q + cloud * Nq * 3;
  const float* tgt = t + cloud * Nt * 3;
  const int64_t qi = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  const int64_t ql = qi < Nq ? qi : Nq - 1;
  const float qx = qry[ql * 3 + 0], qy = qry[ql * 3 + 1], qz = qry[ql * 3 + 2];
  float m = 3.4e38f, s = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
  for (int64_t t0 = 0; t0 < Nt; t0 += TILE) {
    const int cnt = static_cast<int>(Nt - t0 < TILE ? Nt - t0 : TILE);
    __syncthreads();
    stage_tile(tile, tgt, t0, cnt);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 p = tile[j];
      const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
      const float d = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
      if (d < m) {  // new nearest source: rescale what has been accumulated (exp(-inf) = 0 the first time)
        const float r = expf(d - m);
        s *= r;
        ax *= r;
        ay *= r;
        az *= r;
        m = d;
      }
      const float w = expf(m - d);
      s += w;
      ax = fmaf(w, p.x, ax);
      ay = fmaf(w, p.y, ay);
      az = fmaf(w, p.z, az);
    }
  }
  if (qi >= Nq) return;
  const float inv = 1.0f / s;
  float* o = out + (cloud * Nq + qi) * 3;
  o[0] = ax * inv;
  o[1] = ay * inv;
  o[2] = az * inv;
}

// The same on two sources per iteration: the tile as three coordinate planes, one LDS.64 per plane puts sources (j, j + 1)
// into 64-bit registers, the distance arithmetic runs on packed fp32 pairs ((p - t)^2, the same bits as (t - p)^2), and
// the weights use ex2.approx on x log2(e) (<= 2 ulp + |x| 6e-8 relative: the sources that carry weight have |x| of a few
// units) instead of the ~11-instruction expf.  The square root stays correctly rounded: its error is the one that the
// far-from-origin case multiplies by the distance.  Sources are still visited in index order (lo, then hi).
__device__ __forceinline__ float exp_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
__global__ void __launch_bounds__(THREADS)
softmax_interp2_kernel(const float* __restrict__ q, const float* __restrict__ t, int64_t Nq, int64_t Nt,
                       float* __restrict__ out) {
  __shared__ __align__(16) float tx[TILE], ty[TILE], tz[TILE];
  const int64_t cloud = blockIdx.y;
  const float* qry = q + cloud * Nq * 3;
  const float* tgt = t + cloud * Nt * 3;
  const int64_t qi = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  const int64_t ql = qi < Nq ? qi : Nq - 1;
  const float qx = qry[ql * 3 + 0], qy = qry[ql * 3 + 1], qz = qry[ql * 3 + 2];
  const f32x2 nqx = pk2(-qx, -qx), nqy = pk2(-qy, -qy), nqz = pk2(-qz, -qz);
  float m = 3.4e38f, s = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
  for (int64_t t0 = 0; t0 < Nt; t0 += TILE) {
    const int cnt = static_cast<int>(Nt - t0 < TILE ? Nt - t0 : TILE);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 3; i += THREADS) {  // coalesced read of the xyz stream
      const int pt = i / 3, c = i - pt * 3;
      (c == 0 ? tx : c == 1 ? ty : tz)[pt] = tgt[t0 * 3 + i];
    }
    if ((cnt & 1) && threadIdx.x == 0) tx[cnt] = ty[cnt] = tz[cnt] = 1e18f;  // pad source: weight exp(-1e18) = 0
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < cnt; j += 2) {
      const f32x2 px = *reinterpret_cast<const f32x2*>(tx + j), py = *reinterpret_cast<const f32x2*>(ty + j);
      const f32x2 pz = *reinterpret_cast<const f32x2*>(tz + j);
      const f32x2 dx = add2(px, nqx), dy = add2(py, nqy), dz = add2(pz, nqz);
      float d2[2], cx[2], cy[2], cz[2];
      upk2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), d2[0], d2[1]);
      upk2(px, cx[0], cx[1]);
      upk2(py, cy[0], cy[1]);
      upk2(pz, cz[0], cz[1]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float d = sqrtf(d2[h]);
        if (d < m) {  // new nearest source: rescale what has been accumulated (exp(-inf) = 0 the first time)
          const float r = exp_fast(d - m);
          s *= r;
          ax *= r;
          ay *= r;
          az *= r;
          m = d;
        }
        const float w = exp_fast(m - d);
        s += w;
        ax = fmaf(w, cx[h], ax);
        ay = fmaf(w, cy[h], ay);
        az = fmaf(w, cz[h], az);
      }
    }
  }
  if (qi >= Nq) return;
  const float inv = 1.0f / s;
  float* o = out + (cloud * Nq + qi) * 3;
  o[0] = ax * inv;
  o[1] = ay * inv;
  o[2] = az * inv;
}

static bool knn_packed() {  // NOVA_B200_KNN_PACKED=0 keeps the scalar arithmetic for two queries per thread (A/B runs)
  const char* e = std::getenv("NOVA_B200_KNN_PACKED");
  return e == nullptr || std::atoi(e) != 0;
}
template <int KCAP, int MODE, int QPT>
static void launch_one(const float* q, const float* t, int64_t B, int64_t Nq, int64_t Nt, int k, float* dist,
                       int32_t* idx, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(Nq, THREADS * QPT), (unsigned)B);
  if constexpr (QPT == 2 && KCAP <= 9) {  // measured: -4 % (density), -6 % (k = 4), -1 % (k = 9); +0.7 % at k = 16 (insertion-bound)
    if (knn_packed()) {
      knn_kernel<KCAP, MODE, QPT, true><<<grid, THREADS, 0, s>>>(q, t, Nq, Nt, k, dist, idx);
      return;
    }
  }
  knn_kernel<KCAP, MODE, QPT, false><<<grid, THREADS, 0, s>>>(q, t, Nq, Nt, k, dist, idx);
}

template <int MODE>
static int launch_knn(const float* q, const float* t, int64_t B, int64_t Nq, int64_t Nt, int k, float* dist,
                      int32_t* idx, cudaStream_t s) {
  // two queries per thread once the grid still covers the chip (148 SMs x 16 resident CTAs of 128 threads)
  const bool wide = B * ceil_div(Nq, THREADS * 2) >= 148 * 4;
  if (k <= 4)
    wide ? launch_one<4, MODE, 2>(q, t, B, Nq, Nt, k, dist, idx, s) : launch_one<4, MODE, 1>(q, t, B, Nq, Nt, k, dist, idx, s);
  else if (k <= 9)
    wide ? launch_one<9, MODE, 2>(q, t, B, Nq, Nt, k, dist, idx, s) : launch_one<9, MODE, 1>(q, t, B, Nq, Nt, k, dist, idx, s);
  else if (k <= 16)
    wide ? launch_one<16, MODE, 2>(q, t, B, Nq, Nt, k, dist, idx, s) : launch_one<16, MODE, 1>(q, t, B, Nq, Nt, k, dist, idx, s);
  else
    launch_one<32, MODE, 1>(q, t, B, Nq, Nt, k, dist, idx, s);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

}  // namespace knn
}  // namespace nova

extern "C" int nova_knn(const float* q, const float* t, int64_t B, int64_t Nq, int64_t Nt, int32_t k, float* dist,
                        int32_t* idx, void* stream) {
  using namespace nova;
  NOVA_REQUIRE(q && t && dist, "nova_knn: null pointer");
  NOVA_REQUIRE(B >= 0 && Nq > 0 && Nt > 0, "nova_knn: empty point cloud (B=%lld Nq=%lld Nt=%lld)", (long long)B,
               (long long)Nq, (long long)Nt);
  NOVA_REQUIRE(k >= 1 && k <= 32, "nova_knn: k = %d outside [1, 32]", (int)k);
  NOVA_REQUIRE(k <= Nt, "nova_knn: k = %d exceeds the %lld target points (torch.topk raises: selected index k out of "
               "range)", (int)k, (long long)Nt);
  NOVA_REQUIRE(B <= 65535 && Nq < (1ll << 31) && Nt < (1ll << 31), "nova_knn: batch > 65535 or cloud too large");
  if (B == 0) return NOVA_OK;
  return knn::launch_knn<0>(q, t, B, Nq, Nt, k, dist, idx, static_cast<cudaStream_t>(stream));
}

extern "C" int nova_local_density(const float* points, int64_t B, int64_t N, int32_t k_neighbors, float* density,
                                  void* stream) {
  using namespace nova;
  NOVA_REQUIRE(points && density, "nova_local_density: null pointer");
  NOVA_REQUIRE(B >= 0 && N > 0, "nova_local_density: empty point cloud (B=%lld N=%lld)", (long long)B, (long long)N);
  NOVA_REQUIRE(k_neighbors >= 1 && k_neighbors <= 31, "nova_local_density: k_neighbors = %d outside [1, 31]",
               (int)k_neighbors);
  NOVA_REQUIRE(k_neighbors + 1 <= N, "nova_local_density: k_neighbors + 1 = %d exceeds the %lld points (torch.topk "
               "raises: selected index k out of range)", (int)k_neighbors + 1, (long long)N);
  NOVA_REQUIRE(B <= 65535 && N < (1ll << 31), "nova_local_density: batch > 65535 or cloud too large");
  if (B == 0) return NOVA_OK;
  return knn::launch_knn<1>(points, points, B, N, N, k_neighbors + 1, density, nullptr,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int nova_softmax_interp(const float* targets, const float* points, int64_t B, int64_t S, int64_t N,
                                   float* out, void* stream) {
  using namespace nova;
  NOVA_REQUIRE(targets && points && out, "nova_softmax_interp: null pointer");
  NOVA_REQUIRE(B >= 0 && S > 0 && N > 0, "nova_softmax_interp: empty point cloud (B=%lld S=%lld N=%lld)",
               (long long)B, (long long)S, (long long)N);
  NOVA_REQUIRE(B <= 65535 && S < (1ll << 31) && N < (1ll << 31), "nova_softmax_interp: batch > 65535 or cloud too large");
  if (B == 0) return NOVA_OK;
  dim3 grid((unsigned)ceil_div(S, knn::THREADS), (unsigned)B);
  const char* env = std::getenv("NOVA_B200_INTERP_FAST");  // =0: scalar arithmetic and expf (A/B runs)
  if (env == nullptr || std::atoi(env) != 0)
    knn::softmax_interp2_kernel<<<grid, knn::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(targets, points, S, N, out);
  else
    knn::softmax_interp_kernel<<<grid, knn::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(targets, points, S, N, out);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}
