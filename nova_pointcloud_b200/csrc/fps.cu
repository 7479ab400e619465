// Farthest point sampling on the GPU.
//
// The reference's farthest_point_sampling (diffnext/models/transformers/transformer_pointcloud_nova.py:100-125) takes
// `min` over a distance matrix that still contains its zero diagonal, so in exact arithmetic every pick after the
// random start is index 0 (tests/test_oracle_vs_reference.py::test_geometry_live shows it on the live function); what
// its docstring and its caller (adaptive_sampling, :92-97) MEAN is the textbook algorithm, and that is what runs here:
//   picked[0] = start;  d[j] = |p_j - p_start|^2
//   picked[i] = argmax_j d[j]  (lowest index on ties, like torch.argmax);  d[j] = min(d[j], |p_j - p_picked[i]|^2)
// The reference-exact result ([start, 0, 0, ...]) needs no kernel and is offered by the Python mirror as
// mode="reference".  Squared distances are exact fp32 differences, summed left to right WITHOUT fused multiply-add, so
// the numpy float32 oracle reproduces every pick bit for bit.
//
// One CTA per cloud; points (SoA) and running minima live in shared memory; an iteration is one pass over the cloud +
// one block-wide arg-max (warp shuffles, then one warp over the per-warp results).
#include <atomic>

#include "common.cuh"

namespace nova {
namespace fps {

constexpr int THREADS = 512;

__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

static __global__ void __launch_bounds__(THREADS)
fps_kernel(const float* __restrict__ points, const int64_t* __restrict__ start, int64_t N, int S,
           int64_t* __restrict__ picked) {
  extern __shared__ float sm[];  // x[N] | y[N] | z[N] | d[N]
  float *sx = sm, *sy = sm + N, *sz = sm + 2 * N, *sd = sm + 3 * N;
  __shared__ float red_v[THREADS / 32];
  __shared__ int red_i[THREADS / 32];
  __shared__ int cur_s;
  const int64_t b = blockIdx.x;
  const float* p = points + b * N * 3;
  for (int64_t j = threadIdx.x; j < N; j += THREADS) {
    sx[j] = p[j * 3];
    sy[j] = p[j * 3 + 1];
    sz[j] = p[j * 3 + 2];
    sd[j] = __int_as_float(0x7f800000);  // +inf
  }
  if (threadIdx.x == 0) {
    int64_t s0 = start ? start[b] : 0;
    cur_s = static_cast<int>(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
    picked[b * S] = cur_s;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = 1; i < S; ++i) {
    const int c = cur_s;
    const float cx = sx[c], cy = sy[c], cz = sz[c];
    float best = -1.f;
    int best_j = 0x7fffffff;
    for (int64_t j = threadIdx.x; j < N; j += THREADS) {  // ascending j per thread: the first maximum wins
      const float d = fminf(sd[j], dist2(sx[j], sy[j], sz[j], cx, cy, cz));
      sd[j] = d;
      if (d > best) { best = d; best_j = static_cast<int>(j); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
      if (ov > best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
    }
    if (lane == 0) { red_v[warp] = best; red_i[warp] = best_j; }
    __syncthreads();
    if (warp == 0) {
      best = lane < THREADS / 32 ? red_v[lane] : -1.f;
      best_j = lane < THREADS / 32 ? red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ov > best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
      }
      if (lane == 0) {
        cur_s = best_j;
        picked[b * S + i] = best_j;
      }
    }
    __syncthreads();
  }
}

}  // namespace fps
}  // namespace nova

extern "C" int nova_farthest_point_sampling(const float* points, const int64_t* start_idx, int64_t B, int64_t N,
                                            int32_t num_samples, int64_t* picked, void* stream) {
  using namespace nova;
  NOVA_REQUIRE(points && picked, "nova_farthest_point_sampling: null pointer");
  NOVA_REQUIRE(B >= 0 && N > 0 && num_samples >= 1, "nova_farthest_point_sampling: bad sizes (B=%lld N=%lld S=%d)",
               (long long)B, (long long)N, (int)num_samples);
  NOVA_REQUIRE(N <= 14000, "nova_farthest_point_sampling: at most 14000 points per cloud (shared-memory resident)");
  if (B == 0) return NOVA_OK;
  const size_t smem = static_cast<size_t>(N) * 4 * sizeof(float);
  static std::atomic<unsigned long long> done{0ull};
  int dev = 0;
  NOVA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !(done.load() & (1ull << dev))) {
    NOVA_CHECK_CUDA(cudaFuncSetAttribute(fps::fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev);
  }
  fps::fps_kernel<<<(unsigned)B, fps::THREADS, smem, static_cast<cudaStream_t>(stream)>>>(points, start_idx, N,
                                                                                          num_samples, picked);
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}
