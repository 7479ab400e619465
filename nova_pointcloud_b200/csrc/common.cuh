// Shared helpers for libnova_b200: status/error plumbing, element-type traits,
// 8-wide vector load/store, warp reductions.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/nova_b200.h"

namespace nova {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// In-situ kernel timing (nova_profile_*): CUDA events recorded around launches of one kernel class on
// the launching stream, so bench.py can report per-kernel durations measured inside the real step.
enum KernelClass : int { KC_GEMM_ADA = 0, KC_GEMM_FC = 1, KC_ROW = 2, KC_PREP = 3, KC_OTHER = 4, KC_CHAIN = 5, KC_GEMM_TAIL = 6, KC_COUNT = 7 };
bool profile_enabled();
void profile_begin(int kernel_class, cudaStream_t stream);
void profile_end(cudaStream_t stream);
struct ProfileScope {
  cudaStream_t s;
  bool on;
  ProfileScope(int kernel_class, cudaStream_t stream) : s(stream), on(profile_enabled()) {
    if (on) profile_begin(kernel_class, s);
  }
  ~ProfileScope() {
    if (on) profile_end(s);
  }
};

#define NOVA_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::nova::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NOVA_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define NOVA_CHECK_LAUNCH()                                                                    \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      ::nova::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NOVA_ERR_CUDA;                                                                    \
    }                                                                                          \
    ::nova::count_launch();                                                                    \
  } while (0)

#define NOVA_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::nova::set_error(__VA_ARGS__);    \
      return NOVA_ERR_INVALID;           \
    }                                    \
  } while (0)

#define NOVA_PROPAGATE(expr)      \
  do {                            \
    int _s = (expr);              \
    if (_s != NOVA_OK) return _s; \
  } while (0)

using bf16 = __nv_bfloat16;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- element access: 8 consecutive elements <-> float[8] -------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
// 8 bf16 kept packed (4 registers) between the load and the use
__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

// Fast SiLU for bf16 outputs: x*sigmoid(x) = 0.5x + 0.5x*tanh(0.5x); one MUFU op (tanh.approx,
// ~2^-11 relative error, below bf16's 2^-9 rounding) instead of ex2 + rcp -- the epilogue of the
// fc1 GEMM is MUFU-bound on B200 (16 MUFU/clk/SM).
__device__ __forceinline__ float silu(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// fp32 parity mode uses the accurate exponential.
__device__ __forceinline__ float silu_accurate(float x) { return x / (1.0f + expf(-x)); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------
// Every kernel of the sampling step starts with pdl_trigger() (lets the NEXT kernel's CTAs be scheduled and
// run their prologue as soon as resources free up) and calls pdl_wait() before its first access to global
// memory (blocks until every preceding kernel has completed and flushed).  Both are no-ops for a kernel
// launched without the attribute.  Saves the launch latency + prologue at each of the ~700 kernel
// boundaries of a sampling pass; captured into the loop's CUDA graph as programmatic edges.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();                   // decided per call from the row count (runtime.cu); NOVA_B200_PDL forces it
void pdl_set_for_rows(int64_t rows);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace nova
