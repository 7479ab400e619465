// Library runtime: error string, launch counter, device check, TMA tensor-map encoding,
// the host-mapped debug words of the tcgen05 kernels, and the small C-ABI entry points.
#include <cuda.h>

#include <atomic>
#include <unordered_map>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_api.cuh"
#include "chain_api.cuh"
#include "rowwise.cuh"

namespace nova {

static thread_local char g_error[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
// PDL is a per-call decision (head.cu): it pays at small M, where the ~28 kernels of a diffusion step are
// latency-bound (4-5 % measured), and costs 2-3 % at M = 65 536, where early-scheduled dependents sit next
// to the persistent GEMM CTAs.  NOVA_B200_PDL=0 / 1 forces it off / on.
static thread_local bool g_pdl_active = false;
int pdl_forced() {
  static const int forced = [] {
    const char* e = std::getenv("NOVA_B200_PDL");
    return e == nullptr ? -1 : (std::atoi(e) != 0 ? 1 : 0);
  }();
  return forced;
}
void pdl_set_for_rows(int64_t rows) { g_pdl_active = pdl_forced() >= 0 ? pdl_forced() == 1 : rows <= 16384; }
bool pdl_enabled() { return g_pdl_active; }

namespace {
struct ProfileSpan {
  int kernel_class;
  cudaEvent_t start, stop;
};
thread_local bool g_profile_on = false;
thread_local std::vector<ProfileSpan> g_spans;
}  // namespace

bool profile_enabled() { return g_profile_on; }
void profile_begin(int kernel_class, cudaStream_t stream) {
  ProfileSpan sp{kernel_class, nullptr, nullptr};
  if (cudaEventCreate(&sp.start) != cudaSuccess || cudaEventCreate(&sp.stop) != cudaSuccess) return;
  cudaEventRecord(sp.start, stream);
  g_spans.push_back(sp);
}
void profile_end(cudaStream_t stream) {
  if (!g_spans.empty()) cudaEventRecord(g_spans.back().stop, stream);
}

namespace tc {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

namespace {
// Encoded tensor maps are pure functions of (pointer, shape, stride, box): a launch-heavy eager pass (training step,
// first pass of a sampling call before its graph exists) re-creates the same few dozen maps over and over, so they
// are kept per host thread (no locking; the map is a 128-byte POD that is copied into the kernel parameters anyway).
struct TmapKey {
  const void* ptr;
  int64_t rows, K, ld;
  int box_rows;
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && rows == o.rows && K == o.K && ld == o.ld && box_rows == o.box_rows; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    for (int64_t v : {k.rows, k.K, k.ld, static_cast<int64_t>(k.box_rows)}) h = h * 1000003u ^ std::hash<int64_t>()(v);
    return h;
  }
};
thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t TMAP_CACHE_MAX = 4096;
}  // namespace

int make_tmap_kmajor(CUtensorMap* map, const bf16* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  const TmapKey key{ptr, rows, K, ld, box_rows};
  auto hit = g_tmap_cache.find(key);
  if (hit != g_tmap_cache.end()) {
    *map = hit->second;
    return NOVA_OK;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return NOVA_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * sizeof(bf16)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld K=%lld ld=%lld", (int)r, (long long)rows,
              (long long)K, (long long)ld);
    return NOVA_ERR_CUDA;
  }
  if (g_tmap_cache.size() >= TMAP_CACHE_MAX) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *map);
  return NOVA_OK;
}

uint32_t* debug_word() {
  static uint32_t* dev = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    uint32_t* host = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&host), 4 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      std::memset(host, 0, 4 * sizeof(uint32_t));
      if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev), host, 0) != cudaSuccess) dev = nullptr;
      g_debug_host = host;
    }
  });
  return dev;
}
uint32_t* g_debug_host = nullptr;

int num_sms() {  // of the CURRENT device: one process may drive several GPUs
  static std::atomic<int> sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  int v = sms[slot].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    sms[slot].store(v, std::memory_order_relaxed);
  }
  return v;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to a function ON A DEVICE: set it once per (kernel, device),
// from any host thread (setting it twice is harmless, so a relaxed flag per device is enough).
int ensure_smem_attr(const void* func, int bytes, std::atomic<unsigned long long>* done_mask) {
  int dev = 0;
  NOVA_CHECK_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = dev >= 0 && dev < 64 ? 1ull << dev : 0ull;
  if (bit == 0ull || (done_mask->load(std::memory_order_acquire) & bit) == 0ull) {
    NOVA_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (bit) done_mask->fetch_or(bit, std::memory_order_release);
  }
  return NOVA_OK;
}

int default_cta_group(int M) {
  static int forced = -1;
  if (forced < 0) {
    const char* env = std::getenv("NOVA_B200_CTA_GROUP");
    forced = env ? std::atoi(env) : 0;
    if (forced != 1 && forced != 2) forced = 0;
  }
  if (forced) return forced;
  return M > BM ? 2 : 1;  // a CTA pair needs more than one 128-row slab to be worth it
}

int pdl_epi_mask() {
  static const int m = [] { const char* e = std::getenv("NOVA_B200_PDL_EPI_MASK"); return e ? std::atoi(e) : 7; }();  // default: bias, SiLU, AdaLN kinds; not the tail
  return m;
}

bool fixed_column_grid() {  // NOVA_B200_FIXED_N=1: GEMM grids rounded down to a multiple of the column tiles (launch_epi)
  const char* env = std::getenv("NOVA_B200_FIXED_N");  // read per call
  return env != nullptr && std::atoi(env) != 0;
}

int tile_columns_override() {
  const char* env = std::getenv("NOVA_B200_TILE_N");  // read per call: tests switch it between launches
  const int v = env ? std::atoi(env) : 0;
  return (v == 64 || v == 128 || v == 256) ? v : 0;
}

}  // namespace tc
namespace rw {
int num_sms_rw() { return tc::num_sms(); }
}  // namespace rw
}  // namespace nova

using namespace nova;

extern "C" const char* nova_last_error(void) { return g_error; }
extern "C" int nova_abi_version(void) { return NOVA_B200_ABI_VERSION; }
extern "C" int64_t nova_launch_count(void) { return g_launches; }
extern "C" void nova_launch_count_reset(void) { g_launches = 0; }

extern "C" int nova_profile_enable(int32_t on) {
  g_profile_on = on != 0;
  return NOVA_OK;
}

// Sum the recorded spans per kernel class (ms) and clear them.  Synchronises on the recorded events.
extern "C" int nova_profile_read(double* ms_by_class, int64_t* launches_by_class, int32_t n_classes) {
  NOVA_REQUIRE(ms_by_class && launches_by_class && n_classes >= KC_COUNT, "nova_profile_read: need %d classes", KC_COUNT);
  for (int i = 0; i < n_classes; ++i) {
    ms_by_class[i] = 0.0;
    launches_by_class[i] = 0;
  }
  int rc = NOVA_OK;
  for (ProfileSpan& sp : g_spans) {
    float ms = 0.f;
    if (cudaEventSynchronize(sp.stop) == cudaSuccess && cudaEventElapsedTime(&ms, sp.start, sp.stop) == cudaSuccess) {
      ms_by_class[sp.kernel_class] += ms;
      launches_by_class[sp.kernel_class] += 1;
    } else {
      rc = NOVA_ERR_CUDA;
    }
    cudaEventDestroy(sp.start);
    cudaEventDestroy(sp.stop);
  }
  g_spans.clear();
  if (rc != NOVA_OK) set_error("nova_profile_read: an event could not be read");
  return rc;
}

extern "C" int nova_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  NOVA_CHECK_CUDA(cudaGetDevice(&dev));
  NOVA_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  NOVA_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("libnova_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return NOVA_ERR_DEVICE;
  }
  return NOVA_OK;
}

// Debug words of the last tcgen05 barrier timeout: [0] = 0xDEAD0000 | code, [1] = block, [2] = parity.
extern "C" int nova_debug_words(uint32_t* out4) {
  if (!out4) return NOVA_ERR_INVALID;
  for (int i = 0; i < 4; ++i) out4[i] = tc::g_debug_host ? tc::g_debug_host[i] : 0u;
  return NOVA_OK;
}

extern "C" int nova_debug_words_clear(void) {
  tc::debug_word();  // make sure the words exist
  if (tc::g_debug_host)
    for (int i = 0; i < 4; ++i) tc::g_debug_host[i] = 0u;
  return NOVA_OK;
}

extern "C" int nova_euler_step(const void* model_output, const void* sample, double dt, void* prev, int64_t numel,
                               int32_t dtype, void* stream) {
  NOVA_REQUIRE(model_output && sample && prev && numel >= 0, "nova_euler_step: bad arguments");
  if (numel == 0) return NOVA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)ceil_div(numel, 256);
  // dt is a Python float in the reference; torch multiplies a tensor by a Python scalar after
  // converting the scalar to the tensor's opmath type (fp32 for both fp32 and bf16 tensors).
  const float dtf = static_cast<float>(dt);
  if (dtype == NOVA_F32)
    rw::euler_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(model_output),
                                                 static_cast<const float*>(sample), static_cast<float*>(prev), numel, dtf);
  else if (dtype == NOVA_BF16)
    rw::euler_kernel<bf16><<<grid, 256, 0, s>>>(static_cast<const bf16*>(model_output), static_cast<const bf16*>(sample),
                                                static_cast<bf16*>(prev), numel, dtf);
  else {
    set_error("nova_euler_step: unknown dtype %d", dtype);
    return NOVA_ERR_INVALID;
  }
  NOVA_CHECK_LAUNCH();
  return NOVA_OK;
}

extern "C" int nova_debug_gemm(const void* A, const void* W, const float* bias, void* C, int64_t M, int64_t N, int64_t K,
                               int32_t dtype, int32_t impl, int32_t epilogue, void* stream) {
  NOVA_REQUIRE(A && W && C && M >= 0 && N >= 0 && K > 0, "nova_debug_gemm: bad arguments");
  NOVA_REQUIRE(epilogue == EPI_BIAS || epilogue == EPI_BIAS_SILU, "nova_debug_gemm: unknown epilogue %d", epilogue);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == NOVA_F32) {
    NOVA_REQUIRE(impl == 0, "nova_debug_gemm: fp32 runs on the SIMT kernel only");
    return simt::launch<float, float, true>(static_cast<const float*>(A), K, static_cast<const float*>(W), K, bias,
                                            static_cast<float*>(C), N, (int)M, (int)N, (int)K, epilogue, s);
  }
  NOVA_REQUIRE(dtype == NOVA_BF16, "nova_debug_gemm: unknown dtype %d", dtype);
  if (impl == 0)
    return simt::launch<bf16, bf16, false>(static_cast<const bf16*>(A), K, static_cast<const bf16*>(W), K, bias,
                                           static_cast<bf16*>(C), N, (int)M, (int)N, (int)K, epilogue, s);
  NOVA_REQUIRE(impl >= 1 && impl <= 3, "nova_debug_gemm: unknown impl %d", impl);
  NOVA_PROPAGATE(nova_device_check());
  // impl 1 = tcgen05 cta_group::1, 2 = tcgen05 cta_group::2 (CTA pairs), 3 = library default
  return tc::launch(static_cast<const bf16*>(A), K, static_cast<const bf16*>(W), K, bias, static_cast<bf16*>(C), N,
                    (int)M, (int)N, (int)K, epilogue, s, impl == 3 ? 0 : impl);
}

// Test hook for the fused AdaLN GEMM: W [n_stats * D, K] and bias [n_stats * D] in the reference order
// (scale | shift | gate); packs them, computes the row statistics of x and runs the EPI_ADALN kernel.
extern "C" int nova_debug_adaln_gemm(const void* A, const void* W, const float* bias, const void* x, void* h_out,
                                     void* gate_out, int64_t M, int64_t D, int64_t K, int32_t n_stats,
                                     int32_t cta_group, void* stream) {
  NOVA_REQUIRE(A && W && bias && x && h_out && (n_stats == 2 || (n_stats == 3 && gate_out)),
               "nova_debug_adaln_gemm: bad arguments");
  NOVA_REQUIRE(D > 0 && D % 256 == 0 && K % 8 == 0 && M >= 0, "nova_debug_adaln_gemm: D must be a multiple of 256");
  NOVA_PROPAGATE(nova_device_check());
  if (M == 0) return NOVA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t rows = (int64_t)n_stats * D;
  bf16* w_il = nullptr;
  float *b_il = nullptr, *rstat = nullptr;
  NOVA_CHECK_CUDA(cudaMalloc(&w_il, rows * K * sizeof(bf16)));
  NOVA_CHECK_CUDA(cudaMalloc(&b_il, rows * sizeof(float)));
  NOVA_CHECK_CUDA(cudaMalloc(&rstat, M * 2 * sizeof(float)));
  rw::pack_adaln_kernel<bf16, bf16><<<(unsigned)ceil_div(rows * K, 256), 256, 0, s>>>(static_cast<const bf16*>(W), w_il,
                                                                                    rows, K, (int)D);
  rw::pack_adaln_kernel<float, float><<<(unsigned)ceil_div(rows, 256), 256, 0, s>>>(bias, b_il, rows, 1, (int)D);
  rw::rowstats_kernel<bf16><<<(unsigned)ceil_div(M, rw::WARPS), rw::THREADS, 0, s>>>(static_cast<const bf16*>(x), rstat,
                                                                                   M, (int)D, 1e-6f);
  tc::AdaLNArgs ada{};
  ada.x = static_cast<const bf16*>(x); ada.ldx = D; ada.rowstats = rstat;
  ada.gate = static_cast<bf16*>(gate_out); ada.ldg = D; ada.features = (int)D;
  int rc = tc::launch_adaln(static_cast<const bf16*>(A), K, w_il, K, b_il, static_cast<bf16*>(h_out), D, ada, (int)M,
                            (int)rows, (int)K, s, cta_group);
  cudaStreamSynchronize(s);
  cudaFree(w_il);
  cudaFree(b_il);
  cudaFree(rstat);
  return rc;
}

// Test hook: SM-clock stamps of the last chain-kernel launch (cluster 0, CTA 0), 8 per stage; see chain_tcgen05.cu.
// Needs NOVA_B200_CHAIN_TIMELINE=1 in the environment of the process.  Synchronises the device.
extern "C" int nova_debug_chain_timeline(int64_t* out, int32_t n) {
  NOVA_REQUIRE(out != nullptr && n > 0, "nova_debug_chain_timeline: bad arguments");
  long long* buf = chain::timeline_buffer(false);
  NOVA_REQUIRE(buf != nullptr, "nova_debug_chain_timeline: no timeline was recorded (set NOVA_B200_CHAIN_TIMELINE=1)");
  const int m = n < chain::TIMELINE_SLOTS ? n : chain::TIMELINE_SLOTS;
  NOVA_CHECK_CUDA(cudaDeviceSynchronize());
  NOVA_CHECK_CUDA(cudaMemcpy(out, buf, m * sizeof(long long), cudaMemcpyDeviceToHost));
  return NOVA_OK;
}
