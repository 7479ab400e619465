// Chain-kernel instantiations, 128 rows per cluster, and the dispatcher (see chain_tcgen05.cuh).
#include <cstdlib>

#include "chain_tcgen05.cuh"

namespace nova {
namespace chain {

int launch_rows128(const ChainParams& p, const bf16* w_stack, cudaStream_t stream) { return launch_rows<128>(p, w_stack, stream); }

bool supported(int D) { return D > 0 && D % 256 == 0 && D <= 2048; }

// Measured (B200, 25-step calls of 32 clouds x n tokens, scripts/gpu_chain_sweep.sh, ms per call, chain vs launches):
//   D = 768  (64-row clusters up to 896 rows, 128-row clusters above)
//     768 rows 2.64 vs 3.46 | 1024: 3.42 vs 3.77 | 1632: 3.47 vs 4.33 | 1792: 3.51 vs 4.41 | 2048: 6.71 vs 4.57 (second wave)
//   D = 1024   768: 3.32 vs 4.36 | 1024: 4.08 vs 4.62 | 1632: 4.56 vs 5.42 | 1792: 5.09 vs 5.46
//   D = 1536   768: 5.34 vs 6.15 | 1024 (128-row clusters): 6.95 vs 6.74
// i.e. the kernel wins while all of its clusters are resident in ONE wave -- with 64 rows per cluster at any width,
// with 128 rows per cluster up to D = 1024 (at D = 1536 a CTA's 590 KB weight slice per stage costs more than the
// launches it saves).  An earlier build put 128-row clusters level with the launch chain; the group-wise ring barriers,
// the whole-warp MMA loop and the packed-fp32 epilogues moved that.
int64_t profitable_rows(int D) {
  if (const char* env = std::getenv("NOVA_B200_CHAIN_ROWS")) return std::atoll(env);
  if (!supported(D)) return 0;
  const int64_t r64 = static_cast<int64_t>(64) * max_clusters64(D);
  const int64_t r128 = D <= 1024 ? static_cast<int64_t>(128) * max_clusters_rows<128>(D) : 0;
  return r64 > r128 ? r64 : r128;
}

long long* timeline_buffer(bool create) {
  static long long* buf = nullptr;
  if (buf == nullptr && create) {
    if (cudaMalloc(&buf, TIMELINE_SLOTS * sizeof(long long)) != cudaSuccess) buf = nullptr;
    else cudaMemset(buf, 0, TIMELINE_SLOTS * sizeof(long long));
  }
  return buf;
}

// 64 rows per cluster while all ceil(M / 64) clusters are resident at once (cudaOccupancyMaxActiveClusters: clusters
// of 8 must sit inside one GPC, so fewer than 148 / 8 fit), 128 rows above.  A second wave is always CORRECT (clusters
// are independent), it just doubles the time.  NOVA_B200_CHAIN_CLUSTER_ROWS=64|128 forces one (tests, A/B runs).
int launch(const ChainParams& p, const bf16* w_stack, cudaStream_t stream) {
  if (p.M <= 0) return NOVA_OK;
  NOVA_REQUIRE(supported(p.D), "chain kernel: unsupported width %d", p.D);
  int rows = ceil_div(p.M, 64) <= max_clusters64(p.D) ? 64 : 128;
  if (const char* env = std::getenv("NOVA_B200_CHAIN_CLUSTER_ROWS")) {
    const int v = std::atoi(env);
    if (v == 64 || v == 128) rows = v;
  }
  return rows == 64 ? launch_rows64(p, w_stack, stream) : launch_rows128(p, w_stack, stream);
}

}  // namespace chain
}  // namespace nova
