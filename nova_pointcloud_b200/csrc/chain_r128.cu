// Chain-kernel instantiations, 128 rows per cluster, and the dispatcher (see chain_tcgen05.cuh).
#include <cstdlib>

#include "chain_tcgen05.cuh"

namespace nova {
namespace chain {

int launch_rows128(const ChainParams& p, const bf16* w_stack, cudaStream_t stream) { return launch_rows<128>(p, w_stack, stream); }

bool supported(int D) { return D > 0 && D % 256 == 0 && D <= 2048; }

// Measured (B200, D = 768, 25-step calls, scripts/profile_sets.py): the chain kernel wins while every cluster owns 64
// rows and all clusters are resident in one wave (2.42 vs 2.49 ms at 32 rows, 2.58 vs 3.11 ms at 512, 3.01 vs 3.40 ms
// at 768); with 128 rows per cluster it is level with the launch chain (3.84 vs 3.68 ms at 1024, 4.16 vs 4.25 at 1632).
int64_t profitable_rows(int D) {
  if (const char* env = std::getenv("NOVA_B200_CHAIN_ROWS")) return std::atoll(env);
  return supported(D) ? static_cast<int64_t>(64) * max_clusters64(D) : 0;
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
