// Multi-GPU entry points of the C ABI: the ONE collective of the sampling path -- an all-gather of the generated
// points after the last Euler step (clouds are sharded data-parallel, there is no collective inside the denoise loop).
//
// NCCL is not linked: the library is resolved at first use with dlopen("libnccl.so.2"), so libnova_b200.so loads on a
// box without NCCL, and inside a PyTorch process the already-loaded (torch-bundled) NCCL is the one that is used.
// Reference: the reference itself has no multi-GPU sampling code; this is north_star's "single NCCL all-gather over
// NVLink of the generated points" (SURVEY.md 8(e)).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
struct ncclComm;
typedef ncclComm* ncclComm_t;
struct NcclUniqueId {
  char internal[128];
};
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7;  // ncclDataType_t::ncclFloat32
constexpr int kNcclInt8 = 0;     // ncclDataType_t::ncclInt8 (bytes)

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather;
  });
  return api;
}

int need_nccl(const char* who) {
  if (!nccl().ok) {
    nova::set_error("%s: NCCL (libnccl.so.2) could not be loaded: %s", who, dlerror() ? dlerror() : "symbols missing");
    return NOVA_ERR_CUDA;
  }
  return NOVA_OK;
}

int check_nccl(ncclResult_t r, const char* what) {
  if (r == 0) return NOVA_OK;
  nova::set_error("%s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
  return NOVA_ERR_CUDA;
}

}  // namespace

extern "C" int nova_comm_unique_id(char* out128) {
  NOVA_REQUIRE(out128 != nullptr, "nova_comm_unique_id: null argument");
  NOVA_PROPAGATE(need_nccl("nova_comm_unique_id"));
  NcclUniqueId id;
  NOVA_PROPAGATE(check_nccl(nccl().GetUniqueId(&id), "ncclGetUniqueId"));
  std::memcpy(out128, id.internal, sizeof(id.internal));
  return NOVA_OK;
}

extern "C" int nova_comm_init_rank(const char* id128, int32_t world_size, int32_t rank, nova_comm_t** out) {
  NOVA_REQUIRE(id128 && out && world_size >= 1 && rank >= 0 && rank < world_size, "nova_comm_init_rank: bad arguments");
  NOVA_PROPAGATE(need_nccl("nova_comm_init_rank"));
  NcclUniqueId id;
  std::memcpy(id.internal, id128, sizeof(id.internal));
  ncclComm_t comm = nullptr;
  NOVA_PROPAGATE(check_nccl(nccl().CommInitRank(&comm, world_size, id, rank), "ncclCommInitRank"));
  *out = reinterpret_cast<nova_comm_t*>(comm);
  return NOVA_OK;
}

extern "C" int nova_comm_destroy(nova_comm_t* comm) {
  if (!comm) return NOVA_OK;
  NOVA_PROPAGATE(need_nccl("nova_comm_destroy"));
  return check_nccl(nccl().CommDestroy(reinterpret_cast<ncclComm_t>(comm)), "ncclCommDestroy");
}

extern "C" int nova_allgather(nova_comm_t* comm, const void* send, void* recv, int64_t count_bytes, void* stream) {
  NOVA_REQUIRE(comm && recv && count_bytes >= 0 && (send || count_bytes == 0), "nova_allgather: bad arguments");
  if (count_bytes == 0) return NOVA_OK;
  NOVA_PROPAGATE(need_nccl("nova_allgather"));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool words = count_bytes % 4 == 0;  // fp32 outputs: gather as 4-byte elements
  return check_nccl(nccl().AllGather(send, recv, words ? static_cast<size_t>(count_bytes / 4) : static_cast<size_t>(count_bytes),
                                     words ? kNcclFloat32 : kNcclInt8, reinterpret_cast<ncclComm_t>(comm), s),
                    "ncclAllGather");
}
