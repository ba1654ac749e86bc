// The one collective of the path: an all-gather of the per-rank output latents, issued IN-STREAM
// through the C ABI right behind the sampling loop. The path has no cross-sample reduction, so there is nothing to fuse it
// with; NCCL over NVLink 5 / NVSwitch does the transfer. libnccl is not linked: the entry points are resolved at run time
// from the library the host process already uses (torch's bundled libnccl.so.2 by default), so the product library keeps
// having no dependency beyond the CUDA driver.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "engine.h"

namespace dv {
namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
typedef int ncclResult_t;
constexpr int kNcclFloat = 7;          // ncclFloat32 in nccl.h (ncclDataType_t)

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi& nccl_api(const char* path) {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (api.handle) return api;
  const char* cands[] = {path, getenv("DV_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* c : cands) {
    if (!c || !*c) continue;
    api.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
    api.err = dlerror();
  }
  if (!api.handle) return api;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) {
    api.err = "libnccl does not export the expected entry points";
    dlclose(api.handle);
    api.handle = nullptr;
  }
  return api;
}

std::string nccl_err(NcclApi& a, ncclResult_t r, const char* what) {
  return std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error") + " (" + std::to_string(r) + ")";
}

}  // namespace

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = -1;
};

int comm_unique_id(const char* nccl_path, void* id128, std::string& err) {
  NcclApi& a = nccl_api(nccl_path);
  if (!a.handle) { err = "cannot load libnccl: " + a.err; return DV_ERR_UNSUPPORTED; }
  NcclUniqueId id;
  ncclResult_t r = a.GetUniqueId(&id);
  if (r != 0) { err = nccl_err(a, r, "ncclGetUniqueId"); return DV_ERR_CUDA; }
  memcpy(id128, &id, sizeof id);
  return DV_OK;
}

int comm_init(const char* nccl_path, const void* id128, int rank, int nranks, Comm** out, std::string& err) {
  NcclApi& a = nccl_api(nccl_path);
  if (!a.handle) { err = "cannot load libnccl: " + a.err; return DV_ERR_UNSUPPORTED; }
  if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) { err = "bad communicator arguments"; return DV_ERR_INVALID; }
  NcclUniqueId id;
  memcpy(&id, id128, sizeof id);
  Comm* c = new Comm();
  c->rank = rank; c->nranks = nranks;
  if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; err = "cudaGetDevice failed"; return DV_ERR_CUDA; }
  ncclResult_t r = a.CommInitRank(&c->comm, nranks, id, rank);     // collective: every rank calls it with the same id
  if (r != 0) { delete c; err = nccl_err(a, r, "ncclCommInitRank"); return DV_ERR_CUDA; }
  *out = c;
  return DV_OK;
}

int comm_all_gather(Comm* c, const float* send, float* recv, size_t count, cudaStream_t st, std::string& err) {
  if (!c || !c->comm || !send || !recv) { err = "null communicator / buffer"; return DV_ERR_INVALID; }
  NcclApi& a = nccl_api(nullptr);
  ncclResult_t r = a.AllGather(send, recv, count, kNcclFloat, c->comm, st);
  if (r != 0) { err = nccl_err(a, r, "ncclAllGather"); return DV_ERR_CUDA; }
  return DV_OK;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  NcclApi& a = nccl_api(nullptr);
  if (c->comm && a.handle) a.CommDestroy(c->comm);
  delete c;
}

}  // namespace dv
