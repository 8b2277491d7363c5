// comm.cu - the one collective of the path: a sum all-reduce of the root action statistics over
// the GPUs of a host (NCCL over NVLink 5 / NVSwitch). NCCL is dlopen()ed so that the library has
// no link-time dependency on it: inside a torch process the already-loaded torch-bundled
// libnccl.so.2 is picked up; stand-alone C++ hosts get the system one.
#include <dlfcn.h>

#include <cstring>

#include "engine.h"

#if __has_include(<nccl.h>)
#include <nccl.h>
#else
// Minimal declarations matching NCCL 2.x's public ABI (nccl.h).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
#endif

namespace mamcts {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl(mamcts_engine* e) {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.handle) return MAMCTS_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(e, MAMCTS_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
  NcclApi api;
  api.handle = h;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString)
    return fail(e, MAMCTS_ERR_NCCL, "libnccl is missing a required symbol");
  g_nccl = api;
  return MAMCTS_OK;
}

static int nccl_fail(mamcts_engine* e, ncclResult_t r, const char* what) {
  return fail(e, MAMCTS_ERR_NCCL, std::string("NCCL error: ") + g_nccl.GetErrorString(r) + " in " + what);
}

}  // namespace mamcts

using namespace mamcts;

extern "C" {

int mamcts_comm_get_unique_id(mamcts_engine* e, uint8_t id[MAMCTS_NCCL_UNIQUE_ID_BYTES]) {
  if (!e || !id) return fail(e, MAMCTS_ERR_INVALID, "mamcts_comm_get_unique_id: null argument");
  int rc = load_nccl(e);
  if (rc != MAMCTS_OK) return rc;
  static_assert(sizeof(ncclUniqueId) == MAMCTS_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId uid;
  ncclResult_t r = g_nccl.GetUniqueId(&uid);
  if (r != ncclSuccess) return nccl_fail(e, r, "ncclGetUniqueId");
  std::memcpy(id, &uid, sizeof(uid));
  return MAMCTS_OK;
}

int mamcts_comm_init_rank(mamcts_engine* e, const uint8_t id[MAMCTS_NCCL_UNIQUE_ID_BYTES],
                          int32_t world_size, int32_t rank) {
  if (!e || !id) return fail(e, MAMCTS_ERR_INVALID, "mamcts_comm_init_rank: null argument");
  if (world_size < 1 || rank < 0 || rank >= world_size)
    return fail(e, MAMCTS_ERR_INVALID, "mamcts_comm_init_rank: bad rank / world size");
  int rc = load_nccl(e);
  if (rc != MAMCTS_OK) return rc;
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  if (e->nccl_comm) return fail(e, MAMCTS_ERR_INVALID, "mamcts_comm_init_rank: communicator already attached");
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&comm, world_size, uid, rank);
  if (r != ncclSuccess) return nccl_fail(e, r, "ncclCommInitRank");
  e->nccl_comm = comm;
  e->world_size = world_size;
  e->rank = rank;
  return MAMCTS_OK;
}

int mamcts_comm_destroy(mamcts_engine* e) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_comm_destroy: null engine");
  if (!e->nccl_comm) return MAMCTS_OK;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  g_nccl.CommDestroy(static_cast<ncclComm_t>(e->nccl_comm));
  e->nccl_comm = nullptr;
  e->world_size = 1;
  e->rank = 0;
  return MAMCTS_OK;
}

// NOTE: called with e->mu held by root_merge_device, or directly by a host; takes no lock itself.
int mamcts_allreduce_f64(mamcts_engine* e, double* dptr, size_t count) {
  if (!e || !dptr) return fail(e, MAMCTS_ERR_INVALID, "mamcts_allreduce_f64: null argument");
  if (!e->nccl_comm) return MAMCTS_OK;  // single rank: the sum is the local value
  ncclResult_t r = g_nccl.AllReduce(dptr, dptr, count, ncclFloat64, ncclSum,
                                    static_cast<ncclComm_t>(e->nccl_comm), e->stream);
  if (r != ncclSuccess) return nccl_fail(e, r, "ncclAllReduce");
  return MAMCTS_OK;
}

}  // extern "C"
