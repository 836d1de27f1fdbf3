// dist.cu — multi-GPU plumbing: one ctx = one rank = one GPU, NCCL over
// NVLink 5 / NVSwitch.  Used in exactly the places the path has a real exchange
// step: the all-reduce of the partial reduced camera system and
// of the LM scalars in the point-sharded large BA.  The keyframe-pair sweep and
// the batched windows shard by work item and need no collective.
//
// libnccl is resolved at run time (dlopen) so that a process which already
// carries an NCCL (torch's bundled libnccl.so.2) shares that one copy, and a
// plain C++ host picks up the system library.
#include <dlfcn.h>
#include <nccl.h>

#include <new>

#include "dist.cuh"

namespace lorb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
        dlclose(api.handle);
        api.handle = nullptr;
      }
    }
  }
  return api.handle ? &api : nullptr;
}

struct Dist {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

#define LORB_NCCL_TRY(api, expr)                                                            \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != ncclSuccess) {                                                                \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                               \
                (api)->GetErrorString ? (api)->GetErrorString(_r) : "nccl error");          \
      return LORB_ERR_NCCL;                                                                 \
    }                                                                                       \
  } while (0)

bool dist_ready(lorb_ctx* c) { return c && c->dist && c->dist->comm; }

void dist_destroy(lorb_ctx* c) {
  if (!c || !c->dist) return;
  NcclApi* api = nccl_api();
  if (api && c->dist->comm) api->CommDestroy(c->dist->comm);
  delete c->dist;
  c->dist = nullptr;
}

int dist_allreduce_sum(lorb_ctx* c, double* dev, size_t n) {
  NcclApi* api = nccl_api();
  if (!api || !dist_ready(c)) {
    set_error("collective requested without an initialised communicator");
    return LORB_ERR_STATE;
  }
  LORB_NCCL_TRY(api, api->AllReduce(dev, dev, n, ncclDouble, ncclSum, c->dist->comm, c->stream));
  return LORB_OK;
}

int dist_allreduce_max_u64(lorb_ctx* c, double* dev, size_t n) {
  NcclApi* api = nccl_api();
  if (!api || !dist_ready(c)) {
    set_error("collective requested without an initialised communicator");
    return LORB_ERR_STATE;
  }
  LORB_NCCL_TRY(api, api->AllReduce(dev, dev, n, ncclUint64, ncclMax, c->dist->comm, c->stream));
  return LORB_OK;
}

int dist_allreduce_max_i32(lorb_ctx* c, int* dev, size_t n) {
  NcclApi* api = nccl_api();
  if (!api || !dist_ready(c)) {
    set_error("collective requested without an initialised communicator");
    return LORB_ERR_STATE;
  }
  LORB_NCCL_TRY(api, api->AllReduce(dev, dev, n, ncclInt32, ncclMax, c->dist->comm, c->stream));
  return LORB_OK;
}

int dist_rank(lorb_ctx* c) { return dist_ready(c) ? c->dist->rank : 0; }
int dist_world(lorb_ctx* c) { return dist_ready(c) ? c->dist->world : 1; }

}  // namespace lorb

using namespace lorb;

extern "C" {

int lorb_dist_get_unique_id(uint8_t id[LORB_NCCL_UNIQUE_ID_BYTES]) {
  LORB_REQUIRE(id, "id");
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded: %s", dlerror());
    return LORB_ERR_NCCL;
  }
  static_assert(sizeof(ncclUniqueId) == LORB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  LORB_NCCL_TRY(api, api->GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return LORB_OK;
}

int lorb_dist_init(lorb_ctx* c, const uint8_t id[LORB_NCCL_UNIQUE_ID_BYTES], int rank, int world) {
  LORB_REQUIRE(c && id, "ctx / id");
  LORB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank / world");
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded: %s", dlerror());
    return LORB_ERR_NCCL;
  }
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  dist_destroy(c);
  c->dist = new (std::nothrow) Dist();
  if (!c->dist) return LORB_ERR_NOMEM;
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  c->dist->rank = rank;
  c->dist->world = world;
  LORB_NCCL_TRY(api, api->CommInitRank(&c->dist->comm, world, u, rank));
  return LORB_OK;
}

int lorb_dist_finalize(lorb_ctx* c) {
  LORB_REQUIRE(c, "ctx");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  dist_destroy(c);
  return LORB_OK;
}

int lorb_dist_allreduce_f64(lorb_ctx* c, double* data, int n) {
  LORB_REQUIRE(c && (n == 0 || data) && n >= 0, "ctx / data");
  LORB_REQUIRE(dist_ready(c), "lorb_dist_init was not called");
  LORB_CUDA_TRY(cudaSetDevice(c->device));
  if (n == 0) return LORB_OK;
  LORB_TRY(dev_reserve(c, 15, (size_t)n * 8));
  double* d = c->d[15].as<double>();
  LORB_CUDA_TRY(cudaMemcpyAsync(d, data, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  LORB_TRY(dist_allreduce_sum(c, d, (size_t)n));
  LORB_CUDA_TRY(cudaMemcpyAsync(data, d, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  LORB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return LORB_OK;
}

}  // extern "C"
