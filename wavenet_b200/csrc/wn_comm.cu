// Data-parallel exchange step behind the C ABI: ONE sum all-reduce of the flat gradient buffer per step,
// NCCL over NVLink / NVSwitch.  The reference has no multi-GPU path at all; a maintainer who shards the batch of
// train_audio/train.py:58-80 across processes calls wn_comm_unique_id on rank 0, ships the 128 bytes to the other ranks by
// any means, wn_comm_init everywhere, then wn_allreduce_grads between wn_backward and wn_clip_adam_step(grad_scale = 1/N)
// -- the clip then acts on the AVERAGED gradient exactly like the reference's hook order (wavenet.py:477-480).
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": inside a torch process that is the copy torch already loaded), so
// libwavenet_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <string.h>

#include "wn_common.h"

namespace {

typedef struct { char internal[128]; } NcclUniqueId;     // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

struct NcclApi {
  void* lib = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn get_error_string = nullptr;
  GetVersionFn get_version = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.get_unique_id = (GetUniqueIdFn)dlsym(api.lib, "ncclGetUniqueId");
      api.comm_init_rank = (CommInitRankFn)dlsym(api.lib, "ncclCommInitRank");
      api.all_reduce = (AllReduceFn)dlsym(api.lib, "ncclAllReduce");
      api.comm_destroy = (CommDestroyFn)dlsym(api.lib, "ncclCommDestroy");
      api.get_error_string = (GetErrorStringFn)dlsym(api.lib, "ncclGetErrorString");
      api.get_version = (GetVersionFn)dlsym(api.lib, "ncclGetVersion");
      if (!api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.comm_destroy) api.lib = nullptr;
    }
  }
  return api.lib ? &api : nullptr;
}

int nccl_fail(const char* what, int rc) {
  NcclApi* a = nccl();
  wn_set_error("%s -> NCCL error %d (%s)", what, rc, a && a->get_error_string ? a->get_error_string(rc) : "?");
  return WN_ECUDA;
}

}  // namespace

extern "C" int wn_comm_available(void) { return nccl() ? 1 : 0; }

extern "C" int wn_comm_unique_id(char* id_out_host) {
  WN_REQUIRE(id_out_host, WN_EINVAL, "null argument");
  NcclApi* a = nccl();
  WN_REQUIRE(a, WN_ESTATE, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "not found");
  NcclUniqueId id;
  const int rc = a->get_unique_id(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id_out_host, id.internal, sizeof(id.internal));
  return WN_OK;
}

extern "C" int wn_comm_init(wn_handle* h, const char* id_host, int rank, int world) {
  WN_REQUIRE(h && id_host, WN_EINVAL, "null argument");
  WN_REQUIRE(world >= 1 && rank >= 0 && rank < world, WN_EINVAL, "bad rank %d / world %d", rank, world);
  NcclApi* a = nccl();
  WN_REQUIRE(a, WN_ESTATE, "libnccl.so.2 could not be loaded");
  if (h->comm) {
    a->comm_destroy((NcclComm)h->comm);
    h->comm = nullptr;
  }
  NcclUniqueId id;
  memcpy(id.internal, id_host, sizeof(id.internal));
  NcclComm comm = nullptr;
  const int rc = a->comm_init_rank(&comm, world, id, rank);   // binds to the CURRENT CUDA device
  if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
  h->comm = comm;
  h->comm_rank = rank;
  h->comm_world = world;
  return WN_OK;
}

extern "C" int wn_comm_world(const wn_handle* h) { return h && h->comm ? h->comm_world : 1; }

extern "C" int wn_allreduce_grads(wn_handle* h, float* grads, wn_stream_t st) {
  WN_REQUIRE(h && grads, WN_EINVAL, "null argument");
  if (!h->comm || h->comm_world == 1) return WN_OK;
  NcclApi* a = nccl();
  WN_REQUIRE(a, WN_ESTATE, "libnccl.so.2 could not be loaded");
  const int rc = a->all_reduce(grads, grads, (size_t)h->flat_size, /*ncclFloat32*/ 7, /*ncclSum*/ 0, (NcclComm)h->comm,
                               (cudaStream_t)st);
  if (rc != 0) return nccl_fail("ncclAllReduce", rc);
  wn_count_launch();
  return WN_OK;
}

extern "C" int wn_comm_destroy(wn_handle* h) {
  WN_REQUIRE(h, WN_EINVAL, "null handle");
  if (h->comm) {
    NcclApi* a = nccl();
    if (a) a->comm_destroy((NcclComm)h->comm);
    h->comm = nullptr;
  }
  h->comm_world = 1;
  h->comm_rank = 0;
  return WN_OK;
}
