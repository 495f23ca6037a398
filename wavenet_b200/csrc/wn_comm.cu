// Data-parallel exchange step behind the C ABI: ONE sum all-reduce of the flat gradient buffer per step,
// NCCL over NVLink / NVSwitch.  The reference has no multi-GPU path at all; a maintainer who shards the batch of
// train_audio/train.py:58-80 across processes calls wn_comm_unique_id on rank 0, ships the 128 bytes to the other ranks by
// any means, wn_comm_init everywhere, then wn_allreduce_grads between wn_backward and wn_clip_adam_step(grad_scale = 1/N)
// -- the clip then acts on the AVERAGED gradient exactly like the reference's hook order (wavenet.py:477-480).
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": inside a torch process that is the copy torch already loaded), so
// libwavenet_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "wn_common.h"

namespace {

typedef struct { char internal[128]; } NcclUniqueId;     // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

struct NcclApi {
  void* lib = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  AllGatherFn all_gather = nullptr;
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
      api.all_gather = (AllGatherFn)dlsym(api.lib, "ncclAllGather");
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

// ---- one-shot peer-memory all-reduce fused with the optimiser's first pass ------------------------------------------
// Every rank owns an exchange buffer (cudaMalloc + CUDA IPC, opened by all peers over NVLink).  ONE kernel per step:
//   0. wait until every peer has finished reading our buffer from the previous step, copy the local gradient into it, and --
//      once the whole grid has done so -- post "ready = step" into every peer's flag block (remote stores);
//   1. wait for every peer's ready flag;
//   2. read ALL buffers (own + peers, P2P loads) and add them IN RANK ORDER -- every rank computes the same bits --, apply the
//      1/N mean and the weight-decay hook, write the gradient back and accumulate its squared norm (the optimiser's first
//      pass, wavenet.py:175-182, which then does not run as a kernel of its own);
//   3. once the whole grid is through, post "done = step" to the peers.
// Flags only ever grow (the step number), waits are ">= step": no reset races.  Waits are bounded: a peer that never arrives
// sets the error word and the kernel returns instead of hanging the GPU.
constexpr int PEER_MAX = 8;
struct PeerDev {
  float* xbuf[PEER_MAX];
  int* flags[PEER_MAX];        // per rank: ready[PEER_MAX] at +0, done[PEER_MAX] at +16 ints, grid counters at +32 / +33, error at +34
  int world, rank;
};
struct PeerState {
  PeerDev dev;
  void* local = nullptr;       // the cudaMalloc'ed block of this rank (exchange buffer + flag block)
  void* opened[PEER_MAX] = {};
  uint32_t step = 0;
};

__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool wait_flags(const int* flags, int world, int rank, int want) {
  for (int src = 0; src < world; ++src) {
    if (src == rank) continue;
    uint32_t spins = 0;
    while (ld_flag(flags + src) < want) {
      if (++spins > (1u << 24)) return false;
      __nanosleep(64);
    }
  }
  return true;
}

__global__ void __launch_bounds__(1024) peer_allreduce_norm_kernel(PeerDev pd, float* __restrict__ grads, const float* __restrict__ params,
                                                                   int64_t n, int step, float grad_scale, float wd,
                                                                   double* __restrict__ acc, double* __restrict__ det_partials) {
  int* my = pd.flags[pd.rank];
  __shared__ int s_ok;
  __shared__ double part[32];
  // ---- 0. our buffer is free again; publish the local gradient ----
  if (threadIdx.x == 0) s_ok = wait_flags(my + 16, pd.world, pd.rank, step - 1) ? 1 : 0;
  __syncthreads();
  float* mine = pd.xbuf[pd.rank];
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4)
    *reinterpret_cast<float4*>(mine + i) = *reinterpret_cast<const float4*>(grads + i);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(my + 32), 1u) + 1u;
    if (done == (unsigned)step * gridDim.x) {              // last CTA of this step's grid
      __threadfence_system();
      for (int p = 0; p < pd.world; ++p)
        if (p != pd.rank) st_flag(pd.flags[p] + pd.rank, step);
    }
    // ---- 1. every peer has published ----
    if (!wait_flags(my, pd.world, pd.rank, step)) s_ok = 0;
  }
  __syncthreads();
  if (!s_ok) {
    if (threadIdx.x == 0) my[34] = 1;
    return;
  }
  // ---- 2. sum in rank order, mean + weight decay, squared norm ----
  double sq = 0.0;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < pd.world; ++r) {
      const float4 v = ld_peer4(pd.xbuf[r] + i);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    a.x *= grad_scale, a.y *= grad_scale, a.z *= grad_scale, a.w *= grad_scale;
    if (wd > 0.f) {
      const float4 pp = *reinterpret_cast<const float4*>(params + i);
      a.x += wd * pp.x, a.y += wd * pp.y, a.z += wd * pp.z, a.w += wd * pp.w;
    }
    *reinterpret_cast<float4*>(grads + i) = a;
    sq += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    if (det_partials)
      det_partials[blockIdx.x] = t;
    else
      atomicAdd(acc, t);
    // ---- 3. the peers may overwrite their buffers once our whole grid has read them ----
    __threadfence_system();
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(my + 33), 1u) + 1u;
    if (done == (unsigned)step * gridDim.x)
      for (int p = 0; p < pd.world; ++p)
        if (p != pd.rank) st_flag(pd.flags[p] + 16 + pd.rank, step);
  }
}

void peer_free(wn_handle* h) {
  PeerState* ps = (PeerState*)h->peer;
  if (!ps) return;
  for (int p = 0; p < PEER_MAX; ++p)
    if (ps->opened[p]) cudaIpcCloseMemHandle(ps->opened[p]);
  if (ps->local) cudaFree(ps->local);
  delete ps;
  h->peer = nullptr;
}

// Collective (every rank calls it after ncclCommInitRank): all ranks end up with the peer path enabled, or none does.
void peer_setup(wn_handle* h, NcclApi* a) {
  const int world = h->comm_world, rank = h->comm_rank;
  const char* env = getenv("WN_FUSED_ALLREDUCE");
  int want = (env && atoi(env) != 0 && world >= 2 && world <= PEER_MAX && a->all_gather && h->flat_size % 4 == 0) ? 1 : 0;
  cudaStream_t s = 0;
  int* d_ok = nullptr;
  if (cudaMalloc(&d_ok, sizeof(int)) != cudaSuccess) return;
  auto agree = [&](int v) {      // min over ranks
    cudaMemcpy(d_ok, &v, sizeof(int), cudaMemcpyHostToDevice);
    if (a->all_reduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, (NcclComm)h->comm, s) != 0) return 0;
    cudaStreamSynchronize(s);
    cudaMemcpy(&v, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
  };
  if (!agree(want)) {
    cudaFree(d_ok);
    return;
  }
  PeerState* ps = new PeerState();
  const size_t xbytes = ((size_t)h->flat_size * sizeof(float) + 255) / 256 * 256, total = xbytes + 256;
  int ok = cudaMalloc(&ps->local, total) == cudaSuccess && cudaMemset(ps->local, 0, total) == cudaSuccess ? 1 : 0;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, ps->local) != cudaSuccess) ok = 0;
  cudaIpcMemHandle_t* d_h = nullptr;
  if (cudaMalloc(&d_h, sizeof(cudaIpcMemHandle_t) * (world + 1)) != cudaSuccess) ok = 0;
  std::vector<cudaIpcMemHandle_t> all(world);
  if (d_h) {
    cudaMemcpy(d_h + world, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    if (a->all_gather(d_h + world, d_h, sizeof(mine), /*ncclInt8*/ 0, (NcclComm)h->comm, s) != 0) ok = 0;
    cudaStreamSynchronize(s);
    cudaMemcpy(all.data(), d_h, sizeof(mine) * world, cudaMemcpyDeviceToHost);
    cudaFree(d_h);
  }
  ok = agree(ok);                 // everybody allocated and exported
  if (ok) {
    for (int p = 0; p < world && ok; ++p) {
      void* base = ps->local;
      if (p != rank) {
        if (cudaIpcOpenMemHandle(&base, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
          break;
        }
        ps->opened[p] = base;
      }
      ps->dev.xbuf[p] = (float*)base;
      ps->dev.flags[p] = (int*)((char*)base + xbytes);
    }
  }
  ok = agree(ok);                 // everybody opened every peer
  cudaFree(d_ok);
  ps->dev.world = world;
  ps->dev.rank = rank;
  h->peer = ps;
  if (!ok) peer_free(h);
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
  peer_free(h);
  peer_setup(h, a);              // opt-in (WN_FUSED_ALLREDUCE=1): CUDA-IPC exchange buffers for the one-shot fused all-reduce
  return WN_OK;
}

extern "C" int wn_comm_peer_enabled(const wn_handle* h) { return h && h->peer ? 1 : 0; }

// all-reduce + [1/N, weight decay, squared norm] + clip + Adam.  Peer path: ONE fused kernel + the clip/Adam kernel; otherwise
// ncclAllReduce followed by wn_clip_adam_step's two kernels.  grad_scale multiplies the SUMMED gradient (1 / (world * micro_batches)).
extern "C" int wn_allreduce_clip_adam_step(wn_handle* h, float* params, float* grads, float* m, float* v, int t, float lr, float beta1,
                                           float beta2, float eps, float weight_decay, float clip, float grad_scale, void* scratch,
                                           float* norm_out, wn_stream_t st) {
  WN_REQUIRE(h && params && grads && m && v && scratch, WN_EINVAL, "null argument");
  cudaStream_t s = (cudaStream_t)st;
  PeerState* ps = (PeerState*)h->peer;
  if (!ps) {
    WN_TRY(wn_allreduce_grads(h, grads, st));
    return wn_clip_adam_step(h, params, grads, m, v, t, lr, beta1, beta2, eps, weight_decay, clip, grad_scale, scratch, norm_out, st);
  }
  WN_REQUIRE(t >= 1, WN_EINVAL, "Adam step count must be >= 1");
  double* det_partials = h->deterministic ? reinterpret_cast<double*>(h->det_slab + (int64_t)h->det_nslab * h->flat_size) : nullptr;
  WN_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), s));
  ps->step += 1;
  peer_allreduce_norm_kernel<<<h->sm_count, 1024, 0, s>>>(ps->dev, grads, params, h->flat_size, (int)ps->step, grad_scale,
                                                            weight_decay, (double*)scratch, det_partials);
  WN_CHECK_LAUNCH();
  return optim_adam_after_norm(params, grads, m, v, h->flat_size, t, lr, beta1, beta2, eps, clip, (double*)scratch, norm_out,
                               h->sm_count, s, det_partials, h->sm_count);
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
  peer_free(h);
  if (h->comm) {
    NcclApi* a = nccl();
    if (a) a->comm_destroy((NcclComm)h->comm);
    h->comm = nullptr;
  }
  h->comm_world = 1;
  h->comm_rank = 0;
  return WN_OK;
}
