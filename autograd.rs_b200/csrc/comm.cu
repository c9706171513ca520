// Data-parallel gradient exchange: NCCL all-reduce over NVLink 5 / NVSwitch on a dedicated communication stream,
// ordered against the context's compute stream with events (no host synchronisation anywhere).
// The reference has no multi-device concept; this is the exchange step the BASELINE north star adds: parameters replicated, minibatch rows sharded, dW/db summed.
//
// libnccl.so.2 is resolved at run time with dlopen, so the library loads (and every single-GPU path works) on a
// machine without NCCL, and inside a torch process it binds to the copy torch already loaded.
#include <dlfcn.h>
#include <string.h>

#include "agrs_internal.h"

namespace {

// the few NCCL declarations used (ABI-stable since NCCL 2.0; nccl.h itself is not needed at build time)
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4 };
typedef int (*PFN_GetUniqueId)(NcclUniqueId*);
typedef int (*PFN_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*PFN_CommDestroy)(NcclComm);
typedef int (*PFN_AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*PFN_GetErrorString)(int);

struct NcclApi {
  void* handle = nullptr;
  PFN_GetUniqueId GetUniqueId = nullptr;
  PFN_CommInitRank CommInitRank = nullptr;
  PFN_CommDestroy CommDestroy = nullptr;
  PFN_AllReduce AllReduce = nullptr;
  PFN_GetErrorString GetErrorString = nullptr;
  std::string why;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown");
    return &api;
  }
  api.GetUniqueId = reinterpret_cast<PFN_GetUniqueId>(dlsym(api.handle, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<PFN_CommInitRank>(dlsym(api.handle, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<PFN_CommDestroy>(dlsym(api.handle, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<PFN_AllReduce>(dlsym(api.handle, "ncclAllReduce"));
  api.GetErrorString = reinterpret_cast<PFN_GetErrorString>(dlsym(api.handle, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
    api.why = "libnccl.so.2 lacks a required symbol";
    api.handle = nullptr;
  }
  return &api;
}

}  // namespace

struct agrs_comm {
  agrs_ctx* ctx = nullptr;
  NcclComm comm = nullptr;
  int world = 1, rank = 0;
  cudaStream_t stream = nullptr;   // communication stream
  cudaEvent_t ready = nullptr;     // compute -> comm: the gradient is complete
  cudaEvent_t done = nullptr;      // comm -> compute: the collectives enqueued so far are complete
};

#define AGRS_NCCL(ctx, expr)                                                                                  \
  do {                                                                                                        \
    int _r = (expr);                                                                                          \
    if (_r != kNcclSuccess) {                                                                                 \
      (ctx)->sticky = AGRS_ERR_CUDA;                                                                          \
      return agrs_fail((ctx), AGRS_ERR_CUDA, "%s failed: %s", #expr,                                          \
                       nccl()->GetErrorString ? nccl()->GetErrorString(_r) : "nccl error");                   \
    }                                                                                                         \
  } while (0)

extern "C" {

int agrs_comm_unique_id(void* out_id128) {
  if (!out_id128) return AGRS_ERR_INVALID;
  NcclApi* api = nccl();
  if (!api->handle) return AGRS_ERR_UNSUPPORTED;
  NcclUniqueId id;
  if (api->GetUniqueId(&id) != kNcclSuccess) return AGRS_ERR_CUDA;
  memcpy(out_id128, &id, sizeof(id));
  return AGRS_OK;
}

int agrs_comm_create(agrs_ctx* ctx, int world, int rank, const void* id128, agrs_comm** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_comm_create");
  AGRS_REQUIRE(ctx, out && id128 && world >= 1 && rank >= 0 && rank < world, AGRS_ERR_INVALID, "agrs_comm_create: bad arguments");
  *out = nullptr;
  NcclApi* api = nccl();
  AGRS_REQUIRE(ctx, api->handle, AGRS_ERR_UNSUPPORTED, "agrs_comm_create: %s", api->why.c_str());
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  agrs_comm* c = new agrs_comm();
  c->ctx = ctx;
  c->world = world;
  c->rank = rank;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  int r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != kNcclSuccess) {
    delete c;
    return agrs_fail(ctx, AGRS_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "nccl error");
  }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) != cudaSuccess) {
    api->CommDestroy(c->comm);
    delete c;
    return agrs_fail(ctx, AGRS_ERR_CUDA, "agrs_comm_create: stream/event creation failed");
  }
  *out = c;
  return AGRS_OK;
}

int agrs_comm_destroy(agrs_comm* c) {
  if (!c) return AGRS_ERR_INVALID;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm) nccl()->CommDestroy(c->comm);
  cudaEventDestroy(c->ready);
  cudaEventDestroy(c->done);
  cudaStreamDestroy(c->stream);
  delete c;
  return AGRS_OK;
}

int agrs_comm_info(agrs_comm* c, int* world, int* rank) {
  if (!c) return AGRS_ERR_INVALID;
  if (world) *world = c->world;
  if (rank) *rank = c->rank;
  return AGRS_OK;
}

int agrs_comm_allreduce_f32(agrs_comm* c, float* buf, size_t n, int average) {
  if (!c) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = c->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_comm_allreduce_f32");
  if (n == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, buf, AGRS_ERR_INVALID, "agrs_comm_allreduce_f32: null pointer");
  // the collective starts once everything enqueued so far on the compute stream (the kernel that produced buf) is done
  AGRS_CUDA(ctx, cudaEventRecord(c->ready, ctx->stream));
  AGRS_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ready, 0));
  AGRS_NCCL(ctx, nccl()->AllReduce(buf, buf, n, kNcclFloat32, average ? kNcclAvg : kNcclSum, c->comm, c->stream));
  return AGRS_OK;
}

int agrs_comm_join(agrs_comm* c) {
  if (!c) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = c->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_comm_join");
  AGRS_CUDA(ctx, cudaEventRecord(c->done, c->stream));
  AGRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->done, 0));
  return AGRS_OK;
}

}  // extern "C"
