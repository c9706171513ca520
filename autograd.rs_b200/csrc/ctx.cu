// Context, stream-ordered memory pool and transfers of libagrs_b200.so.
// Replaces CudaData{ptr} + Drop of the reference (src/tensor/storage.rs:5-14).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "agrs_internal.h"

int agrs_fail(agrs_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    ctx->err = buf;
  }
  return code;
}

static int ctx_init(int device, cudaStream_t stream, bool own, agrs_ctx** out) {
  if (!out) return AGRS_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return AGRS_ERR_NODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AGRS_ERR_NODEVICE;
  if (prop.major != 10) return AGRS_ERR_NODEVICE;  // sm_100a code only: no fallback of any kind
  agrs_ctx* ctx = new agrs_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  if (const char* e = getenv("AGRS_CONV_TILED")) ctx->conv_tiled = atoi(e) != 0;
  if (const char* e = getenv("AGRS_TC_WAVE_SYNC")) {
    const int v = atoi(e);
    if (v >= 0 && v <= 2) ctx->tc_wave_sync = v;
  }
  if (const char* e = getenv("AGRS_TC_CROSS")) {  // default of AGRS_TUNE_TC_CROSS (agrs.h)
    const int v = atoi(e);
    if (v >= 0 && v <= 5) ctx->tc_cross = v;
  }
  if (const char* e = getenv("AGRS_TC_L2_HINT")) ctx->tc_l2_hint = atoi(e);
  if (const char* e = getenv("AGRS_TC_ACT_WARPS")) ctx->tc_act_warps = atoi(e) ? 1 : 0;  // default of AGRS_TUNE_TC_ACT_WARPS
  if (const char* e = getenv("AGRS_TC_EPI_TMA")) ctx->tc_epi_tma = atoi(e) ? 1 : 0;  // default of AGRS_TUNE_TC_EPI_TMA
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return AGRS_ERR_CUDA; }
  if (const char* e = getenv("AGRS_L2_PERSIST_MB")) {
    // evict_last lines only persist inside the L2's persisting carve-out, which is empty by default (experiment of AGRS_TC_L2_HINT)
    size_t want = size_t(atoi(e)) << 20;
    if (want > size_t(prop.persistingL2CacheMaxSize)) want = size_t(prop.persistingL2CacheMaxSize);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) (void)cudaGetLastError();
    fprintf(stderr, "agrs: persisting L2 carve-out %zu MB of at most %d MB (L2 %d MB)\n", want >> 20, prop.persistingL2CacheMaxSize >> 20, prop.l2CacheSize >> 20);
  }
  if (own) {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return AGRS_ERR_CUDA; }
    ctx->owns_stream = true;
  } else {
    ctx->stream = stream;
  }
  // the device's default pool, told to keep freed blocks (no trim at sync points): steady-state
  // training steps then allocate without touching the driver.
  if (cudaDeviceGetDefaultMemPool(&ctx->pool, device) != cudaSuccess) { delete ctx; return AGRS_ERR_CUDA; }
  uint64_t keep = UINT64_MAX;
  cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
  ctx->scratch_bytes = 64 * 1024 * sizeof(float);
  if (cudaMalloc(&ctx->scratch, ctx->scratch_bytes) != cudaSuccess || cudaMalloc(&ctx->counter, 64) != cudaSuccess ||
      cudaMemsetAsync(ctx->counter, 0, 64, ctx->stream) != cudaSuccess) {
    delete ctx;
    return AGRS_ERR_CUDA;
  }
  *out = ctx;
  return AGRS_OK;
}

extern "C" {

const char* agrs_version(void) { return "agrs_b200 0.1 (sm_100a)"; }

int agrs_ctx_create(int device, agrs_ctx** out) { return ctx_init(device, nullptr, true, out); }

int agrs_ctx_create_on_stream(int device, void* cuda_stream, agrs_ctx** out) {
  return ctx_init(device, static_cast<cudaStream_t>(cuda_stream), false, out);
}

int agrs_ctx_destroy(agrs_ctx* ctx) {
  if (!ctx) return AGRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->counter) cudaFree(ctx->counter);
  if (ctx->workspace) cudaFree(ctx->workspace);
  if (ctx->tickets) cudaFree(ctx->tickets);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->copy_ready);
    cudaEventDestroy(ctx->copy_done);
  }
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AGRS_OK;
}

const char* agrs_last_error(agrs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int agrs_sync(agrs_ctx* ctx) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sync");
  AGRS_NOT_CAPTURING(ctx, "agrs_sync");
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AGRS_OK;
}

struct agrs_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t kernel_nodes = 0, total_nodes = 0;
  uint64_t launches_per_replay = 0;  // kernels this library launched inside the capture
};

int agrs_capture_begin(agrs_ctx* ctx) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_capture_begin");
  AGRS_NOT_CAPTURING(ctx, "agrs_capture_begin");
  AGRS_REQUIRE(ctx, !ctx->profiling, AGRS_ERR_UNSUPPORTED, "agrs_capture_begin: disable agrs_profile_enable first");
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  AGRS_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true;
  ctx->capture_allocs.clear();
  ctx->deferred_frees.clear();
  ctx->launches_at_capture = ctx->launches;
  return AGRS_OK;
}

int agrs_capture_end(agrs_ctx* ctx, agrs_graph** out) {
  if (!ctx) return AGRS_ERR_INVALID;
  AGRS_REQUIRE(ctx, ctx->capturing, AGRS_ERR_INVALID, "agrs_capture_end without agrs_capture_begin");
  AGRS_REQUIRE(ctx, out, AGRS_ERR_INVALID, "agrs_capture_end: null out");
  *out = nullptr;
  ctx->capturing = false;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
  // buffers that predate the capture and were dropped inside it
  for (void* p : ctx->deferred_frees) cudaFreeAsync(p, ctx->stream);
  ctx->deferred_frees.clear();
  ctx->capture_allocs.clear();
  if (e != cudaSuccess || !g) {
    cudaGetLastError();
    if (ctx->sticky) return ctx->sticky;  // the capture was already broken by a failing call: keep its message
    return agrs_fail(ctx, AGRS_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
  }
  agrs_graph* gr = new agrs_graph();
  gr->graph = g;
  gr->launches_per_replay = ctx->launches - ctx->launches_at_capture;
  size_t n = 0;
  if (cudaGraphGetNodes(g, nullptr, &n) == cudaSuccess && n) {
    std::vector<cudaGraphNode_t> nodes(n);
    if (cudaGraphGetNodes(g, nodes.data(), &n) == cudaSuccess) {
      gr->total_nodes = n;
      for (auto nd : nodes) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nd, &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) gr->kernel_nodes++;
      }
    }
  }
  // buffers allocated in the graph and still owned by a tensor when the capture ended are released by the graph itself at
  // the next replay (and re-allocated at the same address)
  e = cudaGraphInstantiateWithFlags(&gr->exec, g, cudaGraphInstantiateFlagAutoFreeOnLaunch);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g);
    delete gr;
    cudaGetLastError();
    return agrs_fail(ctx, AGRS_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  }
  *out = gr;
  return AGRS_OK;
}

int agrs_graph_launch(agrs_ctx* ctx, agrs_graph* graph, int times) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_graph_launch");
  AGRS_NOT_CAPTURING(ctx, "agrs_graph_launch");
  AGRS_REQUIRE(ctx, graph && graph->exec && times >= 0, AGRS_ERR_INVALID, "agrs_graph_launch: bad arguments");
  for (int i = 0; i < times; ++i) AGRS_CUDA(ctx, cudaGraphLaunch(graph->exec, ctx->stream));
  ctx->launches += graph->launches_per_replay * uint64_t(times);
  return AGRS_OK;
}

int agrs_graph_destroy(agrs_ctx* ctx, agrs_graph* graph) {
  if (!graph) return AGRS_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  if (graph->exec) cudaGraphExecDestroy(graph->exec);
  if (graph->graph) cudaGraphDestroy(graph->graph);
  delete graph;
  return AGRS_OK;
}

int agrs_graph_info(agrs_graph* graph, uint64_t* kernel_nodes, uint64_t* total_nodes) {
  if (!graph) return AGRS_ERR_INVALID;
  if (kernel_nodes) *kernel_nodes = graph->kernel_nodes;
  if (total_nodes) *total_nodes = graph->total_nodes;
  return AGRS_OK;
}

int agrs_timer_begin(agrs_ctx* ctx) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_timer_begin");
  AGRS_NOT_CAPTURING(ctx, "agrs_timer_begin");
  if (!ctx->timer_a) {
    AGRS_CUDA(ctx, cudaEventCreate(&ctx->timer_a));
    AGRS_CUDA(ctx, cudaEventCreate(&ctx->timer_b));
  }
  AGRS_CUDA(ctx, cudaEventRecord(ctx->timer_a, ctx->stream));
  return AGRS_OK;
}

int agrs_timer_end(agrs_ctx* ctx, float* ms) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_timer_end");
  AGRS_NOT_CAPTURING(ctx, "agrs_timer_end");
  AGRS_REQUIRE(ctx, ms && ctx->timer_a, AGRS_ERR_INVALID, "agrs_timer_end without agrs_timer_begin");
  AGRS_CUDA(ctx, cudaEventRecord(ctx->timer_b, ctx->stream));
  AGRS_CUDA(ctx, cudaEventSynchronize(ctx->timer_b));
  AGRS_CUDA(ctx, cudaEventElapsedTime(ms, ctx->timer_a, ctx->timer_b));
  return AGRS_OK;
}

void* agrs_stream(agrs_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

uint64_t agrs_launch_count(agrs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int agrs_device_info(agrs_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_device_info");
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (free_bytes || total_bytes) {
    size_t f = 0, t = 0;
    AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
    AGRS_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
  }
  return AGRS_OK;
}

int agrs_alloc(agrs_ctx* ctx, size_t nbytes, void** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_alloc");
  AGRS_REQUIRE(ctx, out != nullptr, AGRS_ERR_INVALID, "agrs_alloc: null out");
  if (nbytes == 0) nbytes = 4;  // empty tensors still get a distinct address
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  AGRS_CUDA(ctx, cudaMallocFromPoolAsync(out, nbytes, ctx->pool, ctx->stream));
  if (ctx->capturing) ctx->capture_allocs.insert(*out);
  return AGRS_OK;
}

int agrs_pool_reserve(agrs_ctx* ctx, size_t nbytes) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_pool_reserve");
  AGRS_NOT_CAPTURING(ctx, "agrs_pool_reserve");
  if (nbytes == 0) return AGRS_OK;
  void* p = nullptr;
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  AGRS_CUDA(ctx, cudaMallocFromPoolAsync(&p, nbytes, ctx->pool, ctx->stream));
  AGRS_CUDA(ctx, cudaFreeAsync(p, ctx->stream));
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AGRS_OK;
}

int agrs_pool_trim(agrs_ctx* ctx) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_pool_trim");
  AGRS_NOT_CAPTURING(ctx, "agrs_pool_trim");
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  AGRS_CUDA(ctx, cudaMemPoolTrimTo(ctx->pool, 0));
  uint64_t zero = 0;
  AGRS_CUDA(ctx, cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReservedMemHigh, &zero));
  AGRS_CUDA(ctx, cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrUsedMemHigh, &zero));
  return AGRS_OK;
}

int agrs_pool_stats(agrs_ctx* ctx, size_t* reserved_bytes, size_t* used_bytes, size_t* reserved_high, size_t* used_high) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_pool_stats");
  uint64_t v = 0;
  if (reserved_bytes) { AGRS_CUDA(ctx, cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &v)); *reserved_bytes = size_t(v); }
  if (used_bytes) { AGRS_CUDA(ctx, cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &v)); *used_bytes = size_t(v); }
  if (reserved_high) { AGRS_CUDA(ctx, cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemHigh, &v)); *reserved_high = size_t(v); }
  if (used_high) { AGRS_CUDA(ctx, cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemHigh, &v)); *used_high = size_t(v); }
  return AGRS_OK;
}

int agrs_free(agrs_ctx* ctx, void* ptr) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_free");
  if (!ptr) return AGRS_OK;
  if (ctx->capturing) {
    auto it = ctx->capture_allocs.find(ptr);
    if (it == ctx->capture_allocs.end()) {  // allocated before the capture: a graph cannot free it; release it when the capture ends
      ctx->deferred_frees.push_back(ptr);
      return AGRS_OK;
    }
    ctx->capture_allocs.erase(it);
  }
  AGRS_CUDA(ctx, cudaFreeAsync(ptr, ctx->stream));
  return AGRS_OK;
}

int agrs_host_alloc(agrs_ctx* ctx, size_t nbytes, void** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_host_alloc");
  AGRS_REQUIRE(ctx, out != nullptr, AGRS_ERR_INVALID, "agrs_host_alloc: null out");
  AGRS_CUDA(ctx, cudaMallocHost(out, nbytes ? nbytes : 4));
  return AGRS_OK;
}

int agrs_host_free(agrs_ctx* ctx, void* ptr) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_host_free");
  if (ptr) AGRS_CUDA(ctx, cudaFreeHost(ptr));
  return AGRS_OK;
}

int agrs_upload(agrs_ctx* ctx, void* dst, const void* src, size_t nbytes) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_upload");
  if (nbytes == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, dst && src, AGRS_ERR_INVALID, "agrs_upload: null pointer");
  // pageable sources are staged by the runtime before the call returns; pinned ones stay async
  AGRS_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  return AGRS_OK;
}

int agrs_download(agrs_ctx* ctx, void* dst, const void* src, size_t nbytes) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_download");
  if (nbytes == 0) return AGRS_OK;
  AGRS_NOT_CAPTURING(ctx, "agrs_download");
  AGRS_REQUIRE(ctx, dst && src, AGRS_ERR_INVALID, "agrs_download: null pointer");
  AGRS_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AGRS_OK;
}

int agrs_upload_prefetch(agrs_ctx* ctx, void* dst, const void* src, size_t nbytes) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_upload_prefetch");
  if (nbytes == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, dst && src, AGRS_ERR_INVALID, "agrs_upload_prefetch: null pointer");
  if (!ctx->copy_stream) {
    AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
    AGRS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    AGRS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_ready, cudaEventDisableTiming));
    AGRS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
  }
  AGRS_CUDA(ctx, cudaEventRecord(ctx->copy_ready, ctx->stream));
  AGRS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ready, 0));
  AGRS_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  AGRS_CUDA(ctx, cudaEventRecord(ctx->copy_done, ctx->copy_stream));
  return AGRS_OK;
}

int agrs_prefetch_join(agrs_ctx* ctx) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_prefetch_join");
  if (!ctx->copy_stream) return AGRS_OK;
  AGRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
  return AGRS_OK;
}

struct agrs_event {
  cudaEvent_t ev = nullptr;
  bool recorded = false;
};

int agrs_event_create(agrs_ctx* ctx, agrs_event** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, out, AGRS_ERR_INVALID, "agrs_event_create: null out");
  *out = nullptr;
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  agrs_event* e = new agrs_event();
  cudaError_t err = cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming);
  if (err != cudaSuccess) {
    delete e;
    return agrs_fail(ctx, AGRS_ERR_CUDA, "agrs_event_create: %s", cudaGetErrorString(err));
  }
  *out = e;
  return AGRS_OK;
}

int agrs_event_destroy(agrs_ctx* ctx, agrs_event* ev) {
  AGRS_CHECK_CTX(ctx);
  if (!ev) return AGRS_OK;
  if (ev->ev) AGRS_CUDA(ctx, cudaEventDestroy(ev->ev));
  delete ev;
  return AGRS_OK;
}

int agrs_download_async(agrs_ctx* ctx, void* dst, const void* src, size_t nbytes, agrs_event* done) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_download_async");
  AGRS_NOT_CAPTURING(ctx, "agrs_download_async");
  AGRS_REQUIRE(ctx, done && done->ev && (nbytes == 0 || (dst && src)), AGRS_ERR_INVALID, "agrs_download_async: null pointer");
  if (nbytes) AGRS_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  AGRS_CUDA(ctx, cudaEventRecord(done->ev, ctx->stream));
  done->recorded = true;
  return AGRS_OK;
}

int agrs_upload_prefetch_event(agrs_ctx* ctx, void* dst, const void* src, size_t nbytes, agrs_event* landed) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, landed && landed->ev, AGRS_ERR_INVALID, "agrs_upload_prefetch_event: null event");
  if (int rc = agrs_upload_prefetch(ctx, dst, src, nbytes)) return rc;
  if (nbytes == 0) return AGRS_OK;
  AGRS_CUDA(ctx, cudaEventRecord(landed->ev, ctx->copy_stream));
  landed->recorded = true;
  return AGRS_OK;
}

int agrs_wait_event(agrs_ctx* ctx, agrs_event* ev) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, ev && ev->ev, AGRS_ERR_INVALID, "agrs_wait_event: null event");
  AGRS_NOT_CAPTURING(ctx, "agrs_wait_event (touch a prefetched tensor, or call agrs_prefetch_join, before the capture begins)");
  if (ev->recorded) AGRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev->ev, 0));
  return AGRS_OK;
}

int agrs_event_wait(agrs_ctx* ctx, agrs_event* ev) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_event_wait");
  AGRS_NOT_CAPTURING(ctx, "agrs_event_wait");
  AGRS_REQUIRE(ctx, ev && ev->ev, AGRS_ERR_INVALID, "agrs_event_wait: null event");
  if (!ev->recorded) return AGRS_OK;  // nothing was ever enqueued behind it
  AGRS_CUDA(ctx, cudaEventSynchronize(ev->ev));
  return AGRS_OK;
}

int agrs_copy(agrs_ctx* ctx, void* dst, const void* src, size_t nbytes) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_copy");
  if (nbytes == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, dst && src, AGRS_ERR_INVALID, "agrs_copy: null pointer");
  AGRS_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return AGRS_OK;
}

int agrs_set_gemm_path(agrs_ctx* ctx, int path) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_set_gemm_path");
  AGRS_REQUIRE(ctx, path >= AGRS_GEMM_AUTO && path <= AGRS_GEMM_SKINNY, AGRS_ERR_INVALID, "unknown gemm path %d", path);
  ctx->gemm_path = path;
  return AGRS_OK;
}

int agrs_last_gemm_path(agrs_ctx* ctx) { return ctx ? ctx->last_gemm_path : 0; }

}  // extern "C"

int agrs_ensure_workspace(agrs_ctx* ctx, size_t nbytes) {
  if (ctx->workspace_bytes >= nbytes) return AGRS_OK;
  AGRS_NOT_CAPTURING(ctx, "growing the split-K workspace (run one ordinary step before capturing)");
  // growing is rare (first split-K call of a given size); stream-sync so nothing in flight uses the old block
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->workspace) AGRS_CUDA(ctx, cudaFree(ctx->workspace));
  ctx->workspace = nullptr;
  ctx->workspace_bytes = 0;
  size_t want = nbytes < (size_t(16) << 20) ? (size_t(16) << 20) : nbytes;
  AGRS_CUDA(ctx, cudaMalloc(&ctx->workspace, want));
  ctx->workspace_bytes = want;
  return AGRS_OK;
}
