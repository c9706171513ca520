// Fused data-parallel optimiser step over NVLink peer memory: reduce-scatter + SGD + all-gather without NCCL, replacing
//     all-reduce(grad)  ->  w -= lr * grad  (optim.rs:25-34)
// Every rank keeps its parameters and a gradient exchange region in an "arena" (one cudaMalloc block, IPC-exported, mapped by
// every peer at the same offsets).  The flat parameter space is dealt out in granules of G = 2^gshift floats: granule g belongs
// to rank g % world (interleaved: whatever part of it a kernel is working on, its traffic spreads over all ranks' NVLink ports);
// rank r OWNS the slice made of its granules and its gradient region holds one SLOT per source rank: [world][slice].
//     scatter     during backward every rank pushes its gradient values into slot [its rank] of the owner's region with P2P
//                 stores over NVLink -- straight from the dW GEMM's epilogue, tile by tile (gemm_tcgen05_2cta.cu), or with
//                 k_scatter_grad for gradients produced elsewhere (bias sums, conv filters, split-K results)
// Two ways to finish the step:
//  (A) per BUCKET, overlapped with the rest of backward (agrs_fused_begin_step / _bucket_ready / _finish_step; the default).
//      A bucket is a contiguous range of the parameter space (one layer's W and b) whose gradients are complete at the same
//      point of the backward walk.  As soon as they have been pushed:
//        compute stream   k_signal: "rank r's pushes for bucket b of this step have landed" -> push flag PF(b, r) at every rank
//        comm stream      wait PF(b, q) for every q (stream memory operations: no SM is held while waiting)
//                         k_bucket_sgd: LOCAL sum of the slots in rank order, w -= lr * mean(g) on the owned granules
//        copy streams     all-gather with the COPY ENGINES: one strided (2-D) peer copy of the owned granules to every peer,
//                         then weight flag WF(b, r) at that peer
//      and the backward of the earlier layers keeps the tensor cores busy meanwhile.
//      By default the scatter itself also runs on the copy engines in this mode (one strided peer copy per gradient and owner, the
//      push flag written on the same copy stream right behind it): the dW GEMMs then keep their TMA epilogue and their K split,
//      and no SM does communication work at all.  AGRS_FUSED_SCATTER=kernel keeps the pushes in the GEMM epilogue / k_scatter_grad.
//      agrs_fused_finish_step makes the compute
//      stream wait for its own comm stream and for WF(b, q) of every bucket and peer.  Only the first layer's exchange (the last
//      one to become ready) is exposed.
//  (B) one kernel after backward (agrs_fused_sgd_step): cross-rank barrier, slot sum, SGD, all-gather with P2P stores, barrier.
// Hazards of (A), all closed by the flags: a peer's push of step s+1 into my slots comes after its finish_step(s), which waited
// for my WF(b, me), written after my k_bucket_sgd consumed the slots; a peer's weight copy into my W comes after its wait on
// PF(b, me), signalled after my dX and dW GEMMs of that layer (the last readers of W this step) in stream order.
// Replicas stay bit-identical (every element has one owner whose result is copied) and sums run in rank order: deterministic.
#include <string.h>

#include "agrs_internal.h"

namespace {

constexpr int kMaxWorld = 16;
constexpr int kMaxBuckets = 64;
constexpr int kFlagFloats = 4096;  // first 16 KB of every arena: flag words (unsigned), then the data regions
// flag words: [0, 68) barrier state of the one-kernel step (B); PF(b, q) and WF(b, q) of the bucket protocol (A)
__host__ __device__ constexpr size_t PF(int b, int q) { return 128 + size_t(b) * kMaxWorld + q; }
__host__ __device__ constexpr size_t WF(int b, int q) { return 128 + size_t(kMaxBuckets) * kMaxWorld + size_t(b) * kMaxWorld + q; }
static_assert(WF(kMaxBuckets - 1, kMaxWorld - 1) < kFlagFloats, "flag words");
// Bound of every barrier spin: each poll is a system- or device-scope load (~0.5-1 us), so a rank that never shows up traps
// the kernel after roughly half a minute instead of hanging the GPU; honest skew between ranks is milliseconds.
constexpr unsigned kMaxSpins = 1u << 25;

struct FusedArgs {
  float* base[kMaxWorld];  // arena of every rank, as mapped in THIS process (base[rank] is the local one)
  int world, rank;
  size_t w_off, g_off;     // float offsets of the weight and gradient regions inside an arena
  size_t w_out;            // where the bucket protocol WRITES the new weights: the other weight region when double-buffered (else w_off)
  size_t total;            // floats in each region; a multiple of (granule * world)
  float lr;
  unsigned epoch;          // strictly increasing per step: flag / barrier generation
  int gshift;              // log2 of the ownership granule in floats (>= 5: a 32-float piece never straddles two owners)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// data accesses to peer arenas: system-scope, so neither a stale L1 line nor a cached remote line can be read
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- ownership map: flat element e (a multiple of 4) -> owner rank, index inside the owner's slice ----
__device__ __forceinline__ void owner_of(const FusedArgs& a, size_t e, size_t& owner, size_t& idx) {
  const size_t g = e >> a.gshift;
  owner = g % size_t(a.world);
  idx = ((g / size_t(a.world)) << a.gshift) + (e & ((size_t(1) << a.gshift) - 1));
}

// This rank's OWN gradients of a bucket, where they were produced (the parameters' gradient buffers): with the scatter on the copy
// engines nothing copies them into slot [rank] -- a local copy crawls while a GEMM owns the SMs (profiles/r2_fused_exchange.md) --
// the slot sum reads them in place.  Elements outside every segment (padding between parameters) count as zero.
constexpr int kMaxSegs = 8;
struct OwnSegs {
  int n;                     // 0: the own contribution is in slot [rank] like everybody else's
  const float* src[kMaxSegs];
  size_t off[kMaxSegs], len[kMaxSegs];  // flat range [off, off + len) of the parameter space, off a multiple of 4
};
__device__ __forceinline__ float4 own_at(const OwnSegs& sg, size_t e) {
#pragma unroll 1
  for (int i = 0; i < sg.n; ++i) {
    if (e >= sg.off[i] && e < sg.off[i] + sg.len[i]) {
      const float* p = sg.src[i] + (e - sg.off[i]);
      if (e + 4 <= sg.off[i] + sg.len[i]) return __ldcs(reinterpret_cast<const float4*>(p));
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);  // ragged end of a parameter
      const size_t left = sg.off[i] + sg.len[i] - e;
      r.x = p[0];
      if (left > 1) r.y = p[1];
      if (left > 2) r.z = p[2];
      return r;
    }
  }
  return make_float4(0.f, 0.f, 0.f, 0.f);
}

// slot sum in rank order + SGD on element e4 (float4) of the owned slice index idx4; returns the new weights
template <int WORLD>
__device__ __forceinline__ float4 reduce_and_update(const FusedArgs& a, int world, size_t slice, size_t idx, size_t e, float inv, const OwnSegs* own = nullptr) {
  const float* slots = a.base[a.rank] + a.g_off + idx;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (WORLD) {
    float4 v[WORLD ? WORLD : 1];
#pragma unroll
    for (int p = 0; p < (WORLD ? WORLD : 1); ++p)  // all loads in flight before the first add
      v[p] = (own && own->n && p == a.rank) ? own_at(*own, e) : ld_sys_v4(slots + size_t(p) * slice);
#pragma unroll
    for (int p = 0; p < (WORLD ? WORLD : 1); ++p) { g.x += v[p].x; g.y += v[p].y; g.z += v[p].z; g.w += v[p].w; }
  } else {
    for (int p = 0; p < world; ++p) {
      const float4 v = (own && own->n && p == a.rank) ? own_at(*own, e) : ld_sys_v4(slots + size_t(p) * slice);
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
  }
  g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
  float4 w = ld_sys_v4(a.base[a.rank] + a.w_off + e);
  w.x = __fsub_rn(w.x, __fmul_rn(g.x, a.lr));  // same two roundings as agrs_sgd_step_f32
  w.y = __fsub_rn(w.y, __fmul_rn(g.y, a.lr));
  w.z = __fsub_rn(w.z, __fmul_rn(g.z, a.lr));
  w.w = __fsub_rn(w.w, __fmul_rn(g.w, a.lr));
  return w;
}

// ============================================================================ (B) the one-kernel step
// Cross-rank + grid barrier `which` (0 or 1).  Flags of an arena (unsigned words):
//   [which*32 + p]  "rank p reached barrier `which` of generation epoch"      written by rank p into EVERY arena
//   [64 + which]    "all ranks reached it": local go flag for the other CTAs  written by this rank's CTA 0
//   [80 + 2*which]  CTA arrival counter of this rank (64-bit, device scope; never reset: generation g ends at g * gridDim.x)
__device__ void rank_barrier(const FusedArgs& a, int which) {
  unsigned* mine = reinterpret_cast<unsigned*>(a.base[a.rank]);
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(mine + 80) + which;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's loads/stores of the phase before the barrier are performed
    atomicAdd(counter, 1ull);
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {  // wait for every CTA of this rank, then tell the peers
      const unsigned long long want = (unsigned long long)gridDim.x * a.epoch;
      for (unsigned spins = 0; ld_acquire_gpu64(counter) < want; ++spins)
        if (spins > kMaxSpins) { printf("agrs fused step: local barrier timeout (rank %d)\n", a.rank); __trap(); }
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < a.world) {
      const int p = threadIdx.x;
      st_release_sys(reinterpret_cast<unsigned*>(a.base[p]) + which * 32 + a.rank, a.epoch);
      for (unsigned spins = 0; ld_acquire_sys(mine + which * 32 + p) < a.epoch; ++spins)
        if (spins > kMaxSpins) { printf("agrs fused step: rank %d never saw rank %d at barrier %d\n", a.rank, p, which); __trap(); }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      st_release_sys(mine + 64 + which, a.epoch);
    }
  } else if (threadIdx.x == 0) {
    for (unsigned spins = 0; ld_acquire_gpu(mine + 64 + which) < a.epoch; ++spins)
      if (spins > kMaxSpins) { printf("agrs fused step: go-flag timeout (rank %d)\n", a.rank); __trap(); }
    __threadfence_system();
  }
  __syncthreads();
}

template <int WORLD>  // 0 = run-time world size
__global__ void __launch_bounds__(256) k_fused_dp_sgd(const FusedArgs a) {
  const int world = WORLD ? WORLD : a.world;
  rank_barrier(a, 0);
  const size_t slice = a.total / size_t(world);
  const size_t slice4 = slice / 4;  // float4 per rank
  const float inv = 1.0f / float(world);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < slice4; i += stride) {
    // my slice is every world-th granule: slice index 4i lies in my granule (4i >> gshift), which is granule
    // (4i >> gshift) * world + rank of the flat parameter space
    const size_t idx = 4 * i;
    const size_t e = ((((idx >> a.gshift) * size_t(world)) + size_t(a.rank)) << a.gshift) + (idx & ((size_t(1) << a.gshift) - 1));
    const float4 w = reduce_and_update<WORLD>(a, world, slice, idx, e, inv);
#pragma unroll
    for (int j = 0; j < (WORLD ? WORLD : kMaxWorld); ++j) {
      if (j >= world) break;
      int p = a.rank + 1 + j;
      p -= p >= world ? world : 0;
      st_sys_v4(a.base[p] + a.w_off + e, w);
    }
  }
  rank_barrier(a, 1);
}

// ============================================================================ (A) the bucket protocol
// "this rank's pushes for bucket b of step `step` have landed": one release store per rank, after a system-scope fence
__global__ void k_signal(const FusedArgs a, int bucket, unsigned step, int weights) {
  if (int(threadIdx.x) < a.world) {
    __threadfence_system();
    const size_t word = weights ? WF(bucket, a.rank) : PF(bucket, a.rank);
    st_release_sys(reinterpret_cast<unsigned*>(a.base[threadIdx.x]) + word, step);
  }
}
// one flag word at one peer (after a peer copy on a copy stream, when stream memory operations are not available)
__global__ void k_set_flag(unsigned* word, unsigned step) {
  __threadfence_system();
  st_release_sys(word, step);
}
// spin (bounded) until every listed flag word of MY arena has reached `step`; threads stride over the list
struct WaitList {
  int n;
  unsigned short word[kMaxBuckets * kMaxWorld];
};
__global__ void __launch_bounds__(256) k_wait_flags(const unsigned* flags, const __grid_constant__ WaitList wl, unsigned step, int rank) {
  for (int i = threadIdx.x; i < wl.n; i += blockDim.x) {
    for (unsigned spins = 0; ld_acquire_sys(flags + wl.word[i]) < step; ++spins)
      if (spins > kMaxSpins) { printf("agrs fused exchange: rank %d never saw flag word %d reach %u\n", rank, int(wl.word[i]), step); __trap(); }
  }
  __threadfence_system();
}

// LOCAL sum of the slots (rank order) + SGD on the owned granules of the bucket [off, off + n); with PUSH the new weights also go
// to every peer with P2P stores (the all-gather without the copy engines: AGRS_FUSED_AG=kernel)
template <int WORLD, bool PUSH>
__global__ void __launch_bounds__(256) k_bucket_sgd(const FusedArgs a, size_t off, size_t n, const __grid_constant__ OwnSegs own) {
  const int world = WORLD ? WORLD : a.world;
  const size_t slice = a.total / size_t(world);
  const float inv = 1.0f / float(world);
  const size_t g0 = off >> a.gshift;                            // first granule of the bucket: a multiple of world
  const size_t rows = (n >> a.gshift) / size_t(world);          // owned granules in the bucket
  const size_t per4 = (size_t(1) << a.gshift) / 4;              // float4 per granule
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * per4; i += stride) {
    const size_t r = i / per4, j = i - r * per4;
    const size_t e = ((g0 + r * size_t(world) + size_t(a.rank)) << a.gshift) + 4 * j;
    const size_t idx = ((g0 / size_t(world) + r) << a.gshift) + 4 * j;
    const float4 w = reduce_and_update<WORLD>(a, world, slice, idx, e, inv, &own);
    st_sys_v4(a.base[a.rank] + a.w_out + e, w);
    if (PUSH) {
#pragma unroll
      for (int jj = 1; jj < (WORLD ? WORLD : kMaxWorld); ++jj) {
        if (jj >= world) break;
        int p = a.rank + jj;
        p -= p >= world ? world : 0;
        st_sys_v4(a.base[p] + a.w_out + e, w);
      }
    }
  }
}

// The scatter half of the reduce-scatter for a gradient that was NOT produced by the scattering GEMM (bias sums, conv filters,
// split-K results, second accumulations): element e of the flat parameter space belongs to the rank that owns its granule, which
// keeps one slot per source rank; this rank's values go to slot [rank] of the owner with P2P stores.  `off` and `n`: multiples of 4.
__global__ void __launch_bounds__(256) k_scatter_grad(const FusedArgs a, const float* __restrict__ src, size_t off, size_t n) {
  const size_t slice = a.total / size_t(a.world);
  const size_t stride = size_t(gridDim.x) * blockDim.x, n4 = (n + 3) / 4;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const size_t e = off + 4 * i;  // a multiple of 4: a float4 never straddles two granules
    size_t owner, idx;
    owner_of(a, e, owner, idx);
    float* dst = a.base[owner] + a.g_off + size_t(a.rank) * slice + idx;
    if (4 * i + 3 < n) {
      st_sys_v4(dst, __ldcs(reinterpret_cast<const float4*>(src) + i));
    } else {
      for (size_t j = 4 * i; j < n; ++j) asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + (j - 4 * i)), "f"(src[j]) : "memory");
    }
  }
}

typedef CUresult (*PFN_waitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*PFN_writeValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

}  // namespace

struct agrs_fused {
  agrs_ctx* ctx = nullptr;
  FusedArgs args{};
  size_t grid = 0;  // (B): CTAs per launch, fixed at the first call (the barrier counters assume it)
  // (A)
  cudaStream_t comm = nullptr;                 // slot sum + SGD kernels, highest priority
  std::vector<cudaStream_t> copy;              // peer copies (copy engines)
  std::vector<cudaEvent_t> copy_done;          // per copy stream: its copies of the previous step have left my weight region
  cudaEvent_t sgd_done = nullptr, comm_done = nullptr, pushed = nullptr;
  PFN_waitValue32 wait32 = nullptr;            // stream memory operations (no SM held while waiting); null: spin kernels
  PFN_writeValue32 write32 = nullptr;
  // Double-buffered weights (w_alt != 0): a step reads region `cur` and the bucket protocol writes the new weights into the other
  // one, at home and at the peers; finish_step swaps.  No kernel of the running step reads what the exchange writes, so a layer's
  // exchange may start before that layer's dX GEMM (which reads W) has run -- Linear::backward then computes dW first and the
  // first layer's exchange hides under its dX GEMM instead of trailing the step.
  // AGRS_FUSED_TRACE=k: time stamps (CUDA events on the streams involved) of step k, printed to stderr by finish_step
  unsigned trace_step = 0;
  struct Mark { const char* what; int bucket; int peer; cudaEvent_t ev; };
  std::vector<Mark> marks;
  cudaEvent_t trace_base = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> exposed;  // while agrs_profile_enable is on: event pairs around finish_step's waits
  size_t w_regions[2] = {0, 0};
  int cur = 0;
  bool ag_kernel = false;                      // all-gather with P2P stores from k_bucket_sgd instead of the copy engines
  bool scatter_dma = true;                     // (A) only: gradients reach the owners' slots by strided peer copies (copy engines) instead
                                               // of P2P stores from the dW GEMM's epilogue / k_scatter_grad: no SM, no slower epilogue
  struct Seg { const float* src; size_t off, n; cudaEvent_t ready; };
  std::vector<Seg> segs;                       // gradients handed to the DMA scatter in the open step (own part is read in place)
  std::vector<cudaEvent_t> ev_pool;            // "this gradient is complete" / chunk fork-join events for the copy streams
  size_t ev_next = 0;
  std::vector<cudaStream_t> helper;            // a large strided copy is cut into row chunks that ride on these streams: one copy
  size_t helper_next = 0;                      // engine per stream, so several engines drive the same NVLink peer at once
  size_t chunk_bytes = size_t(4) << 20;        // copies above 2 x this are cut (AGRS_FUSED_COPY_CHUNK_MB; 0 = never)
  bool scatter_pending = false;                // scatter copies of the open step are in flight on the copy streams
  unsigned step = 0;                           // generation of the step being built (epoch + 1 between begin and finish)
  bool open = false;
  std::vector<int> buckets;                    // buckets exchanged in the open step
  bool copies_pending = false;
};

// what a scattering GEMM epilogue needs to know (gemm_tcgen05_2cta.cu)
void agrs_fused_scatter_target(const agrs_fused* f, float* const** bases, int* world, int* rank, size_t* g_off, size_t* shard, int* gshift) {
  *bases = f->args.base;
  *world = f->args.world;
  *rank = f->args.rank;
  *g_off = f->args.g_off;
  *shard = f->args.total / size_t(f->args.world);
  *gshift = f->args.gshift;
}

extern "C" {

int agrs_peer_alloc(agrs_ctx* ctx, size_t nbytes, void** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_peer_alloc");
  AGRS_NOT_CAPTURING(ctx, "agrs_peer_alloc");
  AGRS_REQUIRE(ctx, out && nbytes > 0, AGRS_ERR_INVALID, "agrs_peer_alloc: bad arguments");
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  AGRS_CUDA(ctx, cudaMalloc(out, nbytes));  // plain allocation: stream-ordered pool memory cannot be exported with cudaIpcGetMemHandle
  AGRS_CUDA(ctx, cudaMemsetAsync(*out, 0, nbytes, ctx->stream));
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AGRS_OK;
}

int agrs_peer_free(agrs_ctx* ctx, void* ptr) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_peer_free");
  if (!ptr) return AGRS_OK;
  AGRS_CUDA(ctx, cudaDeviceSynchronize());  // comm / copy streams of an exchange that used the arena included
  AGRS_CUDA(ctx, cudaFree(ptr));
  return AGRS_OK;
}

int agrs_peer_export(agrs_ctx* ctx, void* ptr, void* handle64) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_peer_export");
  AGRS_REQUIRE(ctx, ptr && handle64, AGRS_ERR_INVALID, "agrs_peer_export: null pointer");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == AGRS_PEER_HANDLE_BYTES, "IPC handle size");
  AGRS_CUDA(ctx, cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return AGRS_OK;
}

int agrs_peer_open(agrs_ctx* ctx, const void* handle64, void** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_peer_open");
  AGRS_REQUIRE(ctx, handle64 && out, AGRS_ERR_INVALID, "agrs_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  AGRS_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return AGRS_OK;
}

int agrs_peer_close(agrs_ctx* ctx, void* ptr) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_peer_close");
  if (ptr) AGRS_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
  return AGRS_OK;
}

size_t agrs_fused_flag_floats(void) { return kFlagFloats; }
int agrs_fused_max_buckets(void) { return kMaxBuckets; }

int agrs_fused_create(agrs_ctx* ctx, int world, int rank, void* const* arenas, size_t w_off, size_t w_alt, size_t g_off, size_t total,
                      int granule_log2, agrs_fused** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fused_create");
  AGRS_REQUIRE(ctx, out && arenas && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, AGRS_ERR_INVALID,
               "agrs_fused_create: bad arguments (world <= %d)", kMaxWorld);
  AGRS_REQUIRE(ctx, granule_log2 >= 5 && granule_log2 <= 24, AGRS_ERR_INVALID, "agrs_fused_create: granule_log2 must be 5 .. 24");
  const size_t G = size_t(1) << granule_log2;
  AGRS_REQUIRE(ctx, total % (G * size_t(world)) == 0 && w_off % 32 == 0 && g_off % 32 == 0 && w_off >= kFlagFloats && g_off >= kFlagFloats,
               AGRS_ERR_INVALID, "agrs_fused_create: regions must start after the %d flag words, on 32-float boundaries, and hold a multiple of granule*world floats", kFlagFloats);
  agrs_fused* f = new agrs_fused();
  f->ctx = ctx;
  for (int p = 0; p < world; ++p) {
    if (!arenas[p] || !agrs_aligned16(arenas[p])) {
      delete f;
      return agrs_fail(ctx, AGRS_ERR_INVALID, "agrs_fused_create: arena %d is null or unaligned", p);
    }
    f->args.base[p] = static_cast<float*>(arenas[p]);
  }
  f->args.world = world;
  f->args.rank = rank;
  AGRS_REQUIRE(ctx, w_alt == 0 || (w_alt % 32 == 0 && w_alt >= kFlagFloats && w_alt != w_off), AGRS_ERR_INVALID, "agrs_fused_create: bad alternate weight region");
  f->args.w_off = w_off;
  f->args.w_out = w_off;
  f->w_regions[0] = w_off;
  f->w_regions[1] = w_alt;
  f->args.g_off = g_off;
  f->args.total = total;
  f->args.epoch = 0;
  f->args.gshift = granule_log2;
  // streams of the bucket protocol
  int lo = 0, hi = 0;
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  AGRS_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
  AGRS_CUDA(ctx, cudaStreamCreateWithPriority(&f->comm, cudaStreamNonBlocking, hi));
  int ncopy = world - 1 < 4 ? world - 1 : 4;
  if (const char* e = getenv("AGRS_FUSED_COPY_STREAMS")) { const int v = atoi(e); if (v >= 1 && v <= 16) ncopy = v < world - 1 ? v : world - 1; }
  for (int i = 0; i < ncopy; ++i) {
    cudaStream_t s;
    cudaEvent_t ev;
    AGRS_CUDA(ctx, cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
    AGRS_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    f->copy.push_back(s);
    f->copy_done.push_back(ev);
  }
  AGRS_CUDA(ctx, cudaEventCreateWithFlags(&f->sgd_done, cudaEventDisableTiming));
  AGRS_CUDA(ctx, cudaEventCreateWithFlags(&f->comm_done, cudaEventDisableTiming));
  AGRS_CUDA(ctx, cudaEventCreateWithFlags(&f->pushed, cudaEventDisableTiming));
  if (const char* e = getenv("AGRS_FUSED_AG")) f->ag_kernel = strcmp(e, "kernel") == 0;
  if (const char* e = getenv("AGRS_FUSED_SCATTER")) f->scatter_dma = strcmp(e, "kernel") != 0;
  if (const char* e = getenv("AGRS_FUSED_TRACE")) f->trace_step = unsigned(atoi(e));
  if (f->copy.empty()) {  // world == 1: the one local copy stream of the DMA scatter
    cudaStream_t s;
    cudaEvent_t ev;
    AGRS_CUDA(ctx, cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
    AGRS_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    f->copy.push_back(s);
    f->copy_done.push_back(ev);
  }
  for (int i = 0; i < 256; ++i) {
    cudaEvent_t ev;
    AGRS_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    f->ev_pool.push_back(ev);
  }
  int nhelp = 6;
  if (const char* e = getenv("AGRS_FUSED_COPY_HELPERS")) { const int v = atoi(e); if (v >= 0 && v <= 16) nhelp = v; }
  if (const char* e = getenv("AGRS_FUSED_COPY_CHUNK_MB")) { const int v = atoi(e); if (v >= 0 && v <= 1024) f->chunk_bytes = size_t(v) << 20; }
  for (int i = 0; i < nhelp; ++i) {
    cudaStream_t s;
    AGRS_CUDA(ctx, cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
    f->helper.push_back(s);
  }
  // stream memory operations, resolved at run time like the tensor-map encoder (no link-time libcuda dependency); a first wait on an
  // already-satisfied local flag word checks that the driver takes them on this memory
  const char* mo = getenv("AGRS_FUSED_MEMOPS");
  if (!mo || atoi(mo) != 0) {
    void *fw = nullptr, *fs = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fw, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
        cudaGetDriverEntryPoint("cuStreamWriteValue32", &fs, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess && fw && fs) {
      f->wait32 = reinterpret_cast<PFN_waitValue32>(fw);
      f->write32 = reinterpret_cast<PFN_writeValue32>(fs);
      unsigned* probe = reinterpret_cast<unsigned*>(f->args.base[rank]) + 100;  // a spare flag word
      bool ok = f->write32(f->comm, CUdeviceptr(uintptr_t(probe)), 1u, 0) == CUDA_SUCCESS &&
                f->wait32(f->comm, CUdeviceptr(uintptr_t(probe)), 1u, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS;
      for (int p = 0; ok && p < world; ++p)  // and on peer-mapped memory: a spare word per source rank
        if (p != rank) ok = f->write32(f->comm, CUdeviceptr(uintptr_t(reinterpret_cast<unsigned*>(f->args.base[p]) + 101 + rank)), 1u, 0) == CUDA_SUCCESS;
      ok = ok && cudaStreamSynchronize(f->comm) == cudaSuccess;
      if (!ok) {
        cudaGetLastError();
        f->wait32 = nullptr;
        f->write32 = nullptr;
      }
    }
  }
  *out = f;
  return AGRS_OK;
}

int agrs_fused_info(agrs_fused* f, int* memops, int* copy_streams, int* allgather_kernel) {
  if (!f) return AGRS_ERR_INVALID;
  if (memops) *memops = f->wait32 ? 1 : 0;
  if (copy_streams) *copy_streams = int(f->copy.size());
  if (allgather_kernel) *allgather_kernel = f->ag_kernel ? 1 : 0;
  return AGRS_OK;
}

int agrs_fused_destroy(agrs_fused* f) {
  if (!f) return AGRS_OK;
  cudaSetDevice(f->ctx->device);
  if (f->comm) { cudaStreamSynchronize(f->comm); cudaStreamDestroy(f->comm); }
  for (auto s : f->copy) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  for (auto e : f->copy_done) cudaEventDestroy(e);
  for (auto e : f->ev_pool) cudaEventDestroy(e);
  for (auto& pr : f->exposed) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto s : f->helper) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  if (f->sgd_done) cudaEventDestroy(f->sgd_done);
  if (f->comm_done) cudaEventDestroy(f->comm_done);
  if (f->pushed) cudaEventDestroy(f->pushed);
  delete f;
  return AGRS_OK;
}

static void mark(agrs_fused* f, const char* what, int bucket, int peer, cudaStream_t s) {
  if (!f->trace_step || f->step != f->trace_step) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, s);
  f->marks.push_back({what, bucket, peer, ev});
}

// stream that carries this rank's traffic to rank p: a copy stream per peer (round robin); local copies ride on the last one
static cudaStream_t stream_to(agrs_fused* f, int p) {
  const int world = f->args.world, rank = f->args.rank;
  const int j = (p - rank + world) % world;  // 0 = myself
  return f->copy[size_t(j == 0 ? f->copy.size() - 1 : (j - 1) % int(f->copy.size()))];
}

// One strided device-to-device copy, ordered on stream `s` like a plain cudaMemcpy2DAsync there, but cut into row chunks that run
// on helper streams when it is large: every stream is served by its own copy engine, and one engine alone does not fill an NVLink
// port (measured at N = 2, where each rank has a single peer: profiles/r2_fused_exchange.md).
static int copy_2d(agrs_fused* f, float* dst, size_t dpitch, const float* src, size_t spitch, size_t width, size_t rows, cudaStream_t s) {
  agrs_ctx* ctx = f->ctx;
  const size_t bytes = width * rows * sizeof(float);
  size_t parts = f->chunk_bytes && !f->helper.empty() ? bytes / f->chunk_bytes : 1;
  if (parts > f->helper.size() + 1) parts = f->helper.size() + 1;
  if (parts > rows) parts = rows;
  if (parts < 2) {
    AGRS_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch * sizeof(float), src, spitch * sizeof(float), width * sizeof(float), rows, cudaMemcpyDeviceToDevice, s));
    return AGRS_OK;
  }
  cudaEvent_t fork = f->ev_pool[f->ev_next++ % f->ev_pool.size()];
  AGRS_CUDA(ctx, cudaEventRecord(fork, s));
  const size_t per = (rows + parts - 1) / parts;
  for (size_t i = 0, r0 = 0; r0 < rows; ++i, r0 += per) {
    const size_t nr = r0 + per < rows ? per : rows - r0;
    cudaStream_t h = i == 0 ? s : f->helper[f->helper_next++ % f->helper.size()];
    if (i) AGRS_CUDA(ctx, cudaStreamWaitEvent(h, fork, 0));
    AGRS_CUDA(ctx, cudaMemcpy2DAsync(dst + r0 * dpitch, dpitch * sizeof(float), src + r0 * spitch, spitch * sizeof(float), width * sizeof(float), nr,
                                     cudaMemcpyDeviceToDevice, h));
    if (i) {
      cudaEvent_t join = f->ev_pool[f->ev_next++ % f->ev_pool.size()];
      AGRS_CUDA(ctx, cudaEventRecord(join, h));
      AGRS_CUDA(ctx, cudaStreamWaitEvent(s, join, 0));
    }
  }
  return AGRS_OK;
}

// The scatter half of the reduce-scatter on the copy engines (bucket protocol only): the gradient values [off, off + n) of the flat
// parameter space go from `src` (local memory, complete once everything enqueued on the compute stream so far has run) to slot
// [rank] of the rank that owns each granule.  The granules of one owner are world granules apart in src and contiguous in its
// slot: ONE strided 2-D copy per owner, plus 1-D copies for a ragged first / last granule.
static int scatter_dma(agrs_fused* f, const float* src, size_t off, size_t n) {
  agrs_ctx* ctx = f->ctx;
  const FusedArgs& a = f->args;
  const size_t G = size_t(1) << a.gshift, world = size_t(a.world), slice = a.total / world;
  cudaEvent_t ready = f->ev_pool[f->ev_next++ % f->ev_pool.size()];
  AGRS_CUDA(ctx, cudaEventRecord(ready, ctx->stream));
  mark(f, "gradient ready (compute stream)", -1, int(off >> a.gshift), ctx->stream);
  std::vector<char> waited(f->copy.size(), 0);
  auto stream_for = [&](size_t owner, cudaStream_t* s) -> int {
    *s = stream_to(f, int(owner));
    for (size_t i = 0; i < f->copy.size(); ++i)
      if (f->copy[i] == *s && !waited[i]) {
        AGRS_CUDA(ctx, cudaStreamWaitEvent(*s, ready, 0));
        waited[i] = 1;
      }
    return AGRS_OK;
  };
  f->segs.push_back({src, off, n, ready});
  const size_t me = size_t(a.rank);
  auto slot_at = [&](size_t e) {  // where flat element e lives in its owner's slot [rank]
    const size_t g = e >> a.gshift;
    return a.base[g % world] + a.g_off + size_t(a.rank) * slice + ((g / world) << a.gshift) + (e & (G - 1));
  };
  const size_t end = off + n;
  size_t e = off;
  if (e & (G - 1)) {  // ragged head: up to the next granule boundary
    const size_t stop = ((e >> a.gshift) + 1) << a.gshift, len = (stop < end ? stop : end) - e;
    cudaStream_t s;
    if ((e >> a.gshift) % world != me) {  // my own granules stay where they are: k_bucket_sgd reads them in place
      if (int rc = stream_for((e >> a.gshift) % world, &s)) return rc;
      AGRS_CUDA(ctx, cudaMemcpyAsync(slot_at(e), src + (e - off), len * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    e += len;
  }
  const size_t g0 = e >> a.gshift, full = (end - e) >> a.gshift;  // whole granules g0 .. g0 + full - 1
  for (size_t k = 0; k < world && k < full; ++k) {              // owner of granule g0 + k takes every world-th granule from there
    const size_t g = g0 + k, rows = (full - k + world - 1) / world;
    if (g % world == me) continue;
    cudaStream_t s;
    if (int rc = stream_for(g % world, &s)) return rc;
    if (int rc = copy_2d(f, slot_at(g << a.gshift), G, src + ((g << a.gshift) - off), G * world, G, rows, s)) return rc;
  }
  e += full << a.gshift;
  if (e < end && (e >> a.gshift) % world != me) {  // ragged tail
    cudaStream_t s;
    if (int rc = stream_for((e >> a.gshift) % world, &s)) return rc;
    AGRS_CUDA(ctx, cudaMemcpyAsync(slot_at(e), src + (e - off), (end - e) * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  f->scatter_pending = true;
  return AGRS_OK;
}

int agrs_fused_scatter_f32(agrs_fused* f, const float* src, size_t off, size_t n) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fused_scatter_f32");
  if (n == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, src && agrs_aligned16(src) && off % 4 == 0 && off + n <= f->args.total, AGRS_ERR_INVALID,
               "agrs_fused_scatter_f32: source must be 16-byte aligned and [off, off+n) a 4-aligned range of the gradient region");
  if (f->open && f->scatter_dma) return scatter_dma(f, src, off, n);
  size_t blocks = ((n + 3) / 4 + 255) / 256, cap = size_t(ctx->sm_count) * 8;
  k_scatter_grad<<<unsigned(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(f->args, src, off, n);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_fused_scatter_by_copy(agrs_fused* f) { return f && f->open && f->scatter_dma ? 1 : 0; }

// ---------------------------------------------------------------- (A) bucket protocol
int agrs_fused_begin_step(agrs_fused* f, float lr) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fused_begin_step");
  AGRS_NOT_CAPTURING(ctx, "agrs_fused_begin_step (flag generations are launch arguments)");
  AGRS_REQUIRE(ctx, !f->open, AGRS_ERR_INVALID, "agrs_fused_begin_step: the previous step was not finished");
  f->args.lr = lr;
  f->step = f->args.epoch + 1;
  f->open = true;
  f->buckets.clear();
  f->segs.clear();
  f->args.w_off = f->w_regions[f->cur];
  f->args.w_out = f->w_regions[1] ? f->w_regions[f->cur ^ 1] : f->args.w_off;
  if (f->trace_step && f->step == f->trace_step) {
    cudaEventCreate(&f->trace_base);
    cudaEventRecord(f->trace_base, ctx->stream);
  }
  return AGRS_OK;
}

static int wait_flags(agrs_fused* f, cudaStream_t s, const std::vector<size_t>& words) {
  agrs_ctx* ctx = f->ctx;
  unsigned* mine = reinterpret_cast<unsigned*>(f->args.base[f->args.rank]);
  if (f->wait32) {
    for (size_t w : words) {
      const CUresult r = f->wait32(s, CUdeviceptr(uintptr_t(mine + w)), f->step, CU_STREAM_WAIT_VALUE_GEQ);
      AGRS_REQUIRE(ctx, r == CUDA_SUCCESS, AGRS_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", int(r));
    }
    return AGRS_OK;
  }
  WaitList wl{};
  for (size_t w : words) wl.word[wl.n++] = static_cast<unsigned short>(w);
  k_wait_flags<<<1, 256, 0, s>>>(mine, wl, f->step, f->args.rank);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_fused_bucket_ready(agrs_fused* f, int bucket, size_t off, size_t n) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fused_bucket_ready");
  const FusedArgs& a = f->args;
  const size_t GW = (size_t(1) << a.gshift) * size_t(a.world);
  AGRS_REQUIRE(ctx, f->open, AGRS_ERR_INVALID, "agrs_fused_bucket_ready outside agrs_fused_begin_step / agrs_fused_finish_step");
  AGRS_REQUIRE(ctx, bucket >= 0 && bucket < kMaxBuckets && off % GW == 0 && n % GW == 0 && n > 0 && off + n <= a.total, AGRS_ERR_INVALID,
               "agrs_fused_bucket_ready: bucket %d [%zu, +%zu) must be a granule*world aligned range of the parameter space", bucket, off, n);
  for (int b : f->buckets) AGRS_REQUIRE(ctx, b != bucket, AGRS_ERR_INVALID, "agrs_fused_bucket_ready: bucket %d signalled twice in one step", bucket);
  f->buckets.push_back(bucket);
  const int world = a.world, rank = a.rank;
  if (f->scatter_dma) {
    // my gradients of this bucket travel on the copy streams: PF(bucket, me) at rank p follows them on the stream that goes to p
    for (int p = 0; p < world; ++p) {
      if (p == rank) continue;  // nothing travels to myself
      unsigned* word = reinterpret_cast<unsigned*>(a.base[p]) + PF(bucket, rank);
      cudaStream_t s = stream_to(f, p);
      mark(f, "scatter copies to peer done", bucket, p, s);
      if (f->write32) {
        const CUresult r = f->write32(s, CUdeviceptr(uintptr_t(word)), f->step, 0);
        AGRS_REQUIRE(ctx, r == CUDA_SUCCESS, AGRS_ERR_CUDA, "cuStreamWriteValue32 failed (%d)", int(r));
      } else {
        k_set_flag<<<1, 1, 0, s>>>(word, f->step);
        AGRS_LAUNCHED(ctx);
      }
    }
  } else {
    // compute stream: my pushes of this bucket are on their way -> PF(bucket, me) at every rank once they have landed
    k_signal<<<1, 32, 0, ctx->stream>>>(a, bucket, f->step, 0);
    AGRS_LAUNCHED(ctx);
  }
  // comm stream: every rank's pushes have landed in my slots
  std::vector<size_t> words;
  for (int q = 0; q < world; ++q)
    if (q != rank || !f->scatter_dma) words.push_back(PF(bucket, q));
  if (int rc = wait_flags(f, f->comm, words)) return rc;
  mark(f, "all push flags seen (comm stream)", bucket, -1, f->comm);
  // my weight region: the copies of the previous step must have left it before it is rewritten
  if (f->copies_pending) {
    for (auto ev : f->copy_done) AGRS_CUDA(ctx, cudaStreamWaitEvent(f->comm, ev, 0));
  }
  OwnSegs own{};
  if (f->scatter_dma) {
    for (auto& sg : f->segs) {
      const size_t lo = sg.off > off ? sg.off : off, hi = sg.off + sg.n < off + n ? sg.off + sg.n : off + n;
      if (lo >= hi) continue;  // the part of this gradient that lies in the bucket
      AGRS_REQUIRE(ctx, own.n < kMaxSegs, AGRS_ERR_UNSUPPORTED, "agrs_fused_bucket_ready: more than %d gradients in one bucket", kMaxSegs);
      own.src[own.n] = sg.src + (lo - sg.off);
      own.off[own.n] = lo;
      own.len[own.n] = hi - lo;
      ++own.n;
      AGRS_CUDA(ctx, cudaStreamWaitEvent(f->comm, sg.ready, 0));  // read in place: the kernels that produced them must be done
    }
  }
  const size_t rows = (n >> a.gshift) / size_t(world), per4 = (size_t(1) << a.gshift) / 4;
  size_t blocks = (rows * per4 + 255) / 256;
  const size_t cap = size_t(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  const bool push = f->ag_kernel && world > 1;
#define AGRS_BUCKET(W)                                                                              \
  do {                                                                                              \
    if (push) k_bucket_sgd<W, true><<<unsigned(blocks), 256, 0, f->comm>>>(a, off, n, own);         \
    else      k_bucket_sgd<W, false><<<unsigned(blocks), 256, 0, f->comm>>>(a, off, n, own);        \
  } while (0)
  switch (world) {
    case 2: AGRS_BUCKET(2); break;
    case 4: AGRS_BUCKET(4); break;
    case 8: AGRS_BUCKET(8); break;
    default: AGRS_BUCKET(0); break;
  }
#undef AGRS_BUCKET
  AGRS_LAUNCHED(ctx);
  mark(f, "slot sum + SGD done (comm stream)", bucket, -1, f->comm);
  if (world == 1) return AGRS_OK;
  if (push) {  // the kernel stored the new weights at every peer: one flag round after it
    k_signal<<<1, 32, 0, f->comm>>>(a, bucket, f->step, 1);
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  // all-gather with the copy engines: my granules of the bucket (every world-th granule) -> the same places at every peer
  AGRS_CUDA(ctx, cudaEventRecord(f->sgd_done, f->comm));
  const size_t G = size_t(1) << a.gshift;
  const size_t first = a.w_out + off + size_t(rank) * G;
  for (int j = 1; j < world; ++j) {
    const int p = (rank + j) % world;
    cudaStream_t s = stream_to(f, p);
    AGRS_CUDA(ctx, cudaStreamWaitEvent(s, f->sgd_done, 0));
    if (int rc = copy_2d(f, a.base[p] + first, GW, a.base[rank] + first, GW, G, rows, s)) return rc;
    mark(f, "all-gather copy to peer done", bucket, p, s);
    unsigned* word = reinterpret_cast<unsigned*>(a.base[p]) + WF(bucket, rank);
    if (f->write32) {
      const CUresult r = f->write32(s, CUdeviceptr(uintptr_t(word)), f->step, 0);
      AGRS_REQUIRE(ctx, r == CUDA_SUCCESS, AGRS_ERR_CUDA, "cuStreamWriteValue32 failed (%d)", int(r));
    } else {
      k_set_flag<<<1, 1, 0, s>>>(word, f->step);
      AGRS_LAUNCHED(ctx);
    }
  }
  f->copies_pending = true;
  return AGRS_OK;
}

int agrs_fused_finish_step(agrs_fused* f) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fused_finish_step");
  AGRS_REQUIRE(ctx, f->open, AGRS_ERR_INVALID, "agrs_fused_finish_step without agrs_fused_begin_step");
  const FusedArgs& a = f->args;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (ctx->profiling) {  // how long the compute stream sits in the waits below = the part of the exchange backward did not hide
    AGRS_CUDA(ctx, cudaEventCreate(&t0));
    AGRS_CUDA(ctx, cudaEventCreate(&t1));
    AGRS_CUDA(ctx, cudaEventRecord(t0, ctx->stream));
  }
  // my own slices: the comm stream's kernels; the copies leaving my weight region only matter to the NEXT rewrite (copy_done)
  AGRS_CUDA(ctx, cudaEventRecord(f->comm_done, f->comm));
  AGRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, f->comm_done, 0));
  if (f->copies_pending || f->scatter_pending)
    for (size_t i = 0; i < f->copy.size(); ++i) AGRS_CUDA(ctx, cudaEventRecord(f->copy_done[i], f->copy[i]));
  if (f->scatter_pending) {
    // the scatter copies read gradient buffers of the stream-ordered pool: the compute stream (which frees and reuses them) waits
    for (auto ev : f->copy_done) AGRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev, 0));
    f->scatter_pending = false;
    f->copies_pending = true;  // the events also cover this step's all-gather copies
  }
  // the peers' slices: WF(b, q) of every bucket and peer, polled by one small kernel on the (now idle) compute stream
  if (a.world > 1 && !f->buckets.empty()) {
    WaitList wl{};
    for (int b : f->buckets)
      for (int q = 0; q < a.world; ++q)
        if (q != a.rank) wl.word[wl.n++] = static_cast<unsigned short>(WF(b, q));
    k_wait_flags<<<1, 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned*>(a.base[a.rank]), wl, f->step, a.rank);
    AGRS_LAUNCHED(ctx);
  }
  if (t0) {
    AGRS_CUDA(ctx, cudaEventRecord(t1, ctx->stream));
    f->exposed.emplace_back(t0, t1);
  }
  if (f->trace_base) {
    cudaEvent_t fin;
    cudaEventCreate(&fin);
    cudaEventRecord(fin, ctx->stream);
    cudaDeviceSynchronize();
    float t = 0.f;
    for (auto& m : f->marks) {
      cudaEventElapsedTime(&t, f->trace_base, m.ev);
      fprintf(stderr, "agrs fused trace rank %d step %u: %9.3f ms  bucket %2d peer/granule %3d  %s\n", a.rank, f->step, t, m.bucket, m.peer, m.what);
      cudaEventDestroy(m.ev);
    }
    cudaEventElapsedTime(&t, f->trace_base, fin);
    fprintf(stderr, "agrs fused trace rank %d step %u: %9.3f ms  finish_step's waits satisfied (compute stream)\n", a.rank, f->step, t);
    if (t0) {
      cudaEventElapsedTime(&t, f->trace_base, t0);
      fprintf(stderr, "agrs fused trace rank %d step %u: %9.3f ms  backward's last kernel done, finish_step begins to wait (compute stream)\n", a.rank, f->step, t);
    }
    cudaEventDestroy(fin);
    cudaEventDestroy(f->trace_base);
    f->trace_base = nullptr;
    f->marks.clear();
  }
  f->args.epoch = f->step;
  f->open = false;
  if (f->w_regions[1]) {  // the regions swap roles: what this step wrote is what the next one reads
    f->cur ^= 1;
    f->args.w_off = f->w_regions[f->cur];
    f->args.w_out = f->args.w_off;
  }
  return AGRS_OK;
}

int agrs_fused_exposed_ms(agrs_fused* f, double* total_ms, uint64_t* steps) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fused_exposed_ms");
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double ms = 0.0;
  for (auto& pr : f->exposed) {
    float t = 0.f;
    AGRS_CUDA(ctx, cudaEventElapsedTime(&t, pr.first, pr.second));
    ms += t;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if (total_ms) *total_ms = ms;
  if (steps) *steps = f->exposed.size();
  f->exposed.clear();
  return AGRS_OK;
}

int agrs_fused_weights_offset(agrs_fused* f, size_t* w_off_now) {
  if (!f || !w_off_now) return AGRS_ERR_INVALID;
  *w_off_now = f->w_regions[f->cur];
  return AGRS_OK;
}

// ---------------------------------------------------------------- (B) one kernel after backward
int agrs_fused_sgd_step(agrs_fused* f, float lr) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fused_sgd_step");
  AGRS_NOT_CAPTURING(ctx, "agrs_fused_sgd_step (its barrier generation is a launch argument)");
  AGRS_REQUIRE(ctx, !f->open, AGRS_ERR_INVALID, "agrs_fused_sgd_step inside an open bucket step");
  f->args.lr = lr;
  f->args.epoch += 1;
  // Every CTA must be resident at once (they wait on each other): a COOPERATIVE launch of one wave sized by the occupancy of the
  // instantiation actually launched -- if the grid ever did not fit, the launch fails cleanly instead of deadlocking.
  const void* kern;
  switch (f->args.world) {
    case 2: kern = reinterpret_cast<const void*>(k_fused_dp_sgd<2>); break;
    case 4: kern = reinterpret_cast<const void*>(k_fused_dp_sgd<4>); break;
    case 8: kern = reinterpret_cast<const void*>(k_fused_dp_sgd<8>); break;
    default: kern = reinterpret_cast<const void*>(k_fused_dp_sgd<0>); break;
  }
  int per_sm = 0;
  AGRS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const size_t slice4 = f->args.total / 4 / size_t(f->args.world);
  size_t blocks = (slice4 + 255) / 256;
  const size_t cap = size_t(ctx->sm_count) * size_t(per_sm);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // the barrier counter arithmetic assumes the same grid on every call
  if (f->grid == 0) f->grid = blocks;
  AGRS_REQUIRE(ctx, f->grid == blocks, AGRS_ERR_INVALID, "agrs_fused_sgd_step: grid changed between calls");
  void* params[] = {&f->args};
  AGRS_CUDA(ctx, cudaLaunchCooperativeKernel(kern, dim3(unsigned(blocks)), dim3(256), params, 0, ctx->stream));
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

}  // extern "C"
