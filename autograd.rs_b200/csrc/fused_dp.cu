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
//      and the backward of the earlier layers keeps the tensor cores busy meanwhile.  agrs_fused_finish_step makes the compute
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

// slot sum in rank order + SGD on element e4 (float4) of the owned slice index idx4; returns the new weights
template <int WORLD>
__device__ __forceinline__ float4 reduce_and_update(const FusedArgs& a, int world, size_t slice, size_t idx, size_t e, float inv) {
  const float* slots = a.base[a.rank] + a.g_off + idx;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (WORLD) {
    float4 v[WORLD ? WORLD : 1];
#pragma unroll
    for (int p = 0; p < (WORLD ? WORLD : 1); ++p) v[p] = ld_sys_v4(slots + size_t(p) * slice);  // all loads in flight before the first add
#pragma unroll
    for (int p = 0; p < (WORLD ? WORLD : 1); ++p) { g.x += v[p].x; g.y += v[p].y; g.z += v[p].z; g.w += v[p].w; }
  } else {
    for (int p = 0; p < world; ++p) {
      const float4 v = ld_sys_v4(slots + size_t(p) * slice);
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
__global__ void __launch_bounds__(256) k_bucket_sgd(const FusedArgs a, size_t off, size_t n) {
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
    const float4 w = reduce_and_update<WORLD>(a, world, slice, idx, e, inv);
    st_sys_v4(a.base[a.rank] + a.w_off + e, w);
    if (PUSH) {
#pragma unroll
      for (int jj = 1; jj < (WORLD ? WORLD : kMaxWorld); ++jj) {
        if (jj >= world) break;
        int p = a.rank + jj;
        p -= p >= world ? world : 0;
        st_sys_v4(a.base[p] + a.w_off + e, w);
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
  bool ag_kernel = false;                      // all-gather with P2P stores from k_bucket_sgd instead of the copy engines
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
  if (!ptr) return AGRS_OK;
  AGRS_CUDA(ctx, cudaDeviceSynchronize());  // comm / copy streams of an exchange that used the arena included
  AGRS_CUDA(ctx, cudaFree(ptr));
  return AGRS_OK;
}

int agrs_peer_export(agrs_ctx* ctx, void* ptr, void* handle64) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, ptr && handle64, AGRS_ERR_INVALID, "agrs_peer_export: null pointer");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == AGRS_PEER_HANDLE_BYTES, "IPC handle size");
  AGRS_CUDA(ctx, cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return AGRS_OK;
}

int agrs_peer_open(agrs_ctx* ctx, const void* handle64, void** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, handle64 && out, AGRS_ERR_INVALID, "agrs_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  AGRS_CUDA(ctx, cudaSetDevice(ctx->device));
  AGRS_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return AGRS_OK;
}

int agrs_peer_close(agrs_ctx* ctx, void* ptr) {
  AGRS_CHECK_CTX(ctx);
  if (ptr) AGRS_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
  return AGRS_OK;
}

size_t agrs_fused_flag_floats(void) { return kFlagFloats; }
int agrs_fused_max_buckets(void) { return kMaxBuckets; }

int agrs_fused_create(agrs_ctx* ctx, int world, int rank, void* const* arenas, size_t w_off, size_t g_off, size_t total, int granule_log2,
                      agrs_fused** out) {
  AGRS_CHECK_CTX(ctx);
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
  f->args.w_off = w_off;
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
  if (f->sgd_done) cudaEventDestroy(f->sgd_done);
  if (f->comm_done) cudaEventDestroy(f->comm_done);
  if (f->pushed) cudaEventDestroy(f->pushed);
  delete f;
  return AGRS_OK;
}

int agrs_fused_scatter_f32(agrs_fused* f, const float* src, size_t off, size_t n) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  if (n == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, src && agrs_aligned16(src) && off % 4 == 0 && off + n <= f->args.total, AGRS_ERR_INVALID,
               "agrs_fused_scatter_f32: source must be 16-byte aligned and [off, off+n) a 4-aligned range of the gradient region");
  size_t blocks = ((n + 3) / 4 + 255) / 256, cap = size_t(ctx->sm_count) * 8;
  k_scatter_grad<<<unsigned(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(f->args, src, off, n);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

// ---------------------------------------------------------------- (A) bucket protocol
int agrs_fused_begin_step(agrs_fused* f, float lr) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_NOT_CAPTURING(ctx, "agrs_fused_begin_step (flag generations are launch arguments)");
  AGRS_REQUIRE(ctx, !f->open, AGRS_ERR_INVALID, "agrs_fused_begin_step: the previous step was not finished");
  f->args.lr = lr;
  f->step = f->args.epoch + 1;
  f->open = true;
  f->buckets.clear();
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
  const FusedArgs& a = f->args;
  const size_t GW = (size_t(1) << a.gshift) * size_t(a.world);
  AGRS_REQUIRE(ctx, f->open, AGRS_ERR_INVALID, "agrs_fused_bucket_ready outside agrs_fused_begin_step / agrs_fused_finish_step");
  AGRS_REQUIRE(ctx, bucket >= 0 && bucket < kMaxBuckets && off % GW == 0 && n % GW == 0 && n > 0 && off + n <= a.total, AGRS_ERR_INVALID,
               "agrs_fused_bucket_ready: bucket %d [%zu, +%zu) must be a granule*world aligned range of the parameter space", bucket, off, n);
  for (int b : f->buckets) AGRS_REQUIRE(ctx, b != bucket, AGRS_ERR_INVALID, "agrs_fused_bucket_ready: bucket %d signalled twice in one step", bucket);
  f->buckets.push_back(bucket);
  const int world = a.world, rank = a.rank;
  // compute stream: my pushes of this bucket are on their way -> PF(bucket, me) at every rank once they have landed
  k_signal<<<1, 32, 0, ctx->stream>>>(a, bucket, f->step, 0);
  AGRS_LAUNCHED(ctx);
  // comm stream: every rank's pushes have landed in my slots
  std::vector<size_t> words;
  for (int q = 0; q < world; ++q) words.push_back(PF(bucket, q));
  if (int rc = wait_flags(f, f->comm, words)) return rc;
  // my weight region: the copies of the previous step must have left it before it is rewritten
  if (f->copies_pending) {
    for (auto ev : f->copy_done) AGRS_CUDA(ctx, cudaStreamWaitEvent(f->comm, ev, 0));
  }
  const size_t rows = (n >> a.gshift) / size_t(world), per4 = (size_t(1) << a.gshift) / 4;
  size_t blocks = (rows * per4 + 255) / 256;
  const size_t cap = size_t(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  const bool push = f->ag_kernel && world > 1;
#define AGRS_BUCKET(W)                                                                              \
  do {                                                                                              \
    if (push) k_bucket_sgd<W, true><<<unsigned(blocks), 256, 0, f->comm>>>(a, off, n);              \
    else      k_bucket_sgd<W, false><<<unsigned(blocks), 256, 0, f->comm>>>(a, off, n);             \
  } while (0)
  switch (world) {
    case 2: AGRS_BUCKET(2); break;
    case 4: AGRS_BUCKET(4); break;
    case 8: AGRS_BUCKET(8); break;
    default: AGRS_BUCKET(0); break;
  }
#undef AGRS_BUCKET
  AGRS_LAUNCHED(ctx);
  if (world == 1) return AGRS_OK;
  if (push) {  // the kernel stored the new weights at every peer: one flag round after it
    k_signal<<<1, 32, 0, f->comm>>>(a, bucket, f->step, 1);
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  // all-gather with the copy engines: my granules of the bucket (every world-th granule) -> the same places at every peer
  AGRS_CUDA(ctx, cudaEventRecord(f->sgd_done, f->comm));
  const size_t G = size_t(1) << a.gshift;
  const size_t first = a.w_off + off + size_t(rank) * G;
  for (int j = 1; j < world; ++j) {
    const int p = (rank + j) % world;
    cudaStream_t s = f->copy[size_t(j - 1) % f->copy.size()];
    AGRS_CUDA(ctx, cudaStreamWaitEvent(s, f->sgd_done, 0));
    AGRS_CUDA(ctx, cudaMemcpy2DAsync(a.base[p] + first, GW * sizeof(float), a.base[rank] + first, GW * sizeof(float), G * sizeof(float), rows,
                                     cudaMemcpyDeviceToDevice, s));
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
  AGRS_REQUIRE(ctx, f->open, AGRS_ERR_INVALID, "agrs_fused_finish_step without agrs_fused_begin_step");
  const FusedArgs& a = f->args;
  // my own slices: the comm stream's kernels; the copies leaving my weight region only matter to the NEXT rewrite (copy_done)
  AGRS_CUDA(ctx, cudaEventRecord(f->comm_done, f->comm));
  AGRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, f->comm_done, 0));
  if (f->copies_pending)
    for (size_t i = 0; i < f->copy.size(); ++i) AGRS_CUDA(ctx, cudaEventRecord(f->copy_done[i], f->copy[i]));
  // the peers' slices: WF(b, q) of every bucket and peer, polled by one small kernel on the (now idle) compute stream
  if (a.world > 1 && !f->buckets.empty()) {
    WaitList wl{};
    for (int b : f->buckets)
      for (int q = 0; q < a.world; ++q)
        if (q != a.rank) wl.word[wl.n++] = static_cast<unsigned short>(WF(b, q));
    k_wait_flags<<<1, 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned*>(a.base[a.rank]), wl, f->step, a.rank);
    AGRS_LAUNCHED(ctx);
  }
  f->args.epoch = f->step;
  f->open = false;
  return AGRS_OK;
}

// ---------------------------------------------------------------- (B) one kernel after backward
int agrs_fused_sgd_step(agrs_fused* f, float lr) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
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
