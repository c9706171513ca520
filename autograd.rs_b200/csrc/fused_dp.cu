// Fused data-parallel optimiser step over NVLink peer memory: reduce-scatter + SGD + all-gather without NCCL, replacing
//     all-reduce(grad)  ->  w -= lr * grad  (optim.rs:25-34)
// Every rank keeps its parameters and a gradient exchange region in an "arena" (one cudaMalloc block, IPC-exported, mapped by
// every peer at the same offsets).  The flat parameter space is dealt out in granules of 32 floats: granule g belongs to rank
// g % world (interleaved, so whatever part of it a kernel is working on, its traffic spreads over all ranks' NVLink ports);
// rank r OWNS the slice made of its granules and its gradient region holds one SLOT per source rank: [world][slice].
//     scatter     during backward every rank pushes its gradient values into slot [its rank] of the owner's region with P2P
//                 stores over NVLink -- straight from the dW GEMM's epilogue, tile by tile (gemm_tcgen05_2cta.cu), or with
//                 k_scatter_grad for gradients produced elsewhere (bias sums, conv filters)
//     barrier 1   every rank's pushes have landed
//     reduce      g = (slot 0 + slot 1 + ..) / world on the owned slice      LOCAL loads, rank order: deterministic
//     update      w = W_r[i] - lr * g                                       two roundings, as agrs_sgd_step_f32
//     broadcast   W_p[i] = w for every rank p                               P2P stores over NVLink, peers visited from rank+1 on
//     barrier 2   every rank's slice has landed everywhere; the slots may be overwritten by the next backward
// The barriers are flag arrays in the arenas, written with system-scope releases by the peers and polled with system-scope
// acquires; every wait is bounded (a lost rank traps the kernel instead of hanging the GPU).  One process per GPU: each
// rank's kernel runs on its own device, so the ranks always run concurrently.
#include <string.h>

#include "agrs_internal.h"

namespace {

constexpr int kMaxWorld = 16;
constexpr int kFlagFloats = 256;  // first 1 KB of every arena: barrier flags (unsigned), then the data regions
// Bound of every barrier spin: each poll is a system- or device-scope load (~0.5-1 us), so a rank that never shows up traps
// the kernel after roughly half a minute instead of hanging the GPU; honest skew between ranks is milliseconds.
constexpr unsigned kMaxSpins = 1u << 25;

struct FusedArgs {
  float* base[kMaxWorld];  // arena of every rank, as mapped in THIS process (base[rank] is the local one)
  int world, rank;
  size_t w_off, g_off;     // float offsets of the weight and gradient regions inside an arena
  size_t total;            // floats in each region; a multiple of 4 * world
  float lr;
  unsigned epoch;          // strictly increasing per call: barrier generation
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
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Cross-rank + grid barrier `which` (0 or 1).  Flags of an arena (unsigned words):
//   [which*32 + p]  "rank p reached barrier `which` of generation epoch"      written by rank p into EVERY arena
//   [64 + which]    "all ranks reached it": local go flag for the other CTAs  written by this rank's CTA 0
//   [66 + which]    CTA arrival counter of this rank (device scope)
__device__ void rank_barrier(const FusedArgs& a, int which) {
  unsigned* mine = reinterpret_cast<unsigned*>(a.base[a.rank]);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's loads/stores of the phase before the barrier are performed
    atomicAdd(mine + 66 + which, 1u);
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {  // wait for every CTA of this rank, then tell the peers
      const unsigned want = gridDim.x * a.epoch;  // the counter is never reset: generation g ends at g * gridDim.x
      for (unsigned spins = 0; ld_acquire_gpu(mine + 66 + which) < want; ++spins)
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
  const size_t shard4 = a.total / 4 / world;  // float4 per rank
  const float inv = 1.0f / float(world);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < shard4; i += stride) {
    // my slice is every world-th granule of 32 floats: local float index 4i lies in my granule (4i >> 5), which is granule
    // (4i >> 5) * world + rank of the flat parameter space
    const size_t e = ((((4 * i) >> 5) * size_t(world) + size_t(a.rank)) << 5) + ((4 * i) & 31);
    // The reduce-scatter already happened: every rank pushed its piece of MY slice into slot [its rank] of my gradient
    // region (from the dW GEMM's epilogue or k_scatter_grad).  What is left is a LOCAL sum over the slots, in rank order:
    // deterministic, every element has exactly one owner and its result is broadcast, so the replicas stay bit-identical.
    const float* slots = a.base[a.rank] + a.g_off + i * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (WORLD) {
      float4 v[WORLD ? WORLD : 1];
#pragma unroll
      for (int p = 0; p < (WORLD ? WORLD : 1); ++p) v[p] = ld_sys_v4(slots + size_t(p) * shard4 * 4);  // all loads in flight before the first add
#pragma unroll
      for (int p = 0; p < (WORLD ? WORLD : 1); ++p) { g.x += v[p].x; g.y += v[p].y; g.z += v[p].z; g.w += v[p].w; }
    } else {
      for (int p = 0; p < world; ++p) {
        const float4 v = ld_sys_v4(slots + size_t(p) * shard4 * 4);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
    }
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    float4 w = ld_sys_v4(a.base[a.rank] + a.w_off + e);
    w.x = __fsub_rn(w.x, __fmul_rn(g.x, a.lr));
    w.y = __fsub_rn(w.y, __fmul_rn(g.y, a.lr));
    w.z = __fsub_rn(w.z, __fmul_rn(g.z, a.lr));
    w.w = __fsub_rn(w.w, __fmul_rn(g.w, a.lr));
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

// The scatter half of the reduce-scatter for a gradient that was NOT produced by the scattering GEMM (bias sums, conv filters,
// second accumulations): element e of the gradient region belongs to rank e / shard, which keeps one slot per source rank;
// this rank's values go to slot [rank] of the owner with P2P stores.  `off` and `n` are multiples of 4.
__global__ void __launch_bounds__(256) k_scatter_grad(const FusedArgs a, const float* __restrict__ src, size_t off, size_t n) {
  const size_t shard = a.total / size_t(a.world);
  const size_t stride = size_t(gridDim.x) * blockDim.x, n4 = (n + 3) / 4;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const size_t e = off + 4 * i;  // a multiple of 4: a float4 never straddles two granules of 32
    const size_t g = e >> 5, owner = g % size_t(a.world), idx = ((g / size_t(a.world)) << 5) + (e & 31);
    float* dst = a.base[owner] + a.g_off + size_t(a.rank) * shard + idx;
    if (4 * i + 3 < n) {
      st_sys_v4(dst, __ldcs(reinterpret_cast<const float4*>(src) + i));
    } else {
      for (size_t j = 4 * i; j < n; ++j) asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + (j - 4 * i)), "f"(src[j]) : "memory");
    }
  }
}

}  // namespace

struct agrs_fused {
  agrs_ctx* ctx = nullptr;
  FusedArgs args{};
  size_t grid = 0;  // CTAs per launch: fixed at the first call (the barrier counters assume it)
};

// what a scattering GEMM epilogue needs to know (gemm_tcgen05_2cta.cu)
void agrs_fused_scatter_target(const agrs_fused* f, float* const** bases, int* world, int* rank, size_t* g_off, size_t* shard) {
  *bases = f->args.base;
  *world = f->args.world;
  *rank = f->args.rank;
  *g_off = f->args.g_off;
  *shard = f->args.total / size_t(f->args.world);
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
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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

int agrs_fused_create(agrs_ctx* ctx, int world, int rank, void* const* arenas, size_t w_off, size_t g_off, size_t total, agrs_fused** out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, out && arenas && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, AGRS_ERR_INVALID,
               "agrs_fused_create: bad arguments (world <= %d)", kMaxWorld);
  AGRS_REQUIRE(ctx, total % (32 * size_t(world)) == 0 && w_off % 32 == 0 && g_off % 32 == 0 && w_off >= kFlagFloats && g_off >= kFlagFloats,
               AGRS_ERR_INVALID, "agrs_fused_create: regions must start after the %d flag words, on 32-float boundaries, and hold a multiple of 32*world floats", kFlagFloats);
  agrs_fused* f = new agrs_fused();
  f->ctx = ctx;
  for (int p = 0; p < world; ++p) {
    AGRS_REQUIRE(ctx, arenas[p] && agrs_aligned16(arenas[p]), AGRS_ERR_INVALID, "agrs_fused_create: arena %d is null or unaligned", p);
    f->args.base[p] = static_cast<float*>(arenas[p]);
  }
  f->args.world = world;
  f->args.rank = rank;
  f->args.w_off = w_off;
  f->args.g_off = g_off;
  f->args.total = total;
  f->args.epoch = 0;
  *out = f;
  return AGRS_OK;
}

int agrs_fused_destroy(agrs_fused* f) {
  if (!f) return AGRS_OK;
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

int agrs_fused_sgd_step(agrs_fused* f, float lr) {
  if (!f) return AGRS_ERR_INVALID;
  agrs_ctx* ctx = f->ctx;
  AGRS_CHECK_CTX(ctx);
  AGRS_NOT_CAPTURING(ctx, "agrs_fused_sgd_step (its barrier generation is a launch argument)");
  f->args.lr = lr;
  f->args.epoch += 1;
  // Every CTA must be resident at once (they wait on each other): one wave, sized by the occupancy the runtime reports.
  int per_sm = 0;
  AGRS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_dp_sgd<0>, 256, 0));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const size_t shard4 = f->args.total / 4 / size_t(f->args.world);
  size_t blocks = (shard4 + 255) / 256;
  const size_t cap = size_t(ctx->sm_count) * size_t(per_sm);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // the barrier counter arithmetic assumes the same grid on every call
  static_assert(kFlagFloats >= 68, "flag words");
  if (f->grid == 0) f->grid = blocks;
  AGRS_REQUIRE(ctx, f->grid == blocks, AGRS_ERR_INVALID, "agrs_fused_sgd_step: grid changed between calls");
  switch (f->args.world) {
    case 2: k_fused_dp_sgd<2><<<unsigned(blocks), 256, 0, ctx->stream>>>(f->args); break;
    case 4: k_fused_dp_sgd<4><<<unsigned(blocks), 256, 0, ctx->stream>>>(f->args); break;
    case 8: k_fused_dp_sgd<8><<<unsigned(blocks), 256, 0, ctx->stream>>>(f->args); break;
    default: k_fused_dp_sgd<0><<<unsigned(blocks), 256, 0, ctx->stream>>>(f->args); break;
  }
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

}  // extern "C"
