// PTX wrappers, UMMA descriptors and tile scheduling shared by the tcgen05 GEMM kernels (gemm_tcgen05.cu: one CTA per
// tile; gemm_tcgen05_2cta.cu: cta_group::2 CTA pairs).  Device-only helpers; included inside namespace tc.
#pragma once
#include <cuda.h>

#include "agrs_internal.h"

namespace tc {

constexpr int BM = 128;    // accumulator rows per CTA (TMEM lanes)
constexpr int BK = 32;     // 32 f32 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 8;  // kind::tf32

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("agrs tcgen05 gemm: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
// AGRS_WAIT_NS > 0: every wait carries that suspend-time hint (see mbar_wait_relaxed below); 0: plain try_wait, re-polled in a loop
#ifndef AGRS_WAIT_NS
#define AGRS_WAIT_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spins = 0;; ++spins) {
#if AGRS_WAIT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(uint32_t(AGRS_WAIT_NS))
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
    if (done) break;
    if (spins > (1u << 24)) mbar_timeout(bar, parity);
  }
}
// The same wait for warps that expect to wait LONG (an epilogue warp waits a whole K chunk for its accumulator, the TMA producer is
// several stages ahead): the suspend-time hint lets the hardware park the warp for up to `ns` nanoseconds per try instead of
// returning to the polling loop every ~100 cycles -- the polls of a waiting warp are issue slots and power the splitter warps on
// the same sub-partition want (38 % of all executed instructions of the fp16 x3 kernel were polling loops).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    if (done) break;
    if (spins > (1u << 24)) mbar_timeout(bar, parity);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// the same load with an L2 eviction policy (createpolicy): an operand that the next group of tile rows reads again can be kept
// (evict_last), one that streams through can be marked evict_first
__device__ __forceinline__ void tma_load_2d_hint(const CUtensorMap* map, uint32_t bar, uint32_t dst, int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (SM100 UMMA), version 1.
//   K-major  tile: rows of 128 B (32 f32 of K), SWIZZLE_128B (16-B chunks ^ row%8), 8-row atoms 1024 B apart
//                  -> layout type 2, SBO = 1024, LBO unused (1)
//   MN-major tile: 32-bit operands must use SWIZZLE_128B_BASE32B (32-B chunks ^ row%4; TMA swizzle 128B_ATOM_32B):
//                  blocks of [32 K-rows x 128 B (32 f32 of M/N)] = 4096 B per 32 of M/N -> layout type 1,
//                  LBO = 4096 (next 32 of M/N), SBO = 512 (next 4 rows of K)
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);                    // start address, bits [0,14)
  d |= uint64_t(MN_MAJOR ? (4096u >> 4) : 1u) << 16;            // leading byte offset, bits [16,30)
  d |= uint64_t(MN_MAJOR ? (512u >> 4) : (1024u >> 4)) << 32;   // stride byte offset, bits [32,46)
  d |= uint64_t(1) << 46;                                       // descriptor version (Blackwell)
  d |= uint64_t(MN_MAJOR ? 1 : 2) << 61;                        // layout type: SWIZZLE_128B_BASE32B / SWIZZLE_128B
  return d;
}
// bytes to advance the descriptor start address per UMMA_K (=8 f32) slice of the 32-wide k-block
template <bool MN_MAJOR>
__device__ __forceinline__ constexpr uint32_t k_slice_bytes() {
  return MN_MAJOR ? 1024u : 32u;
}

// Instruction descriptor: D=f32, A=B=tf32, majors, N>>3 at [17,23), M>>4 at [24,29)
template <bool A_MN, bool B_MN, int BN, int MMA_M = BM>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(A_MN) << 15) | (uint32_t(B_MN) << 16) | (uint32_t(BN >> 3) << 17) |
         (uint32_t(MMA_M >> 4) << 24);
}

// x -> lo = tf32(x - trunc_tf32(x));  optionally also returns the truncated hi
__device__ __forceinline__ float split_lo(float x, float& hi) {
  uint32_t u = __float_as_uint(x);
  hi = __uint_as_float(u & 0xFFFFE000u);
  float lo = x - hi;
  if ((u & 0x7F800000u) == 0x7F800000u) lo = 0.f;  // inf/NaN: keep them in `hi` only
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(lo));
  return __uint_as_float(r);
}

__device__ __forceinline__ void tile_coords(int64_t t, int64_t m_tiles, int64_t n_tiles, int64_t& mb, int64_t& nb) {
#ifndef AGRS_TILE_GROUP_M
#define AGRS_TILE_GROUP_M 8
#endif
  constexpr int64_t GROUP_M = AGRS_TILE_GROUP_M;  // walk 8 row-blocks per column sweep: concurrent CTAs share A rows and B columns in L2
  int64_t per_group = GROUP_M * n_tiles;
  int64_t g = t / per_group, r = t % per_group;
  int64_t first = g * GROUP_M;
  int64_t gsize = (m_tiles - first < GROUP_M) ? m_tiles - first : GROUP_M;
  mb = first + r % gsize;
  nb = r / gsize;
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int resolve_encode(agrs_ctx* ctx) {
  if (ctx->encode_tiled) return AGRS_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  AGRS_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  AGRS_REQUIRE(ctx, fn && qres == cudaDriverEntryPointSuccess, AGRS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  ctx->encode_tiled = fn;
  return AGRS_OK;
}

// 2-D f32 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements; 128-byte swizzle
static inline int encode_2d(agrs_ctx* ctx, CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer, bool mn_major) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AGRS_REQUIRE(ctx, r == CUDA_SUCCESS, AGRS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", int(r),
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return AGRS_OK;
}

// 3-D f32 tensor map of an output [planes][rows][cols] (row pitch `ld` elements, planes `rows * ld` apart) for TMA stores of
// 32 x 32 pieces staged 128-byte-swizzled in shared memory; pieces that cross the last row or column are clipped by the hardware
static inline int encode_3d_out(agrs_ctx* ctx, CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint64_t planes, uint64_t ld) {
  cuuint64_t dims[3] = {cols, rows, planes};
  cuuint64_t strides[2] = {ld * sizeof(float), rows * ld * sizeof(float)};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AGRS_REQUIRE(ctx, r == CUDA_SUCCESS, AGRS_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (%d) cols=%llu rows=%llu planes=%llu ld=%llu", int(r),
               (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)planes, (unsigned long long)ld);
  return AGRS_OK;
}

}  // namespace tc
