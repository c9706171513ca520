// f32-faithful split GEMM, CTA-pair version: tcgen05.mma.cta_group::2 on a 256 x 256 pair tile.
//
// Same contract as gemm_tcgen05.cu (Tensor::dot's CUDA arm, reference src/tensor/tensor.rs:319; the three GEMMs of
// Linear::forward/backward, operators.rs:144,159,160).  Numerics: hi = raw f32 container read by the tensor core as tf32,
// lo = x - hi; a*b ~ a_lo*b + a*b_lo + a_hi*b_hi accumulated in TMEM, the two correction terms on bf16 operands
// (template parameter CROSS = 1 / 2, the default) or on tf32 operands (CROSS = 0, "3xTF32"); K cut into chunks that the
// epilogue folds into C with round-to-nearest adds.  Data flow (the one-CTA kernel is bound by shared-memory bandwidth and by
// its two-deep pipeline, profiles/r1_gemm_v1_ncu_full_summary.txt: tensor pipe 47 % active):
//   * a cluster of two CTAs (two SMs of one TPC) owns a 256 x 256 tile; each CTA stages 128 rows of A and 128 columns
//     of B per 32-wide k-block, so every operand byte in shared memory feeds twice the math;
//   * raw f32 tiles (TMA) and lo / bf16 tiles (splitter warps) live in separate rings: 4 raw stages cover the TMA latency, 2
//     splitter stages cover the split; with 32 KB of epilogue staging that is 224 KB of shared memory per CTA;
//   * the epilogue hands 32 x 32 pieces to the TMA (tensor store, then reduce-add per later K chunk) instead of storing one row
//     per lane from registers (see tma_store_3d).
// Per CTA, 16 warps: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator,
// warps 4-7 epilogue, warps 8-15 splitters.  Cross-CTA signalling: splitters and epilogue warps of both CTAs arrive on
// the LEADER's mbarriers (CTA-scope release); the leader's tcgen05.commit multicasts to the same barrier in both CTAs.
#include "tc_common.cuh"
#include "act.cuh"
#include "mag.cuh"

namespace tc2 {
using namespace tc;

constexpr int kThreads = 512;
constexpr int kEpiWarp0 = 4;
constexpr int kActWarp0 = 2;  // ACT instantiations: warps 2 and 3 (idle otherwise once TMEM is allocated) apply the activation
constexpr int kActWarps = 2;
constexpr int kSplitWarp0 = 8;
constexpr int kSplitWarps = 8;
#ifndef AGRS_SPLIT_GROUPS
#define AGRS_SPLIT_GROUPS 2
#endif
constexpr int kSplitGroups = AGRS_SPLIT_GROUPS;  // groups of splitter warps that alternate k-blocks
constexpr int BN = 256;   // pair-tile width
constexpr int BNH = 128;  // columns of B staged per CTA
#ifndef AGRS_PAIR_RAW_STAGES
#define AGRS_PAIR_RAW_STAGES 4
#endif
#ifndef AGRS_PAIR_LO_STAGES
#define AGRS_PAIR_LO_STAGES 2
#endif
constexpr int R = AGRS_PAIR_RAW_STAGES;  // raw (TMA) stages
constexpr int L = AGRS_PAIR_LO_STAGES;   // lo (splitter) stages
// A splitter group waits on a stage barrier one phase after the phase it last saw complete only if the stage's previous user is not
// newer than the group's own previous k-block (mbarrier parity waits alias two phases apart): groups <= ring depth, for both rings.
static_assert(kSplitGroups <= R && kSplitGroups <= L, "splitter groups must not outnumber the stages of either ring");
constexpr uint32_t A_BYTES = BM * BK * 4;    // 16 KB
constexpr uint32_t B_BYTES = BNH * BK * 4;   // 16 KB
constexpr uint32_t RAW_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t EPI_TILE = 32 * 32 * 4;         // one 32 x 32 f32 piece of C staged for a TMA store
constexpr uint32_t EPI_BYTES = 4 * 2 * EPI_TILE;   // two buffers per epilogue warp
constexpr uint32_t SMEM_BYTES = (R + L) * RAW_BYTES + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 512;  // two accumulator stages of 256 columns
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of the leader CTA.  Default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive does:
// a release at cluster scope compiles to MEMBAR.ALL.GPU + ERRBAR and cost the splitter warps ~20 % of their time
// (profiles/r1_gemm_pair_ncu_full_summary.txt).  What the leader's MMA reads afterwards is THIS CTA's shared memory, written by this
// warp and already made visible to the async proxy by fence.proxy.async, so CTA scope is what the data needs.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a LOCAL barrier whose arrivals come from both CTAs of the cluster
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(uint16_t(3))
               : "memory");
}

// ---- epilogue through shared memory and TMA ---------------------------------------------------------------------------------
// A thread of an epilogue warp holds 32 consecutive columns of ONE row (tcgen05.ld 32x32b), so a direct st.global.v4 touches 32
// different 128-byte lines per warp instruction: profiles/r1_gemm_bf16cross_ncu_summary.txt has those stores and the read-modify-
// write of the K chunks at 44 % of the LSU data-pipe wavefronts, the pipe the splitter warps live on.  Instead each warp stages
// its 32 x 32 piece in shared memory (128-byte-swizzled, conflict-free) and one lane hands it to the TMA: a tensor store for the
// first K chunk, a tensor reduce-add (f32 add in L2, round-to-nearest, one writer per element: deterministic) for the later ones.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // at most N of this thread's bulk groups still read their shared-memory source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }  // ... and are performed

__device__ __forceinline__ float lo_of(float x) {
  const uint32_t u = __float_as_uint(x);
  const float hi = __uint_as_float(u & 0xFFFFE000u);  // what the tensor core reads from the raw container
  // exact; an inf input gives inf - inf = NaN here, so its products are NaN where plain f32 says inf (documented in agrs.h: the
  // other correction term already multiplies inf by a lo part that is usually exactly zero)
  return x - hi;
}

// ---- cross terms in bf16 (CROSS = 1, 2) -----------------------------------------------------------------------------------
// The two correction terms a_lo*b and a*b_lo are 2^-10 of the product, so 8 significant bits on each factor keep them to
// ~2^-19 of the product -- and kind::f16 runs at twice the tf32 rate with half the operand bytes: a k-block costs
// 4 (hi*hi, tf32) + 2 + 2 (bf16) MMA slots instead of 12.  The splitter warps therefore write, per k-block and operand, two
// K-major bf16 tiles of [128 rows][32 k] (8 KB each): lo = bf16(x - trunc_tf32(x)) and x16 = bf16(x); the raw f32 tile still
// feeds the hi*hi pass untouched.  The splitters transpose while they convert, so the bf16 tiles are K-major whatever the
// storage order of the f32 operand.  Tile layouts (both canonical UMMA K-major forms):
//   CROSS 1  SWIZZLE_64B : row r at r*64, its 16-byte chunk j (k = 8j..8j+7) at chunk slot j ^ ((r >> 1) & 3); SBO = 512
//   CROSS 2  INTERLEAVE  : chunk j of row r at j*2048 + r*16 (core matrices of 8 rows x 16 B); SBO = 128, LBO = 2048
constexpr uint32_t BF_TILE = BM * BK * 2;  // 8 KB
static_assert(BM == BNH, "one splitter thread per row of A and per row of B");

//   CROSS 3  "bf16x3": no tf32 pass at all -- x = x1 + x2 with x1 = bf16(x), x2 = bf16(x - x1); a*b ~ a2*b1 + a1*b2 + a1*b1, three
//            bf16 MMAs = 6 slots per k-block instead of 8.  Same tiles as CROSS 1 (x2 in the lo tile, x1 in the x16 tile).
//            Representation error ~4e-6 of max|C| instead of ~1e-6 (tools/split_numerics.py): opt-in.
template <int CROSS>
__device__ __forceinline__ uint32_t bf_chunk_offset(uint32_t row, uint32_t j) {
  return CROSS != 2 ? row * 64u + ((j ^ ((row >> 1) & 3u)) << 4) : j * 2048u + row * 16u;
}
// descriptor of the i-th 16-wide K slice (kind::f16 UMMA_K) of a bf16 tile
template <int CROSS>
__device__ __forceinline__ uint64_t make_bf_desc(uint32_t tile, int i) {
  constexpr bool kSw64 = CROSS != 2;
  const uint32_t addr = tile + (kSw64 ? uint32_t(i) * 32u : uint32_t(i) * 4096u);
  uint64_t d = 0;
  d |= uint64_t((addr & 0x3FFFF) >> 4);
  d |= uint64_t(kSw64 ? 1u : (2048u >> 4)) << 16;          // LBO: unused with a swizzle / next 16-byte chunk of K
  d |= uint64_t(kSw64 ? (512u >> 4) : (128u >> 4)) << 32;  // SBO: next 8 rows
  d |= uint64_t(1) << 46;
  d |= uint64_t(kSw64 ? 4 : 0) << 61;                      // SWIZZLE_64B / no swizzle
  return d;
}
// D = f32, A = B = bf16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t((2 * BM) >> 4) << 24);

__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float e0, float e1) {  // e0 at the lower address
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16(float e0, float e1) {  // e0 at the lower address
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}
__device__ __forceinline__ void unpack_f16(uint32_t p, float& e0, float& e1) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(e0), "=f"(e1) : "r"(p));
}
// One row (32 k) of a raw f32 operand tile -> the same row of its two 16-bit tiles (`mult`: the row's scale, CROSS 4 only).
//   raw K-major  tile: [128 rows][128 B], SWIZZLE_128B: 16-byte chunk c of row r sits in slot c ^ (r & 7)
//   raw MN-major tile: 4 blocks of [32 k][32 rows], SWIZZLE_128B_ATOM_32B: 32-byte chunk c of k-line k sits in slot c ^ (k & 3)
template <bool MN_MAJOR, int CROSS>
__device__ __forceinline__ void split_row_bf16(uint32_t raw, uint32_t out_lo, uint32_t out_x, uint32_t row, float mult) {
  float v[BK];
  if (MN_MAJOR) {
    const uint32_t base = raw + (row >> 5) * 4096u + (row & 7u) * 4u, c = (row & 31u) >> 3;
#pragma unroll
    for (uint32_t k = 0; k < uint32_t(BK); ++k)
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k]) : "r"(base + k * 128u + ((c ^ (k & 3u)) << 5)));
  } else {
    const uint32_t base = raw + row * 128u;
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c)
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3])
                   : "r"(base + ((c ^ (row & 7u)) << 4)));
  }
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) {
    uint32_t lo[4], x[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float e0 = v[8 * j + 2 * p], e1 = v[8 * j + 2 * p + 1];
      if (CROSS == 4) {  // fp16 pair of the scaled value: h = fp16(x'), l = fp16(x' - h)
        const float s0 = e0 * mult, s1 = e1 * mult;
        float h0, h1;
        x[p] = pack_f16(s0, s1);
        unpack_f16(x[p], h0, h1);
        lo[p] = pack_f16(s0 - h0, s1 - h1);
        continue;
      }
      x[p] = pack_bf16(e0, e1);
      if (CROSS == 3) {  // second bf16 term: what the first one left, exactly (a bf16 is the top half of an f32)
        lo[p] = pack_bf16(e0 - __uint_as_float(x[p] << 16), e1 - __uint_as_float(x[p] & 0xFFFF0000u));
      } else {
        lo[p] = pack_bf16(lo_of(e0), lo_of(e1));
      }
    }
    const uint32_t off = bf_chunk_offset<CROSS>(row, j);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(out_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(out_x + off), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
  }
}

#ifndef AGRS_SPLIT_EARLY_CONVERT
#define AGRS_SPLIT_EARLY_CONVERT 1
#endif
// Packed f32x2 arithmetic (FMUL2 / FADD2, sm_100) in the 16-bit splits: scale and residual of two neighbouring elements per
// instruction -- 6 instead of 8 instructions per element pair in the fp16 mode, 5 instead of 6 in bf16 x3.  Same roundings.
#ifndef AGRS_SPLIT_F32X2
#define AGRS_SPLIT_F32X2 1
#endif
// In the all-16-bit modes (CROSS >= 3) the tensor core never reads the raw f32 tile: the splitter warps that converted it hand the
// raw stage back to the TMA producer themselves, one or two k-block periods before the MMAs that consume the converted tiles
// complete, so the same number of raw stages covers that much more load latency.
#ifndef AGRS_RAW_EARLY_RELEASE
#define AGRS_RAW_EARLY_RELEASE 1
#endif
// Suspend-time hints (ns) of the long waits: epilogue warps on their accumulator stage, TMA producer on a free raw stage.  0 = plain polling.
#ifndef AGRS_EPI_WAIT_NS
#define AGRS_EPI_WAIT_NS 4000
#endif
#ifndef AGRS_TMA_WAIT_NS
#define AGRS_TMA_WAIT_NS 500
#endif
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity, uint32_t ns) {
  if (ns) mbar_wait_relaxed(bar, parity, ns);
  else    mbar_wait(bar, parity);
}
template <int CROSS>
__device__ __forceinline__ constexpr bool raw_early() { return AGRS_RAW_EARLY_RELEASE && AGRS_SPLIT_EARLY_CONVERT && CROSS >= 3; }
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
#if AGRS_SPLIT_EARLY_CONVERT
// The row conversion cut in two (default; -DAGRS_SPLIT_EARLY_CONVERT=0 builds the one-piece loop): the splitter reads and converts
// a row as soon as its raw tile has landed and waits for the destination stage only before the stores -- the chain "stage free ->
// split written" then holds the stores alone.  Measured +3 % in the fp16 mode, +0..1 % with bf16 cross terms
// (profiles/r2_gemm_pipeline_variants.txt).
template <bool MN_MAJOR, int CROSS>
__device__ __forceinline__ void convert_row_16(uint32_t raw, uint32_t row, float mult, uint32_t (&lo)[16], uint32_t (&x)[16]) {
  float v[BK];
  if (MN_MAJOR) {
    const uint32_t base = raw + (row >> 5) * 4096u + (row & 7u) * 4u, c = (row & 31u) >> 3;
#pragma unroll
    for (uint32_t k = 0; k < uint32_t(BK); ++k)
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k]) : "r"(base + k * 128u + ((c ^ (k & 3u)) << 5)));
  } else {
    const uint32_t base = raw + row * 128u;
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c)
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3])
                   : "r"(base + ((c ^ (row & 7u)) << 4)));
  }
#if AGRS_SPLIT_F32X2
  const uint64_t mult2 = pack_f32x2(mult, mult);
#endif
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float e0 = v[2 * p], e1 = v[2 * p + 1];
#if AGRS_SPLIT_F32X2
    if (CROSS == 4) {  // the same two roundings per element as the scalar form below, two elements per FMUL2 / FADD2
      const uint64_t s2 = mul_f32x2(pack_f32x2(e0, e1), mult2);
      float s0, s1, h0, h1, l0, l1;
      unpack_f32x2(s2, s0, s1);
      x[p] = pack_f16(s0, s1);
      unpack_f16(x[p], h0, h1);
      unpack_f32x2(sub_f32x2(s2, pack_f32x2(h0, h1)), l0, l1);
      lo[p] = pack_f16(l0, l1);
      continue;
    }
    if (CROSS == 3) {
      float l0, l1;
      x[p] = pack_bf16(e0, e1);
      unpack_f32x2(sub_f32x2(pack_f32x2(e0, e1), pack_f32x2(__uint_as_float(x[p] << 16), __uint_as_float(x[p] & 0xFFFF0000u))), l0, l1);
      lo[p] = pack_bf16(l0, l1);
      continue;
    }
#endif
    if (CROSS == 4) {
      const float s0 = e0 * mult, s1 = e1 * mult;
      float h0, h1;
      x[p] = pack_f16(s0, s1);
      unpack_f16(x[p], h0, h1);
      lo[p] = pack_f16(s0 - h0, s1 - h1);
    } else {
      x[p] = pack_bf16(e0, e1);
      lo[p] = CROSS == 3 ? pack_bf16(e0 - __uint_as_float(x[p] << 16), e1 - __uint_as_float(x[p] & 0xFFFF0000u)) : pack_bf16(lo_of(e0), lo_of(e1));
    }
  }
}
template <int CROSS>
__device__ __forceinline__ void store_row_16(uint32_t out_lo, uint32_t out_x, uint32_t row, const uint32_t (&lo)[16], const uint32_t (&x)[16]) {
#pragma unroll
  for (uint32_t j = 0; j < 4; ++j) {
    const uint32_t off = bf_chunk_offset<CROSS>(row, j);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(out_lo + off), "r"(lo[4 * j]), "r"(lo[4 * j + 1]), "r"(lo[4 * j + 2]), "r"(lo[4 * j + 3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(out_x + off), "r"(x[4 * j]), "r"(x[4 * j + 1]), "r"(x[4 * j + 2]), "r"(x[4 * j + 3]) : "memory");
  }
}
#endif

// ---- CROSS 4: "fp16x3" with operand scaling ---------------------------------------------------------------------------------
// fp16 has tf32's 11 significant bits at twice the MMA rate but only 5 exponent bits.  The range comes from an exact power-of-two
// scale per row of op(A) and per column of op(B): the row's largest magnitude is brought into [2^13, 2^14), after which
// x' = h + l with h = fp16(x'), l = fp16(x' - h) carries 22 bits, and a*b ~ l_a*h_b + h_a*l_b + h_a*h_b is three fp16 MMAs
// (6 slots per k-block, like bf16x3) with a representation error of 7e-8 of max|C| (tools/split_numerics.py).  The epilogue
// undoes the two scales (exact).  Two small kernels find the magnitudes first: k_absmax_rows / k_absmax_cols write, per row or
// column, the largest |x| as its bit pattern (unsigned compare == float compare for magnitudes; NaN sorts above inf and so
// propagates).  Tiles: h in the x16 slot, l in the lo slot, layout of CROSS 1.
__device__ __forceinline__ void scale_of_max(uint32_t max_bits, float& mult, float& inv) {
  uint32_t e = max_bits >> 23;            // biased exponent of the largest magnitude
  e = e < 14u ? 14u : (e > 254u ? 254u : e);
  mult = __uint_as_float((267u - e) << 23);  // 2^(140 - e): max * mult in [2^13, 2^14)
  inv = __uint_as_float((e - 13u) << 23);    // 2^(e - 140)
}
constexpr uint32_t kIdescF16 = (1u << 4) | (uint32_t(BN >> 3) << 17) | (uint32_t((2 * BM) >> 4) << 24);  // D f32, A = B = fp16, K-major
// epilogue side: undo the scales on the 32 accumulators a thread holds (one row, columns col0 .. col0+31)
__device__ __forceinline__ void unscale_piece(uint32_t (&v)[32], float inv_a, const uint32_t* __restrict__ b_max, int64_t col0, int64_t N, int mx_scalar) {
  if (mx_scalar) {  // one magnitude for the whole of op(B): the same two exact power-of-two factors for every element
    float mu, ib;
    scale_of_max(__ldg(b_max), mu, ib);
    // both factors are powers of two: when their product is a normal number one multiplication applies both (exactly what the two
    // steps give, minus an intermediate underflow / overflow they could hit), and FMUL2 does two accumulators at a time
    const float f = inv_a * ib;
    if ((__float_as_uint(f) & 0x7F800000u) != 0u && (__float_as_uint(f) & 0x7F800000u) != 0x7F800000u) {
      const uint64_t f2 = pack_f32x2(f, f);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a, b;
        unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), f2), a, b);
        v[2 * j] = __float_as_uint(a);
        v[2 * j + 1] = __float_as_uint(b);
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * inv_a * ib);
    return;
  }
  if (col0 + 32 <= N && ((reinterpret_cast<uintptr_t>(b_max + col0) & 15u) == 0)) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(b_max + col0) + j);
      float mu, i0, i1, i2, i3;
      scale_of_max(m.x, mu, i0);
      scale_of_max(m.y, mu, i1);
      scale_of_max(m.z, mu, i2);
      scale_of_max(m.w, mu, i3);
      v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) * inv_a * i0);
      v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) * inv_a * i1);
      v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) * inv_a * i2);
      v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) * inv_a * i3);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float mu, ib;
      scale_of_max(col0 + j < N ? __ldg(b_max + col0 + j) : 0u, mu, ib);
      v[j] = __float_as_uint(__uint_as_float(v[j]) * inv_a * ib);
    }
  }
}

// out[r] = max_c |X[r*ld + c]| as bits; one block per row
__global__ void __launch_bounds__(256) k_absmax_rows(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, uint32_t* __restrict__ out) {
  __shared__ uint32_t part[8];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* row = X + r * ld;
    uint32_t m = 0;
    if ((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0) {
      for (int64_t c = threadIdx.x * 4; c + 3 < cols; c += 1024) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + c));
        m = max(max(m, v.x & 0x7FFFFFFFu), max(max(v.y & 0x7FFFFFFFu, v.z & 0x7FFFFFFFu), v.w & 0x7FFFFFFFu));
      }
      for (int64_t c = (cols & ~int64_t(3)) + threadIdx.x; c < cols; c += 256) m = max(m, __float_as_uint(__ldg(row + c)) & 0x7FFFFFFFu);
    } else {
      for (int64_t c = threadIdx.x; c < cols; c += 256) m = max(m, __float_as_uint(__ldg(row + c)) & 0x7FFFFFFFu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 0; w < 8; ++w) m = max(m, part[w]);
      out[r] = m;
    }
    __syncthreads();
  }
}
// out[c] = max_r |X[r*ld + c]| as bits (out zeroed beforehand); blockIdx.x: 256-column strip, blockIdx.y: slab of rows
__global__ void __launch_bounds__(256) k_absmax_cols(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, int64_t slab,
                                                     uint32_t* __restrict__ out) {
  const int64_t c = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (c >= cols) return;
  const int64_t r0 = int64_t(blockIdx.y) * slab, r1 = r0 + slab < rows ? r0 + slab : rows;
  uint32_t m = 0;
  for (int64_t r = r0; r < r1; ++r) m = max(m, __float_as_uint(__ldg(X + r * ld + c)) & 0x7FFFFFFFu);
  atomicMax(out + c, m);  // a maximum does not depend on the order: deterministic
}

// GEMM -> reduce-scatter (data-parallel weight gradients, csrc/fused_dp.cu): with SCATTER the epilogue also pushes every finished
// 32-column piece of C to the rank that owns it -- element e = off + row*N + col of the flat parameter space lies in granule
// e >> gshift, which belongs to rank granule % world; every owner keeps one slot per source rank -- with P2P stores over NVLink,
// tile by tile, while the tensor cores are already on the next tile.
struct ScatterArgs {
  float* base[16];  // arena of every rank as mapped here
  int world, rank;
  size_t g_off;     // float offset of the gradient region in an arena
  size_t shard;     // floats per owner (a multiple of the granule)
  size_t off;       // float offset of C inside the flat parameter space (a multiple of 32)
  int gshift;       // log2 of the ownership granule in floats (>= 5)
};

// ACT (forward activation next to the pre-activation, agrs_gemm_bias_act_f32; staged epilogue, unsplit tiles only): the epilogue
// is the plain one -- K chunks fold into Z by TMA store / reduce-add -- plus one signal per tile: "my pieces of Z have been
// performed".  Warps 2 and 3, idle in the plain kernel, are the ACTIVATION WARPS: once the four epilogue warps of their CTA have
// signalled a tile they read its 128 x 256 part of Z back from L2 (written microseconds ago; coalesced 16-byte loads, 16 in flight
// per lane), apply the activation and store Y, while tensor cores, splitters and epilogue are on the next tile.  Y is thus computed
// from exactly the bits a separate activation kernel would read, Z is never re-read from HBM, and nothing waits for it: the
// epilogue warps are the tightest loop of the kernel after the splitters (two earlier versions did the work there -- adding the
// fetched partial sum to the last chunk in registers, then a post-pass after releasing the accumulator -- and cost the forward
// GEMMs of C2 / C5 8-14 %).  The largest finite |act(Z)| goes to y_max (the magnitude word of Y: the next layer's GEMM scales its
// operand by it).
struct ActArgs {
  int kind;         // AGRS_ACT_RELU / AGRS_ACT_SIGMOID
  float* y;         // [M, N], row pitch ldy (a multiple of 4 floats, 16-byte aligned base)
  int64_t ldy;
  uint32_t* y_max;  // may be null
};
template <bool A_MN, bool B_MN, bool SCATTER, int CROSS, bool ACT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_gemm_3xtf32_2cta(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* C, int64_t ldc,
                   const float* __restrict__ bias, int64_t M, int64_t N, int64_t K, int kb_per_chunk, int splits, float* Cw,
                   const __grid_constant__ ScatterArgs sc, const __grid_constant__ CUtensorMap tmap_c,
                   const __grid_constant__ CUtensorMap tmap_w, int epi_tma, const uint32_t* __restrict__ a_max,
                   const uint32_t* __restrict__ b_max, int mx_scalar, unsigned int* wave_sync, int64_t split_from,
                   const ActArgs act, uint32_t l2_hint) {
  static_assert(!(ACT && SCATTER), "the activation epilogue is the staged one; the scattering epilogue stores from registers");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t lo_base = smem_base + R * RAW_BYTES;
  const uint32_t epi_base = lo_base + L * RAW_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto raw_full = [&](int s) { return bar_base + 8u * s; };
  auto raw_free = [&](int s) { return bar_base + 8u * (R + s); };
  auto lo_free = [&](int s) { return bar_base + 8u * (2 * R + s); };
  auto split_done = [&](int s) { return bar_base + 8u * (2 * R + L + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * R + 2 * L + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * R + 2 * L + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * R + 2 * L + 4);
  const uint32_t z_done_bar = bar_base + 8u * (2 * R + 2 * L + 5);    // ACT: the four epilogue warps have completed a tile of Z
  const uint32_t act_done_bar = bar_base + 8u * (2 * R + 2 * L + 6);  // ACT: the two activation warps have finished with that tile
  static_assert(8u * (2 * R + 2 * L + 7) <= 256u, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int64_t pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int64_t m_tiles = (M + 2 * BM - 1) / (2 * BM), n_tiles = (N + BN - 1) / BN;
  const int num_kb = int((K + BK - 1) / BK);
  // Work items: tiles [0, split_from) whole, tiles [split_from, tiles_mn) cut into `splits` pieces of K.  Splitting K fills the
  // last wave when the tile count is just above a multiple of the 74 CTA pairs (dW of the 4096-wide layers: 256 tiles = 3.46
  // waves: the 34 tiles of the last wave are halved): piece s > 0 writes its partial to plane s-1 of the workspace Cw
  // ([splits-1][M][N]) and a small kernel adds the partials of the split tiles to C in split order afterwards (deterministic).
  // split_from = 0 splits every tile (a forced AGRS_TUNE_TC_SPLITK, or few tiles).
  const int64_t tiles_mn = m_tiles * n_tiles;
  const int64_t tail_tiles = tiles_mn - split_from;
  const int64_t num_tiles = split_from + tail_tiles * splits;
  const int kb_per_split = (num_kb + splits - 1) / splits;
  struct Work {
    int64_t mb, nb;
    int kb0, kb1, split;
  };
  auto work_item = [&](int64_t t) {
    Work w;
    if (t < split_from) {
      w.split = 0;
      tile_coords(t, m_tiles, n_tiles, w.mb, w.nb);
      w.kb0 = 0;
      w.kb1 = num_kb;
      return w;
    }
    const int64_t u = t - split_from;
    w.split = int(u / tail_tiles);
    tile_coords(split_from + (u - w.split * tail_tiles), m_tiles, n_tiles, w.mb, w.nb);
    w.kb0 = w.split * kb_per_split;
    w.kb1 = w.kb0 + kb_per_split < num_kb ? w.kb0 + kb_per_split : num_kb;
    return w;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(raw_free(s), raw_early<CROSS>() ? kSplitWarps / kSplitGroups : 1);  // MMA completion, or this CTA's splitter group
    }
    for (int s = 0; s < L; ++s) {
      mbar_init(lo_free(s), 1);
      mbar_init(split_done(s), 2 * kSplitWarps / kSplitGroups);  // the splitter warps of one group, in both CTAs, arrive on the leader's barrier
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * 4);  // both CTAs' epilogue warps arrive on the leader's barrier
    }
    if constexpr (ACT) {
      mbar_init(z_done_bar, 4);
      mbar_init(act_done_bar, kActWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything is signalled across CTAs
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): my 128 rows of A, my 128 columns of B =====================
    if (lane == 0) {
      int r = 0;
      uint32_t rph = 0;
      unsigned long long started = 0;  // work items started by all pairs once every pair has begun its (wave+1)-th item
      // L2 eviction policies of the operand loads (launch(): AGRS_TC_L2_HINT): bit 0 = B evict_last, bit 1 = A evict_first
      const uint64_t pol_b = l2_policy_evict_last(), pol_a = l2_policy_evict_first();
      const bool hint_b = (l2_hint & 1u) != 0, hint_a = (l2_hint & 2u) != 0;
      for (int64_t t = pair; t < num_tiles; t += num_pairs) {
        const Work w = work_item(t);
        const int32_t m0 = int32_t(w.mb * 2 * BM + rank * BM), n0 = int32_t(w.nb * BN + rank * BNH);
        // Wave alignment (large problems only, see launch()): the pairs of a wave share A row-blocks and B column-blocks through
        // the L2, but only while they walk K within a few dozen k-blocks of each other; left alone they drift apart tile after tile
        // (16384^3: 64 GB of DRAM reads for 2 GB of operands).  The leader's producer therefore announces each work item and waits,
        // for a bounded time, until every pair has announced its item of the same wave.  A performance hint, not a dependency:
        // when the wait times out (pairs not co-resident) the loads simply start.
        started = started + num_pairs < (unsigned long long)num_tiles ? started + num_pairs : (unsigned long long)num_tiles;
        if (wave_sync && rank == 0) {
          atomicAdd(wave_sync, 1u);
          for (int polls = 0; polls < 2048; ++polls) {
            unsigned int seen;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(wave_sync) : "memory");
            if (seen >= (unsigned int)started) break;
            __nanosleep(64);
          }
        }
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait_long(raw_free(r), rph ^ 1u, AGRS_TMA_WAIT_NS);
          const uint32_t a_dst = smem_base + r * RAW_BYTES, b_dst = a_dst + A_BYTES;
          mbar_arrive_expect_tx(raw_full(r), RAW_BYTES);
          const int32_t k0 = kb * BK;
          if (A_MN) {  // A stored [K, M]: boxes of 32 (M, contiguous) x 32 (K rows)
#pragma unroll
            for (int g = 0; g < BM / 32; ++g) {
              if (hint_a) tma_load_2d_hint(&tmap_a, raw_full(r), a_dst + g * 4096, m0 + g * 32, k0, pol_a);
              else        tma_load_2d(&tmap_a, raw_full(r), a_dst + g * 4096, m0 + g * 32, k0);
            }
          } else {     // A stored [M, K]: one box of 32 (K, contiguous) x 128 rows
            if (hint_a) tma_load_2d_hint(&tmap_a, raw_full(r), a_dst, k0, m0, pol_a);
            else        tma_load_2d(&tmap_a, raw_full(r), a_dst, k0, m0);
          }
          if (B_MN) {  // B stored [K, N]
#pragma unroll
            for (int g = 0; g < BNH / 32; ++g) {
              if (hint_b) tma_load_2d_hint(&tmap_b, raw_full(r), b_dst + g * 4096, n0 + g * 32, k0, pol_b);
              else        tma_load_2d(&tmap_b, raw_full(r), b_dst + g * 4096, n0 + g * 32, k0);
            }
          } else {     // B stored [N, K]
            if (hint_b) tma_load_2d_hint(&tmap_b, raw_full(r), b_dst, k0, n0, pol_b);
            else        tma_load_2d(&tmap_b, raw_full(r), b_dst, k0, n0);
          }
          if (++r == R) { r = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc<A_MN, B_MN, BN, 2 * BM>();
      int r = 0, l = 0, acc = 0;
      uint32_t lph = 0, acc_ph = 0;
      for (int64_t t = pair; t < num_tiles; t += num_pairs) {
        const Work w = work_item(t);
        for (int c0 = w.kb0; c0 < w.kb1; c0 += kb_per_chunk) {
          mbar_wait_cluster(tempty_bar(acc), acc_ph ^ 1u);  // both epilogues have drained this accumulator stage
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
          const int kb_end = c0 + kb_per_chunk < w.kb1 ? c0 + kb_per_chunk : w.kb1;
          for (int kb = c0; kb < kb_end; ++kb) {
            mbar_wait_cluster(split_done(l), lph);  // both CTAs: raw tile landed and lo tile written
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_hi = smem_base + r * RAW_BYTES, b_hi = a_hi + A_BYTES;
              const bool first_kb = (kb == c0);
              if constexpr (CROSS == 0) {
                const uint32_t a_lo = lo_base + l * RAW_BYTES, b_lo = a_lo + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  const uint64_t da_hi = make_smem_desc<A_MN>(a_hi + k * k_slice_bytes<A_MN>());
                  const uint64_t da_lo = make_smem_desc<A_MN>(a_lo + k * k_slice_bytes<A_MN>());
                  const uint64_t db_hi = make_smem_desc<B_MN>(b_hi + k * k_slice_bytes<B_MN>());
                  const uint64_t db_lo = make_smem_desc<B_MN>(b_lo + k * k_slice_bytes<B_MN>());
                  umma_tf32_2cta(d_tmem, da_lo, db_hi, idesc, (first_kb && k == 0) ? 0u : 1u);  // small terms first
                  umma_tf32_2cta(d_tmem, da_hi, db_lo, idesc, 1u);
                  umma_tf32_2cta(d_tmem, da_hi, db_hi, idesc, 1u);
                }
              } else {
                // lo stage: [A lo | A x16 | B lo | B x16], 8 KB each, K-major bf16
                const uint32_t a_lo = lo_base + l * RAW_BYTES, a_x = a_lo + BF_TILE, b_lo = a_x + BF_TILE, b_x = b_lo + BF_TILE;
                constexpr uint32_t idesc16 = CROSS == 4 ? kIdescF16 : kIdescBf16;
#pragma unroll
                for (int i = 0; i < BK / 16; ++i) {  // cross terms first (they are the small ones), two 16-wide slices each
                  umma_bf16_2cta(d_tmem, make_bf_desc<CROSS>(a_lo, i), make_bf_desc<CROSS>(b_x, i), idesc16, (first_kb && i == 0) ? 0u : 1u);
                  umma_bf16_2cta(d_tmem, make_bf_desc<CROSS>(a_x, i), make_bf_desc<CROSS>(b_lo, i), idesc16, 1u);
                  if constexpr (CROSS >= 3)  // bf16x3 / fp16x3: the leading term is 16-bit too (x16 tiles hold the first piece)
                    umma_bf16_2cta(d_tmem, make_bf_desc<CROSS>(a_x, i), make_bf_desc<CROSS>(b_x, i), idesc16, 1u);
                }
                if constexpr (CROSS < 3) {
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_tf32_2cta(d_tmem, make_smem_desc<A_MN>(a_hi + k * k_slice_bytes<A_MN>()),
                                   make_smem_desc<B_MN>(b_hi + k * k_slice_bytes<B_MN>()), idesc, 1u);
                }
              }
              if constexpr (!raw_early<CROSS>()) umma_commit_pair(raw_free(r));  // both CTAs may refill this raw stage ...
              umma_commit_pair(lo_free(l));   // ... and this lo stage once the MMAs above have read them
              if (kb == kb_end - 1) umma_commit_pair(tfull_bar(acc));
            }
            __syncwarp();
            if (++r == R) r = 0;
            if (++l == L) { l = 0; lph ^= 1u; }
          }
          if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
      }
    }
  } else if (warp >= kSplitWarp0) {
    // ===================== splitters (both CTAs): lo = x - trunc_tf32(x), same swizzled offsets as the raw tile =====================
    // The 8 warps form kSplitGroups groups that take k-blocks in turn (group g: k-blocks g, g + kSplitGroups, ..): every warp
    // then has kSplitGroups k-block times for its wait -> load -> split -> store -> fence -> arrive chain, whose fixed latencies
    // (barrier polls, LDS round trip, proxy fence) would otherwise cap the pipeline with all 8 warps serialised on each k-block.
    constexpr int kGroupWarps = kSplitWarps / kSplitGroups;
    const int grp = (warp - kSplitWarp0) / kGroupWarps;
    const int tid = threadIdx.x - (kSplitWarp0 + grp * kGroupWarps) * 32;  // 0 .. 32*kGroupWarps-1 inside the group
    const uint32_t done_remote = mapa_rank(split_done(0), 0);
    constexpr int kPerThread = RAW_BYTES / (kGroupWarps * 32 * 16);  // float4 per thread per stage
    constexpr uint32_t kStride = kGroupWarps * 32 * 16;
    uint32_t it = 0;  // k-blocks since kernel start: ring positions are it % R and it % L
    constexpr int kItems = CROSS == 0 ? 1 : (2 * BM) / (kGroupWarps * 32);  // operand rows per splitter thread and k-block
    for (int64_t t = pair; t < num_tiles; t += num_pairs) {
      const Work w = work_item(t);
      float mult[kItems];
#pragma unroll
      for (int i = 0; i < kItems; ++i) mult[i] = 1.f;
      if constexpr (CROSS == 4) {  // the scales of this thread's rows of A and of B for this tile
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
          const int item = tid + i * kGroupWarps * 32;
          const int64_t g = item < BM ? w.mb * 2 * BM + rank * BM + item : w.nb * BN + rank * BNH + (item - BM);
          // mx_scalar: one magnitude per operand (kept by the host engine with the tensor), otherwise one per row / column
          const uint32_t bits = mx_scalar ? __ldg(item < BM ? a_max : b_max)
                                          : (item < BM ? (g < M ? __ldg(a_max + g) : 0u) : (g < N ? __ldg(b_max + g) : 0u));
          float inv;
          scale_of_max(bits, mult[i], inv);
        }
      }
      for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
        if (int(it % kSplitGroups) != grp) continue;
        const uint32_t r = it % R, rph = (it / R) & 1u, l = it % L, lph = (it / L) & 1u;
        mbar_wait(raw_full(r), rph);
#if AGRS_SPLIT_EARLY_CONVERT
        if constexpr (CROSS != 0) {  // variant build: convert from the raw tile first, wait for the destination stage afterwards
          const uint32_t raw = smem_base + r * RAW_BYTES, out = lo_base + l * RAW_BYTES;
          uint32_t lo16[kItems][16], x16[kItems][16];
#pragma unroll
          for (int i = 0; i < kItems; ++i) {
            const int item = tid + i * kGroupWarps * 32;
            const uint32_t row = uint32_t(item) & uint32_t(BM - 1);
            if (item < BM) convert_row_16<A_MN, CROSS>(raw, row, mult[i], lo16[i], x16[i]);
            else           convert_row_16<B_MN, CROSS>(raw + A_BYTES, row, mult[i], lo16[i], x16[i]);
          }
          if constexpr (raw_early<CROSS>()) {  // every lane's loads of the raw tile are complete (their values were just used)
            __syncwarp();
            if (lane == 0) mbar_arrive(raw_free(r));
          }
          mbar_wait(lo_free(l), lph ^ 1u);
#pragma unroll
          for (int i = 0; i < kItems; ++i) {
            const int item = tid + i * kGroupWarps * 32;
            const uint32_t row = uint32_t(item) & uint32_t(BM - 1);
            if (item < BM) store_row_16<CROSS>(out, out + BF_TILE, row, lo16[i], x16[i]);
            else           store_row_16<CROSS>(out + 2 * BF_TILE, out + 3 * BF_TILE, row, lo16[i], x16[i]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(done_remote + 8u * l);
          continue;
        }
#endif
        mbar_wait(lo_free(l), lph ^ 1u);
        if constexpr (CROSS == 0) {
          const uint32_t src = smem_base + r * RAW_BYTES + tid * 16, dst = lo_base + l * RAW_BYTES + tid * 16;
          float4 v[kPerThread];
#pragma unroll
          for (int i = 0; i < kPerThread; ++i)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                         : "r"(src + i * kStride));
#pragma unroll
          for (int i = 0; i < kPerThread; ++i)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst + i * kStride), "f"(lo_of(v[i].x)), "f"(lo_of(v[i].y)),
                         "f"(lo_of(v[i].z)), "f"(lo_of(v[i].w))
                         : "memory");
        } else {
          // the group's threads share the 128 rows of A and the 128 rows of B of this k-block (a warp always holds 32 consecutive rows)
          static_assert((2 * BM) % (kGroupWarps * 32) == 0 && kGroupWarps * 32 <= 2 * BM, "bf16 cross terms: whole rows per splitter thread");
          const uint32_t raw = smem_base + r * RAW_BYTES, out = lo_base + l * RAW_BYTES;
#pragma unroll
          for (int i = 0; i < kItems; ++i) {
            const int item = tid + i * kGroupWarps * 32;
            const uint32_t row = uint32_t(item) & uint32_t(BM - 1);
            if (item < BM) split_row_bf16<A_MN, CROSS>(raw, out, out + BF_TILE, row, mult[i]);
            else           split_row_bf16<B_MN, CROSS>(raw + A_BYTES, out + 2 * BF_TILE, out + 3 * BF_TILE, row, mult[i]);
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(done_remote + 8u * l);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (both CTAs): my 128 rows of the pair tile =====================
    const int q = warp - kEpiWarp0;  // == warp % 4: TMEM lanes [32q, 32q+32)
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
    const bool vec_w = ((N & 3) == 0) && (((M * N) & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cw) & 15u) == 0);
    const uint32_t tempty_remote = mapa_rank(tempty_bar(0), 0);
    int acc = 0;
    uint32_t acc_ph = 0;
    if (!SCATTER && epi_tma) {
      // ---- staged epilogue: TMEM -> registers -> swizzled shared memory -> TMA store / reduce-add (see tma_store_3d) ----
      const uint32_t stage0 = epi_base + uint32_t(q) * (2 * EPI_TILE);
      uint32_t issued = 0;  // pieces this warp has handed to the TMA (buffer = issued & 1)
      // one 32 x 32 piece: registers -> swizzled shared memory -> TMA tensor store (or reduce-add)
      auto hand_over = [&](const uint32_t (&v)[32], const CUtensorMap* map, bool reduce, int32_t col0, int32_t row0, int32_t plane) {
        const uint32_t buf = stage0 + (issued & 1u) * EPI_TILE;
        if (lane == 0) bulk_wait_read<1>();  // the piece staged in this buffer two pieces ago has been read out
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)  // row `lane` of the piece, 16-byte chunk j in slot j ^ (lane & 7): SWIZZLE_128B
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + uint32_t(lane) * 128u + (uint32_t(j ^ (lane & 7)) << 4)),
                       "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (reduce) tma_reduce_add_3d(map, buf, col0, row0, plane);
          else        tma_store_3d(map, buf, col0, row0, plane);
          bulk_commit();
        }
        ++issued;
      };
      uint32_t tiles_done = 0;  // ACT: tiles this warp has finished (phase bookkeeping of the hand-over to the activation warps)
      for (int64_t t = pair; t < num_tiles; t += num_pairs) {
        const Work w = work_item(t);
        const int64_t row0 = w.mb * 2 * BM + rank * BM + q * 32;
        // split 0 produces C (with the bias); split s > 0 plane s-1 of the workspace [splits-1][M][N]
        const CUtensorMap* const cmap = w.split == 0 ? &tmap_c : &tmap_w;
        const int32_t plane = w.split == 0 ? 0 : w.split - 1;
        const float* const bias_s = w.split == 0 ? bias : nullptr;
        float inv_a = 1.f;
        if constexpr (CROSS == 4) {
          float mu;
          scale_of_max(mx_scalar ? __ldg(a_max) : (row0 + lane < M ? __ldg(a_max + row0 + lane) : 0u), mu, inv_a);
        }
        for (int c0 = w.kb0, ch = 0; c0 < w.kb1; c0 += kb_per_chunk, ++ch) {
          mbar_wait_long(tfull_bar(acc), acc_ph, AGRS_EPI_WAIT_NS);
          tc_fence_after();
          // the reductions of this chunk go to the addresses the previous chunk's pieces went to: those must have landed
          if (ch > 0 && lane == 0) bulk_wait_all();
          const bool final_chunk = ACT && c0 + kb_per_chunk >= w.kb1;  // ACT launches never split K over pairs: Z is final here
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN);
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_wait_ld();
            const int64_t col0 = w.nb * BN + c * 32;
            if (row0 < M && col0 < N) {  // warp-uniform; rows and columns past the edge are clipped by the tensor map
              if constexpr (CROSS == 4) unscale_piece(v, inv_a, b_max, col0, N, mx_scalar);
              if (ch == 0 && bias_s) {
                if (col0 + 32 <= N) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias_s + col0) + j);
                    v[4 * j] = __float_as_uint(b.x + __uint_as_float(v[4 * j]));
                    v[4 * j + 1] = __float_as_uint(b.y + __uint_as_float(v[4 * j + 1]));
                    v[4 * j + 2] = __float_as_uint(b.z + __uint_as_float(v[4 * j + 2]));
                    v[4 * j + 3] = __float_as_uint(b.w + __uint_as_float(v[4 * j + 3]));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (col0 + j < N) v[j] = __float_as_uint(__ldg(bias_s + col0 + j) + __uint_as_float(v[j]));
                }
              }
              hand_over(v, cmap, ch != 0, int32_t(col0), int32_t(row0), plane);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(tempty_remote + 8u * acc);
          if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
          if constexpr (ACT) {
            if (final_chunk) {
              // the tile's Z is final once this chunk's pieces have been performed: tell the activation warps (see below).  They
              // are at most one tile behind: a tile is only announced after they have finished with the one before it.
              if (lane == 0) {
                bulk_wait_all();
                if (tiles_done > 0) mbar_wait(act_done_bar, (tiles_done - 1u) & 1u);
                mbar_arrive(z_done_bar);
              }
              ++tiles_done;
            }
          }
        }
      }
      if (lane == 0) bulk_wait_all();  // shared memory stays valid, and C is complete, before the CTA leaves
    } else
    for (int64_t t = pair; t < num_tiles; t += num_pairs) {
      const Work w = work_item(t);
      const int64_t nb = w.nb;
      const int64_t row = w.mb * 2 * BM + rank * BM + q * 32 + lane;
      // split 0 produces C (with the bias); split s > 0 its own partial [M][N] in the workspace
      float* const Cs = w.split == 0 ? C : Cw + int64_t(w.split - 1) * M * N;
      const int64_t lds = w.split == 0 ? ldc : N;
      const float* const bias_s = w.split == 0 ? bias : nullptr;
      const bool vec_s = w.split == 0 ? vec_ok : vec_w;
      float inv_a = 1.f;
      if constexpr (CROSS == 4) {
        float mu;
        scale_of_max(mx_scalar ? __ldg(a_max) : (row < M ? __ldg(a_max + row) : 0u), mu, inv_a);
      }
      for (int c0 = w.kb0, ch = 0; c0 < w.kb1; c0 += kb_per_chunk, ++ch) {
        mbar_wait_long(tfull_bar(acc), acc_ph, AGRS_EPI_WAIT_NS);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_wait_ld();
          const int64_t col0 = nb * BN + c * 32;
          if (row < M && col0 < N) {
            if constexpr (CROSS == 4) unscale_piece(v, inv_a, b_max, col0, N, mx_scalar);
            float* dst = Cs + row * lds + col0;
            float* push = nullptr;  // where the finished piece also goes: this rank's slot at the piece's owner
            if (SCATTER && c0 + kb_per_chunk >= w.kb1) {
              // ownership is interleaved in granules of 2^gshift floats (64 KB by default: a few rows of C): granule g of the flat
              // parameter space belongs to rank g % world and is the (g / world)-th granule of that rank's slice -- the 32 rows of
              // an epilogue warp spread over all ranks, so no rank's NVLink ingress is everybody's target
              const size_t e = sc.off + size_t(row) * size_t(N) + size_t(col0);
              const size_t g = e >> sc.gshift;
              const size_t owner = g % size_t(sc.world);
              push = sc.base[owner] + sc.g_off + size_t(sc.rank) * sc.shard + ((g / size_t(sc.world)) << sc.gshift) +
                     (e & ((size_t(1) << sc.gshift) - 1));
            }
            if (vec_s && col0 + 32 <= N) {
              float4 base[8];
              if (ch == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  base[j] = bias_s ? __ldg(reinterpret_cast<const float4*>(bias_s + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) base[j] = *reinterpret_cast<const float4*>(dst + 4 * j);  // previous chunks (L2)
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o;
                o.x = base[j].x + __uint_as_float(v[4 * j]);
                o.y = base[j].y + __uint_as_float(v[4 * j + 1]);
                o.z = base[j].z + __uint_as_float(v[4 * j + 2]);
                o.w = base[j].w + __uint_as_float(v[4 * j + 3]);
                *reinterpret_cast<float4*>(dst + 4 * j) = o;
                if (SCATTER && push) *reinterpret_cast<float4*>(push + 4 * j) = o;  // P2P store (a local one when I am the owner)
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) {
                  float b = ch == 0 ? (bias_s ? __ldg(bias_s + col0 + j) : 0.f) : dst[j];
                  dst[j] = b + __uint_as_float(v[j]);
                }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_remote + 8u * acc);
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
    }
  } else if (ACT && warp >= kActWarp0) {
    // ===================== activation warps (ACT, both CTAs): my CTA's 128 x 256 part of every finished tile =====================
    constexpr int kRowsPerWarp = BM / kActWarps;
    constexpr int kBatch = 16;  // 16-byte loads in flight per lane
    const int aw = warp - kActWarp0;
    uint32_t ph = 0, y_mag = 0;
    for (int64_t t = pair; t < num_tiles; t += num_pairs) {
      const Work w = work_item(t);
      mbar_wait_long(z_done_bar, ph, AGRS_EPI_WAIT_NS);  // the bulk operations that wrote Z have completed: visible to plain loads
      ph ^= 1u;
      const int64_t r0 = w.mb * 2 * BM + rank * BM + aw * kRowsPerWarp, c0 = w.nb * BN;
      const int rows = r0 < M ? int(M - r0 < kRowsPerWarp ? M - r0 : kRowsPerWarp) : 0;
      const int vpr = int((N - c0 < BN ? N - c0 : BN) >> 2);  // 16-byte vectors per row (N % 4 == 0 on this path)
      const int total = rows * vpr;
      const float* const z_tile = C + r0 * ldc + c0;
      float* const y_tile = act.y + r0 * act.ldy + c0;
      auto activate = [&](const float4 x) {
        float4 y;
        if (act.kind == AGRS_ACT_RELU) y = make_float4(agrs_act_relu(x.x), agrs_act_relu(x.y), agrs_act_relu(x.z), agrs_act_relu(x.w));
        else y = make_float4(agrs_act_sigmoid(x.x), agrs_act_sigmoid(x.y), agrs_act_sigmoid(x.z), agrs_act_sigmoid(x.w));
        y_mag = mag4(y_mag, y);
        return y;
      };
      if (vpr == BN / 4) {
        // a whole-width tile (all but the last column block): a batch is 8 rows, a lane owns vectors `lane` and `lane + 32` of each
        // row -- every address is the batch's base plus a constant, no index arithmetic per vector (these two warps share their
        // sub-partitions' issue slots with four splitter warps: a first version with a division per vector slowed THOSE down)
#pragma unroll 1
        for (int rb = 0; rb < rows; rb += kBatch / 2) {
          const float* zr = z_tile + int64_t(rb) * ldc + lane * 4;
          float* yr = y_tile + int64_t(rb) * act.ldy + lane * 4;
          float4 x[kBatch];
#pragma unroll
          for (int i = 0; i < kBatch; i += 2, zr += ldc)
            if (rb + (i >> 1) < rows) {  // from L2, never through this SM's L1
              x[i] = __ldcg(reinterpret_cast<const float4*>(zr));
              x[i + 1] = __ldcg(reinterpret_cast<const float4*>(zr + 128));
            }
#pragma unroll
          for (int i = 0; i < kBatch; i += 2, yr += act.ldy)
            if (rb + (i >> 1) < rows) {
              *reinterpret_cast<float4*>(yr) = activate(x[i]);
              *reinterpret_cast<float4*>(yr + 128) = activate(x[i + 1]);
            }
        }
      } else {  // the clipped last column block: generic indexing, four vectors in flight
        constexpr int kEdgeBatch = 4;
#pragma unroll 1
        for (int base = 0; base < total; base += kEdgeBatch * 32) {
          float4 x[kEdgeBatch];
          int r[kEdgeBatch], v[kEdgeBatch];
#pragma unroll
          for (int i = 0; i < kEdgeBatch; ++i) {
            const int e = base + i * 32 + lane;
            r[i] = e / vpr;
            v[i] = e - r[i] * vpr;
            if (e < total) x[i] = __ldcg(reinterpret_cast<const float4*>(z_tile + r[i] * ldc) + v[i]);
          }
#pragma unroll
          for (int i = 0; i < kEdgeBatch; ++i)
            if (base + i * 32 + lane < total) *(reinterpret_cast<float4*>(y_tile + r[i] * act.ldy) + v[i]) = activate(x[i]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(act_done_bar);
    }
    if (act.y_max) mag_commit(y_mag, act.y_max);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still read its shared memory or signal its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

__global__ void k_zero_u32(unsigned int* p) { *p = 0u; }

// C[m, n] += P_1[m, n] + P_2[m, n] + ..  (split order; P_s contiguous [M][N])
__global__ void __launch_bounds__(256) k_add_partials(float* __restrict__ C, int64_t ldc, const float* __restrict__ P, int nparts, int64_t M,
                                                      int64_t N) {
  const int64_t total = M * N, stride = int64_t(gridDim.x) * blockDim.x;
  if ((N & 3) == 0 && (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(P)) & 15u) == 0) {
    for (int64_t i4 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i4 < total / 4; i4 += stride) {
      const int64_t i = i4 * 4, m = i / N, n = i - m * N;
      float4 c = *reinterpret_cast<const float4*>(C + m * ldc + n);
      for (int s = 0; s < nparts; ++s) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(P + int64_t(s) * total + i));
        c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
      }
      *reinterpret_cast<float4*>(C + m * ldc + n) = c;
    }
  } else {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
      const int64_t m = i / N, n = i - m * N;
      float c = C[m * ldc + n];
      for (int s = 0; s < nparts; ++s) c += P[int64_t(s) * total + i];
      C[m * ldc + n] = c;
    }
  }
}

// the same for the tiles [tile_from, tiles_mn) only (tile order of the GEMM: tile_coords): eight blocks per tile, 32 rows each
__global__ void __launch_bounds__(256) k_add_partials_tiles(float* __restrict__ C, int64_t ldc, const float* __restrict__ P, int nparts, int64_t M, int64_t N,
                                                            int64_t m_tiles, int64_t n_tiles, int64_t tile_from, int64_t tiles_mn) {
  const bool vec = (N & 3) == 0 && (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(P)) & 15u) == 0;
  const int64_t total = M * N;
  constexpr int kSlices = 8, kSliceRows = 2 * BM / kSlices;
  for (int64_t u = blockIdx.x; u < (tiles_mn - tile_from) * kSlices; u += gridDim.x) {
    const int64_t t = tile_from + u / kSlices;
    int64_t mb, nb;
    tile_coords(t, m_tiles, n_tiles, mb, nb);
    const int64_t r0 = mb * 2 * BM + (u % kSlices) * kSliceRows, c0 = nb * BN;
    if (r0 >= M) continue;
    const int rows = int(M - r0 < kSliceRows ? M - r0 : kSliceRows), cols = int(N - c0 < BN ? N - c0 : BN);
    if (vec) {  // cols is a multiple of 4 (N and the tile width are)
      const int c4n = cols >> 2;
      for (int e = threadIdx.x; e < rows * c4n; e += 256) {
        const int r = e / c4n, c = (e - r * c4n) << 2;
        float4 acc = *reinterpret_cast<const float4*>(C + (r0 + r) * ldc + c0 + c);
        for (int sp = 0; sp < nparts; ++sp) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(P + int64_t(sp) * total + (r0 + r) * N + c0 + c));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(C + (r0 + r) * ldc + c0 + c) = acc;
      }
    } else {
      for (int e = threadIdx.x; e < rows * cols; e += 256) {
        const int r = e / cols, c = e - r * cols;
        float acc = C[(r0 + r) * ldc + c0 + c];
        for (int sp = 0; sp < nparts; ++sp) acc += P[int64_t(sp) * total + (r0 + r) * N + c0 + c];
        C[(r0 + r) * ldc + c0 + c] = acc;
      }
    }
  }
}

// K splits that fill the last wave of CTA pairs.  Time in units of one whole tile on one pair: no split ceil(T / P); every tile in s
// pieces ceil(T s / P) / s; only the tiles of the last, partial wave in s pieces: floor(T / P) + 1 / s, possible when those tiles
// times s still fit one wave.  The last form adds partial planes for a few tiles only, so it wins ties.
struct SplitPlan {
  int splits;          // pieces of K per split tile (1: nothing is split)
  int64_t split_from;  // first split tile (0: all of them)
};
static SplitPlan choose_plan(int64_t tiles, int64_t pairs, int64_t num_kb, int forced) {
  if (forced > 0) return {int(forced < num_kb ? forced : num_kb), 0};
  const double none = double((tiles + pairs - 1) / pairs);
  SplitPlan best{1, 0};
  double best_t = none;
  for (int sp = 2; sp <= 4; ++sp) {
    if (num_kb / sp < 8) break;  // at least 256 of K per piece
    const double t = double((tiles * sp + pairs - 1) / pairs) / sp;
    if (t < best_t / 1.05) {  // worth a second pass over C only for a real gain
      best = {sp, 0};
      best_t = t;
    }
  }
  const int64_t full = tiles / pairs, rem = tiles - full * pairs;
  if (full >= 1 && rem > 0) {
    int sp = int(pairs / rem);
    if (sp > 4) sp = 4;
    while (sp >= 2 && num_kb / sp < 8) --sp;
    if (sp >= 2) {
      const double t = double(full) + 1.0 / sp;
      if (t < none / 1.05 && t <= best_t * 1.001) best = {sp, full * pairs};
    }
  }
  return best;
}
static int choose_splits(int64_t tiles, int64_t pairs, int64_t num_kb, int forced) { return choose_plan(tiles, pairs, num_kb, forced).splits; }

template <bool A_MN, bool B_MN, bool SCATTER, int CROSS, bool ACT = false>
static int launch(agrs_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, float* C, int64_t ldc, const float* bias, int64_t M,
                  int64_t N, int64_t K, int kb_per_chunk, const ScatterArgs& sc, const float* A, int64_t lda, const float* B, int64_t ldb) {
  auto kern = k_gemm_3xtf32_2cta<A_MN, B_MN, SCATTER, CROSS, ACT>;
  AGRS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
  const int64_t tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int64_t pairs = ctx->sm_count / 2;
  const int64_t num_kb = (K + BK - 1) / BK;
  SplitPlan plan = SCATTER ? SplitPlan{1, 0} : choose_plan(tiles, pairs, num_kb, ctx->tc_splitk);  // the scattering epilogue writes final values only
  int splits = plan.splits;
  {  // the kernel gives every split ceil(num_kb / splits) k-blocks: drop splits that would be left with none
    const int64_t kps = (num_kb + splits - 1) / splits;
    splits = int((num_kb + kps - 1) / kps);
  }
  const int64_t split_from = splits > 1 ? plan.split_from : 0;
  if constexpr (ACT) {
    // the activation warps need Z final when the kernel's own epilogue has finished a tile (no K split over pairs, staged epilogue)
    // and 16-byte vectors on both outputs; otherwise this launch is the plain one and the caller runs the activation as a second pass
    const agrs_ctx::GemmAct& ga = ctx->gemm_act;
    if (!(ctx->tc_epi_tma && (ldc & 3) == 0 && agrs_aligned16(C) && splits == 1 && (N & 3) == 0 && (ga.ldy & 3) == 0 && agrs_aligned16(ga.y)))
      return launch<A_MN, B_MN, SCATTER, CROSS, false>(ctx, ta, tb, C, ldc, bias, M, N, K, kb_per_chunk, sc, A, lda, B, ldb);
  }
  // workspace = [row magnitudes of op(A) | column magnitudes of op(B)] (CROSS 4 only) followed by the split-K partial planes
  const size_t a_slots = (size_t(M) + 63) / 64 * 64, b_slots = (size_t(N) + 63) / 64 * 64;
  // cross mode 4 with magnitudes supplied by the caller (agrs_gemm_operand_absmax: one per operand): no passes, no scale region
  const bool ext_mx = CROSS == 4 && ctx->gemm_a_mx && ctx->gemm_b_mx;
  const size_t scale_bytes = CROSS == 4 && !ext_mx ? (a_slots + b_slots) * sizeof(uint32_t) : 0;
  const size_t part_bytes = splits > 1 ? size_t(splits - 1) * size_t(M) * size_t(N) * sizeof(float) : 0;
  float* Cw = nullptr;
  uint32_t *a_max = nullptr, *b_max = nullptr;
  if (scale_bytes + part_bytes) {
    int rc = agrs_ensure_workspace(ctx, scale_bytes + part_bytes);
    if (rc) return rc;
    Cw = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->workspace) + scale_bytes);
  }
  if (ext_mx) {
    a_max = const_cast<uint32_t*>(ctx->gemm_a_mx);
    b_max = const_cast<uint32_t*>(ctx->gemm_b_mx);
  } else if (CROSS == 4) {
    a_max = reinterpret_cast<uint32_t*>(ctx->workspace);
    b_max = a_max + a_slots;
    // op(A) row m: a row of A stored [M, K], a column of A stored [K, M]; op(B) column n: a column of B stored [K, N], a row of [N, K]
    auto absmax = [&](const float* X, bool along_columns, int64_t vecs, int64_t ld, uint32_t* out) -> int {
      if (along_columns) {  // X is [K rows][vecs columns]
        AGRS_CUDA(ctx, cudaMemsetAsync(out, 0, size_t(vecs) * sizeof(uint32_t), ctx->stream));
        const int64_t slab = 32;
        dim3 grid(unsigned((vecs + 255) / 256), unsigned((K + slab - 1) / slab));
        k_absmax_cols<<<grid, 256, 0, ctx->stream>>>(X, K, vecs, ld, slab, out);
      } else {              // X is [vecs rows][K columns]
        const int64_t cap = int64_t(ctx->sm_count) * 16;
        k_absmax_rows<<<unsigned(vecs < cap ? vecs : cap), 256, 0, ctx->stream>>>(X, vecs, K, ld, out);
      }
      AGRS_LAUNCHED(ctx);
      return AGRS_OK;
    };
    // measurement aid (AGRS_TUNE_TC_REUSE_SCALES): skip the two passes and use what the previous launch left in the workspace
    int rc = ctx->tc_reuse_scales ? AGRS_OK : absmax(A, A_MN, M, lda, a_max);
    if (!rc && !ctx->tc_reuse_scales) rc = absmax(B, B_MN, N, ldb, b_max);
    if (rc) return rc;
  }
  // staged TMA epilogue: C (and the split workspace) must be describable by a tensor map -- 16-byte aligned base and row pitch
  CUtensorMap tc_map = ta, tw_map = ta;  // placeholders when unused
  int epi_tma = 0;
  if (!SCATTER && ctx->tc_epi_tma && (ldc & 3) == 0 && agrs_aligned16(C) && (splits == 1 || (N & 3) == 0)) {
    int rc = encode_3d_out(ctx, &tc_map, C, uint64_t(N), uint64_t(M), 1, uint64_t(ldc));
    if (!rc && splits > 1) rc = encode_3d_out(ctx, &tw_map, Cw, uint64_t(N), uint64_t(M), uint64_t(splits - 1), uint64_t(N));
    if (rc) return rc;
    epi_tma = 1;
  }
  ActArgs act{0, nullptr, 0, nullptr};
  if constexpr (ACT) {
    agrs_ctx::GemmAct& ga = ctx->gemm_act;
    if (int rc = agrs_take_mx(ctx, &act.y_max)) return rc;  // agrs_next_output_absmax: zeroed on the stream, filled by the activation warps
    act.kind = ga.kind;
    act.y = ga.y;
    act.ldy = ga.ldy;
    ga.fused = true;
  }
  const int64_t items = split_from + (tiles - split_from) * splits;
  const int grid = int(2 * (items < pairs ? items : pairs));
  // wave alignment of the TMA producers: 1 = whenever there is more than one wave; 2 (default) = large problems (>= 4 waves of >= 128
  // k-blocks and (M + N) * K >= 2^27, i.e. from 8192^3: the C5 layers, not the C2 ones).  Measured (profiles/r2_gemm_wave_sync.txt):
  // 16384^3 312 -> 385 TFLOP/s, DRAM reads 64.4 -> 17.8 GB, L2 hit rate 50 -> 76 %; 8192^3 +1 % alone, +2 % inside the power-capped
  // C5 step; the C2 shapes unchanged alone and 0..2 % slower in the step, hence the threshold.
  unsigned int* wave_sync = nullptr;
  if (items > pairs && (ctx->tc_wave_sync == 1 || (ctx->tc_wave_sync == 2 && items >= 4 * pairs && num_kb / splits >= 128 && (M + N) * K >= (int64_t(1) << 27)))) {
    wave_sync = ctx->counter + 12;  // its own 4 bytes of the context's counter block
    k_zero_u32<<<1, 1, 0, ctx->stream>>>(wave_sync);
    AGRS_LAUNCHED(ctx);
  }
  // L2 eviction policies of the operand loads.  The raster walks 8 row-blocks of tiles per column sweep, so op(B) is read once per
  // group of 8 row-blocks: when it is small enough to stay in the L2 next to the streams (and there is more than one group), its
  // loads ask to be evicted last.  AGRS_TC_L2_HINT: 0 off, 1 that rule, 2 / 3 / 4 force (B last / B last + A first / A first).
  uint32_t l2_hint = 0;
  {
    const int64_t m_tiles = (M + 2 * BM - 1) / (2 * BM);
    const double b_bytes = double(N) * double(K) * 4.0;
    if (ctx->tc_l2_hint == 1) l2_hint = (m_tiles > 8 && b_bytes >= double(8 << 20) && b_bytes <= double(80 << 20)) ? 1u : 0u;
    else if (ctx->tc_l2_hint == 2) l2_hint = 1u;
    else if (ctx->tc_l2_hint == 3) l2_hint = 3u;
    else if (ctx->tc_l2_hint == 4) l2_hint = 2u;
  }
  kern<<<grid, kThreads, SMEM_BYTES, ctx->stream>>>(ta, tb, C, ldc, bias, M, N, K, kb_per_chunk, splits, Cw, sc, tc_map, tw_map, epi_tma, a_max, b_max, ext_mx ? 1 : 0, wave_sync, split_from, act, l2_hint);
  AGRS_LAUNCHED(ctx);
  if (splits > 1 && split_from > 0) {  // only the tiles of the last wave were split
    const int64_t m_tiles = (M + 2 * BM - 1) / (2 * BM), n_tiles = (N + BN - 1) / BN;
    k_add_partials_tiles<<<unsigned((tiles - split_from) * 8), 256, 0, ctx->stream>>>(C, ldc, Cw, splits - 1, M, N, m_tiles, n_tiles, split_from, tiles);
    AGRS_LAUNCHED(ctx);
  } else if (splits > 1) {
    const int64_t blocks = (M * N / 4 + 255) / 256, cap = int64_t(ctx->sm_count) * 8;
    k_add_partials<<<unsigned(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap), 256, 0, ctx->stream>>>(C, ldc, Cw, splits - 1, M, N);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}

}  // namespace tc2

static int gemm_tc_2cta_impl(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                             const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias, const tc2::ScatterArgs* sc) {
  int rc = tc::resolve_encode(ctx);
  if (rc) return rc;
  const bool A_MN = transA != 0;  // A stored [K, M]: M contiguous
  const bool B_MN = transB == 0;  // B stored [K, N]: N contiguous
  CUtensorMap ta, tb;
  if (A_MN) rc = tc::encode_2d(ctx, &ta, A, uint64_t(M), uint64_t(K), uint64_t(lda), 32, 32, true);
  else      rc = tc::encode_2d(ctx, &ta, A, uint64_t(K), uint64_t(M), uint64_t(lda), 32, tc::BM, false);
  if (rc) return rc;
  if (B_MN) rc = tc::encode_2d(ctx, &tb, B, uint64_t(N), uint64_t(K), uint64_t(ldb), 32, 32, true);
  else      rc = tc::encode_2d(ctx, &tb, B, uint64_t(K), uint64_t(N), uint64_t(ldb), 32, uint32_t(tc2::BNH), false);
  if (rc) return rc;
  const int64_t num_kb = (K + tc::BK - 1) / tc::BK;
  const int kpc = ctx->tc_kchunk > 0 ? int((ctx->tc_kchunk + tc::BK - 1) / tc::BK) : int(num_kb);
  static const tc2::ScatterArgs none{};
  // mode 5 (the default): the fp16 split when the caller announced the operands' magnitudes (the host engine keeps one per
  // buffer, reported by the kernel that wrote it), bf16 correction terms when it did not -- no magnitude passes either way
  const int cross = ctx->tc_cross == 5 ? (ctx->gemm_a_mx && ctx->gemm_b_mx ? 4 : 1) : ctx->tc_cross;
  // forward activation in the epilogue (agrs_gemm_bias_act_f32): built for the two split schemes the default mode chooses between
  // and opt-in (AGRS_TUNE_TC_ACT_WARPS): measured time-neutral on C5 and 1-2 % slower on C2 against the activation as its own
  // launch -- see DESIGN 4.1 -- so by default the entry point runs the activation as a second pass after this kernel
  const bool want_act = !sc && ctx->tc_act_warps && ctx->gemm_act.kind != AGRS_ACT_NONE && (cross == 1 || cross == 4);
#define AGRS_PAIR_X(AM, BMJ, X)                                                                                     \
  (sc ? tc2::launch<AM, BMJ, true, X>(ctx, ta, tb, C, ldc, bias, M, N, K, kpc, *sc, A, lda, B, ldb)                  \
      : tc2::launch<AM, BMJ, false, X>(ctx, ta, tb, C, ldc, bias, M, N, K, kpc, none, A, lda, B, ldb))
#define AGRS_PAIR_ACT(AM, BMJ, X) tc2::launch<AM, BMJ, false, X, true>(ctx, ta, tb, C, ldc, bias, M, N, K, kpc, none, A, lda, B, ldb)
#define AGRS_PAIR(AM, BMJ)                                                                      \
  (want_act ? (cross == 1 ? AGRS_PAIR_ACT(AM, BMJ, 1) : AGRS_PAIR_ACT(AM, BMJ, 4))             \
   : cross == 1 ? AGRS_PAIR_X(AM, BMJ, 1) : cross == 2 ? AGRS_PAIR_X(AM, BMJ, 2) \
   : cross == 3 ? AGRS_PAIR_X(AM, BMJ, 3) : cross == 4 ? AGRS_PAIR_X(AM, BMJ, 4) : AGRS_PAIR_X(AM, BMJ, 0))
  if (A_MN && B_MN) return AGRS_PAIR(true, true);
  if (A_MN) return AGRS_PAIR(true, false);
  if (B_MN) return AGRS_PAIR(false, true);
  return AGRS_PAIR(false, false);
#undef AGRS_PAIR
#undef AGRS_PAIR_ACT
#undef AGRS_PAIR_X
}

int agrs_gemm_tc_2cta(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                      const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias) {
  return gemm_tc_2cta_impl(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, nullptr);
}

// fused_dp.cu
extern "C" int agrs_fused_scatter_by_copy(agrs_fused* f);
void agrs_fused_scatter_target(const agrs_fused* f, float* const** bases, int* world, int* rank, size_t* g_off, size_t* shard, int* gshift);

extern "C" int agrs_gemm_f32_scatter(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                                     const float* B, int64_t ldb, float* C, const float* bias, agrs_fused* f, size_t off) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_gemm_f32_scatter");
  const uint32_t *a_mx = ctx->gemm_a_mx, *b_mx = ctx->gemm_b_mx;  // agrs_gemm_operand_absmax: consumed by this call
  ctx->gemm_a_mx = ctx->gemm_b_mx = nullptr;
  AGRS_REQUIRE(ctx, f, AGRS_ERR_INVALID, "agrs_gemm_f32_scatter: null exchange");
  tc2::ScatterArgs sc{};
  float* const* bases = nullptr;
  agrs_fused_scatter_target(f, &bases, &sc.world, &sc.rank, &sc.g_off, &sc.shard, &sc.gshift);
  for (int p = 0; p < sc.world; ++p) sc.base[p] = bases[p];
  sc.off = off;
  // the pair kernel's vector epilogue must own whole 32-float pieces that never straddle two owners
  const bool fused_ok = ctx->gemm_path != AGRS_GEMM_SIMT && ctx->tc_2cta && M > 128 && K > 0 && N % 32 == 0 && off % 32 == 0 && sc.shard % 32 == 0 &&
                        agrs_aligned16(C) && agrs_gemm_tc_eligible(transA, transB, M, N, K, A, lda, B, ldb, C, N, bias) &&
                        (M * N) * K >= ctx->tc_min_flops;
  // When splitting K fills the last wave of CTA pairs (dW of the 4096-wide layers: +7 %), the final values only exist after the
  // partials are added: that shape keeps the split-K GEMM and pushes the result with the scatter kernel instead.
  const int64_t tiles = ((M + 2 * tc::BM - 1) / (2 * tc::BM)) * ((N + tc2::BN - 1) / tc2::BN);
  const bool wants_split = tc2::choose_splits(tiles, ctx->sm_count / 2, (K + tc::BK - 1) / tc::BK, ctx->tc_splitk) > 1;
  // when the exchange moves gradients with the copy engines (bucket protocol, default) the GEMM keeps its staged TMA epilogue
  if (fused_ok && !wants_split && !agrs_fused_scatter_by_copy(f)) {
    ctx->last_gemm_path = AGRS_GEMM_TCGEN05;
    agrs_ctx::ProfRec rec;  // timed like every other GEMM launch (bench.py's roofline counts all of them)
    if (int prc = agrs_prof_begin(ctx, AGRS_GEMM_TCGEN05, 2.0 * double(M) * double(N) * double(K), &rec)) return prc;
    ctx->gemm_a_mx = a_mx;
    ctx->gemm_b_mx = b_mx;
    const int grc = gemm_tc_2cta_impl(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, N, bias, &sc);
    ctx->gemm_a_mx = ctx->gemm_b_mx = nullptr;
    if (int prc = agrs_prof_end(ctx, rec)) return prc;
    return grc;
  }
  ctx->gemm_a_mx = a_mx;
  ctx->gemm_b_mx = b_mx;
  int rc = agrs_gemm_f32(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, N, bias);
  if (rc) return rc;
  return agrs_fused_scatter_f32(f, C, off, size_t(M) * size_t(N));
}
