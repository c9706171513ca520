// f32-faithful GEMM on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma
// (kind::tf32, 3xTF32 split) -> TMEM accumulator -> tcgen05.ld epilogue (+bias) -> global.
//
// Replaces the CUDA arm of Tensor::dot (reference src/tensor/tensor.rs:319) for the three GEMMs of
// Linear (operators.rs:144,159,160): forward X*W, dX = G*W^T, dW = X^T*G.  The transposes are operand
// MAJORS of the UMMA shared-memory descriptors, never copies.
//
// 3xTF32: x = hi + lo with hi = x truncated to tf32 (the tensor core reads the top 19 bits of the
// 32-bit container) and lo = tf32(x - hi);  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, accumulated in
// f32 in TMEM.  The dropped lo*lo term and the rounding of lo are ~2^-21 relative: f32-level results.
//
// One persistent CTA per SM, 12 warps, warp-specialised:
//   warp 0      TMA producer: raw f32 tiles of A and B into a STAGES-deep ring (128-byte swizzle)
//   warps 8-11  splitters: read the landed tile, write the `lo` tile next to it (same swizzled offsets)
//   warp 1      MMA issuer: 4 k-slices x 3 tcgen05.mma per stage, tcgen05.commit frees the stage
//   warp 2      TMEM allocator (2 accumulator stages x BN columns)
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns, +bias, vector stores; overlaps the next tile
// Every mbarrier wait is bounded: a protocol bug traps the kernel instead of hanging the GPU.
#include "tc_common.cuh"

namespace tc {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;    // warps 4..7  (warp % 4 selects the TMEM lane quarter)
constexpr int kSplitWarp0 = 8;  // warps 8..11
constexpr int kSplitWarps = 4;

template <int BN, int STAGES>
struct Cfg {
  static constexpr uint32_t A_BYTES = BM * BK * 4;
  static constexpr uint32_t B_BYTES = BN * BK * 4;
  static constexpr uint32_t STAGE_BYTES = 2 * (A_BYTES + B_BYTES);  // hi + lo of both operands
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  static_assert(TMEM_COLS == 512 || TMEM_COLS == 256 || TMEM_COLS == 128, "TMEM columns must be a power of two");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <bool A_MN, bool B_MN, int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_3xtf32(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* C,
              int64_t ldc, const float* __restrict__ bias, int64_t M, int64_t N, int64_t K, int explicit_hi, int kb_per_chunk) {
  using cfg = Cfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * cfg::STAGE_BYTES;
  // barrier map (8 B each): full[s], split[s], empty[s], tmem_full[2], tmem_empty[2], then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
  const int64_t num_tiles = m_tiles * n_tiles;
  const int num_kb = int((K + BK - 1) / BK);
  // The tensor core's f32 accumulator truncates on every MMA (bias ~ number of accumulation steps).  K is therefore
  // cut into chunks: each chunk accumulates in its own TMEM stage and the epilogue warps fold it into C with
  // round-to-nearest f32 adds (the C tile is re-read from L2).  kb_per_chunk = k-blocks (of 32) per chunk.
  const int num_chunks = (num_kb + kb_per_chunk - 1) / kb_per_chunk;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), kSplitWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int64_t mb, nb;
        tile_coords(t, m_tiles, n_tiles, mb, nb);
        const int32_t m0 = int32_t(mb * BM), n0 = int32_t(nb * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_hi = smem_base + s * cfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + 2 * cfg::A_BYTES;
          mbar_arrive_expect_tx(full_bar(s), cfg::A_BYTES + cfg::B_BYTES);
          const int32_t k0 = kb * BK;
          if (A_MN) {  // A stored [K, M]: boxes of 32 (M, contiguous) x 32 (K rows)
#pragma unroll
            for (int g = 0; g < BM / 32; ++g) tma_load_2d(&tmap_a, full_bar(s), a_hi + g * 4096, m0 + g * 32, k0);
          } else {     // A stored [M, K]: one box of 32 (K, contiguous) x 128 rows
            tma_load_2d(&tmap_a, full_bar(s), a_hi, k0, m0);
          }
          if (B_MN) {  // B stored [K, N]
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) tma_load_2d(&tmap_b, full_bar(s), b_hi + g * 4096, n0 + g * 32, k0);
          } else {     // B stored [N, K]
            tma_load_2d(&tmap_b, full_bar(s), b_hi, k0, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc<A_MN, B_MN, BN>();
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(tempty_bar(acc), acc_ph ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        const int kb_end = (ch + 1) * kb_per_chunk < num_kb ? (ch + 1) * kb_per_chunk : num_kb;
        for (int kb = ch * kb_per_chunk; kb < kb_end; ++kb) {
          mbar_wait(split_bar(s), ph);  // TMA landed AND lo tiles written
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_hi = smem_base + s * cfg::STAGE_BYTES;
            const uint32_t a_lo = a_hi + cfg::A_BYTES;
            const uint32_t b_hi = a_hi + 2 * cfg::A_BYTES;
            const uint32_t b_lo = b_hi + cfg::B_BYTES;
            const bool first_kb = (kb == ch * kb_per_chunk);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da_hi = make_smem_desc<A_MN>(a_hi + k * k_slice_bytes<A_MN>());
              const uint64_t da_lo = make_smem_desc<A_MN>(a_lo + k * k_slice_bytes<A_MN>());
              const uint64_t db_hi = make_smem_desc<B_MN>(b_hi + k * k_slice_bytes<B_MN>());
              const uint64_t db_lo = make_smem_desc<B_MN>(b_lo + k * k_slice_bytes<B_MN>());
              umma_tf32(d_tmem, da_lo, db_hi, idesc, (first_kb && k == 0) ? 0u : 1u);  // small terms first
              umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
              umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
            }
            umma_commit(empty_bar(s));  // fires when the MMAs above have finished reading this stage
            if (kb == kb_end - 1) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
    }
  } else if (warp >= kSplitWarp0) {
    // ===================== splitters: lo = tf32(x - trunc(x)) =====================
    const int tid = threadIdx.x - kSplitWarp0 * 32;  // 0..127
    int s = 0;
    uint32_t ph = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        const uint32_t stage = smem_base + s * cfg::STAGE_BYTES;
        // A: hi at [0, A_BYTES), lo at [A_BYTES, 2A);  B: hi at [2A, 2A+B), lo at [2A+B, 2A+2B)
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          const uint32_t hi_base = stage + (part ? 2 * cfg::A_BYTES : 0);
          const uint32_t bytes = part ? cfg::B_BYTES : cfg::A_BYTES;
          const uint32_t lo_base = hi_base + bytes;
#pragma unroll 4
          for (uint32_t off = tid * 16; off < bytes; off += kSplitWarps * 32 * 16) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(hi_base + off));
            float4 h, l;
            l.x = split_lo(v.x, h.x);
            l.y = split_lo(v.y, h.y);
            l.z = split_lo(v.z, h.z);
            l.w = split_lo(v.w, h.w);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo_base + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
            if (explicit_hi)
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(hi_base + off), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(s));
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp - kEpiWarp0;  // == warp % 4: TMEM lanes [32q, 32q+32)
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int64_t mb, nb;
      tile_coords(t, m_tiles, n_tiles, mb, nb);
      const int64_t row = mb * BM + q * 32 + lane;
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_wait_ld();
          const int64_t col0 = nb * BN + c * 32;
          if (row < M && col0 < N) {
            float* dst = C + row * ldc + col0;
            if (vec_ok && col0 + 32 <= N) {
              float4 base[8];
              if (ch == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  base[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
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
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) {
                  float b = ch == 0 ? (bias ? __ldg(bias + col0 + j) : 0.f) : dst[j];
                  dst[j] = b + __uint_as_float(v[j]);
                }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cfg::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
template <bool A_MN, bool B_MN, int BN, int STAGES>
static int launch(agrs_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, float* C, int64_t ldc, const float* bias, int64_t M,
                  int64_t N, int64_t K, int explicit_hi, int kb_per_chunk) {
  using cfg = Cfg<BN, STAGES>;
  auto kern = k_gemm_3xtf32<A_MN, B_MN, BN, STAGES>;
  AGRS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cfg::SMEM_BYTES)));
  int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int grid = int(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  kern<<<grid, kThreads, cfg::SMEM_BYTES, ctx->stream>>>(ta, tb, C, ldc, bias, M, N, K, explicit_hi, kb_per_chunk);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

}  // namespace tc

bool agrs_gemm_tc_eligible(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                           int64_t ldb, float* C, int64_t ldc, const float* bias) {
  (void)transA; (void)transB; (void)C; (void)ldc;
  if (M < 1 || N < 1 || K < 1) return false;
  if (!agrs_aligned16(A) || !agrs_aligned16(B)) return false;  // TMA: 16-byte aligned base ...
  if ((lda & 3) || (ldb & 3)) return false;                     // ... and row pitch
  if (bias && !agrs_aligned16(bias)) return false;
  if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31)) return false;
  return true;
}

int agrs_gemm_tc(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                 int64_t ldb, float* C, int64_t ldc, const float* bias) {
  int rc = tc::resolve_encode(ctx);
  if (rc) return rc;
  const bool A_MN = transA != 0;   // A stored [K, M]: M contiguous
  const bool B_MN = transB == 0;   // B stored [K, N]: N contiguous
  const int BN = ctx->tc_bn;       // 256 (2 stages) or 128 (3 stages)
  CUtensorMap ta, tb;
  if (A_MN) rc = tc::encode_2d(ctx, &ta, A, uint64_t(M), uint64_t(K), uint64_t(lda), 32, 32, true);
  else      rc = tc::encode_2d(ctx, &ta, A, uint64_t(K), uint64_t(M), uint64_t(lda), 32, tc::BM, false);
  if (rc) return rc;
  if (B_MN) rc = tc::encode_2d(ctx, &tb, B, uint64_t(N), uint64_t(K), uint64_t(ldb), 32, 32, true);
  else      rc = tc::encode_2d(ctx, &tb, B, uint64_t(K), uint64_t(N), uint64_t(ldb), 32, uint32_t(BN), false);
  if (rc) return rc;
  const int eh = ctx->tc_explicit_hi;
  const int64_t num_kb = (K + tc::BK - 1) / tc::BK;
  const int kpc = ctx->tc_kchunk > 0 ? int((ctx->tc_kchunk + tc::BK - 1) / tc::BK) : int(num_kb);
#define AGRS_TC(AM, BMJ)                                                                                   \
  (BN == 256 ? tc::launch<AM, BMJ, 256, 2>(ctx, ta, tb, C, ldc, bias, M, N, K, eh, kpc)                       \
             : tc::launch<AM, BMJ, 128, 3>(ctx, ta, tb, C, ldc, bias, M, N, K, eh, kpc))
  if (A_MN && B_MN) return AGRS_TC(true, true);
  if (A_MN) return AGRS_TC(true, false);
  if (B_MN) return AGRS_TC(false, true);
  return AGRS_TC(false, false);
#undef AGRS_TC
}
