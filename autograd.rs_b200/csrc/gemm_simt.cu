// SIMT f32 GEMM for the shapes the TMA/tcgen05 kernel cannot map: tiny, skinny or unaligned
// operands (fit_sine: K=3,N=1; LeNet: K=25,N=6 and N=10; anything whose leading dimension is not
// a multiple of 4 floats).  Plain FFMA, f32 accumulate, any M/N/K, all four operand-major
// combinations.  Skinny outputs with a long K (dW of the conv GEMM: 25x6 with K = B*P) are split
// along K into a workspace and folded in a fixed order (deterministic).
#include "agrs_internal.h"
#include "act.cuh"
#include "mag.cuh"

namespace {

// forward activation next to the pre-activation (agrs_gemm_bias_act_f32): y = act(c) written by the thread that holds c
struct SimtAct {
  int kind;       // AGRS_ACT_NONE: nothing to do
  float* y;
  int64_t ldy;
  uint32_t* mx;   // magnitude word of y, or null
};
__device__ __forceinline__ float simt_act(int kind, float v) { return kind == AGRS_ACT_RELU ? agrs_act_relu(v) : agrs_act_sigmoid(v); }

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;  // 256 threads, 4x4 micro-tile each

// C[m,n] = sum_k A(m,k) * B(k,n);  A(m,k) = A[m*lda+k] or A[k*lda+m] (transA);  B(k,n) = B[k*ldb+n] or B[n*ldb+k] (transB)
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm_simt(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda,
                                                   const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
                                                   const float* __restrict__ bias, int64_t k_per_split, float* __restrict__ part,
                                                   const SimtAct act) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid
  const int64_t m0 = int64_t(blockIdx.y) * BM, n0 = int64_t(blockIdx.x) * BN;
  const int64_t kb = int64_t(blockIdx.z) * k_per_split;
  const int64_t ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  float acc[TM][TN] = {};
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    // stage A tile: BM x BK, choose the thread->element map so global reads are coalesced
#pragma unroll
    for (int r = 0; r < (BM * BK) / 256; ++r) {
      int e = tid + r * 256;
      int mm, kk;
      if (TA) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
      int64_t gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < ke) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int r = 0; r < (BN * BK) / 256; ++r) {
      int e = tid + r * 256;
      int nn, kk;
      if (TB) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int64_t gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < ke) v = TB ? B[gn * ldb + gk] : B[gk * ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  uint32_t y_mag = 0;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      if (part) {
        part[(int64_t(blockIdx.z) * M + gm) * N + gn] = acc[i][j];
      } else {
        float v = acc[i][j];
        if (bias) v += bias[gn];
        C[gm * ldc + gn] = v;
        if (act.kind) {
          const float y = simt_act(act.kind, v);
          act.y[gm * act.ldy + gn] = y;
          y_mag = max(y_mag, mag_bits(y));
        }
      }
    }
  }
  if (act.kind && act.mx) mag_commit(y_mag, act.mx);
}

__global__ void k_fold_splitk(const float* __restrict__ part, int splits, int64_t M, int64_t N, float* __restrict__ C, int64_t ldc,
                              const float* __restrict__ bias, const SimtAct act) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t y_mag = 0;
  if (i < M * N) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[int64_t(z) * M * N + i];
    int64_t m = i / N, n = i % N;
    if (bias) s += bias[n];
    C[m * ldc + n] = s;
    if (act.kind) {
      const float y = simt_act(act.kind, s);
      act.y[m * act.ldy + n] = y;
      y_mag = mag_bits(y);
    }
  }
  if (act.kind && act.mx) mag_commit(y_mag, act.mx);
}

__global__ void k_bias_rows(int64_t M, int64_t N, float* C, int64_t ldc, const float* bias) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  C[(i / N) * ldc + (i % N)] = bias ? bias[i % N] : 0.f;
}

}  // namespace

int agrs_gemm_simt(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                   const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias) {
  if (M == 0 || N == 0) return AGRS_OK;
  // agrs_gemm_bias_act_f32: both kernels that write C hold its final value in a register, so they write act(C) as well
  SimtAct act{AGRS_ACT_NONE, nullptr, 0, nullptr};
  if (ctx->gemm_act.kind != AGRS_ACT_NONE && K > 0) {
    act.kind = ctx->gemm_act.kind;
    act.y = ctx->gemm_act.y;
    act.ldy = ctx->gemm_act.ldy;
    if (int rc = agrs_take_mx(ctx, &act.mx)) return rc;
    ctx->gemm_act.fused = true;
  }
  if (K == 0) {  // empty inner dimension: C = bias (ndarray: dot over an empty axis is 0)
    k_bias_rows<<<unsigned((M * N + 255) / 256), 256, 0, ctx->stream>>>(M, N, C, ldc, bias);
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  int64_t gx = (N + BN - 1) / BN, gy = (M + BM - 1) / BM;
  int64_t tiles = gx * gy;
  int splits = 1;
  if (tiles < ctx->sm_count && K >= 8 * BK) {
    int64_t want = (2LL * ctx->sm_count + tiles - 1) / tiles;
    int64_t maxs = K / (4 * BK);
    splits = int(want < maxs ? want : maxs);
    if (splits > 512) splits = 512;
    if (splits < 1) splits = 1;
  }
  int64_t kps = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  splits = int((K + kps - 1) / kps);
  float* part = nullptr;
  if (splits > 1) {
    int rc = agrs_ensure_workspace(ctx, size_t(splits) * size_t(M) * size_t(N) * sizeof(float));
    if (rc) return rc;
    part = ctx->workspace;
  }
  // grid.y is limited to 65535 blocks: fold very tall problems by looping in row slabs
  const int64_t slab_rows = 65535LL * BM;
  for (int64_t r0 = 0; r0 < M; r0 += slab_rows) {
    int64_t rows = (M - r0 < slab_rows) ? M - r0 : slab_rows;
    dim3 grid((unsigned)gx, (unsigned)((rows + BM - 1) / BM), (unsigned)splits);
    const float* Ar = transA ? A + r0 : A + r0 * lda;
    float* Cr = C + r0 * ldc;
    float* pr = part;  // split-K only happens when tiles < sm_count, i.e. a single slab
    SimtAct act_r = act;
    if (splits > 1) act_r.kind = AGRS_ACT_NONE;  // the fold kernel holds the final values
    else if (act.kind) act_r.y = act.y + r0 * act.ldy;
#define AGRS_SIMT(TA_, TB_) \
  k_gemm_simt<TA_, TB_><<<grid, 256, 0, ctx->stream>>>(rows, N, K, Ar, lda, B, ldb, Cr, ldc, splits > 1 ? nullptr : bias, kps, pr, act_r)
    if (transA && transB) AGRS_SIMT(true, true);
    else if (transA) AGRS_SIMT(true, false);
    else if (transB) AGRS_SIMT(false, true);
    else AGRS_SIMT(false, false);
#undef AGRS_SIMT
    AGRS_LAUNCHED(ctx);
  }
  if (splits > 1) {
    k_fold_splitk<<<unsigned((M * N + 255) / 256), 256, 0, ctx->stream>>>(part, splits, M, N, C, ldc, bias, act);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}
