// HBM-bound f32 GEMMs with a skinny output: the Conv2D lowering and the narrow Linear layers of the LeNet-style
// configuration (reference operators.rs:378-379, :422, :445 and :144, :160 at N = 6, 10, 25).  The generic 64x64-tile SIMT
// kernel wastes 90 % of its tile on them; these kernels stream the big operand once, coalesced, and keep the small one on chip.
//   rows    C[M,N] = A[M,K] . op(B)      K <= 64, N <= 32, M large       one thread per output row      (conv forward, dxcol)
//   dot     C[M,N] = A[M,K] . op(B)      N <= 16, K long                 one warp per output row        (Linear 3456 -> 10)
//   tn      C[M,N] = A[K,M]^T . B[K,N]   N <= 16, any M, K               split-K over blocks, fixed-order fold   (dW of Linear)
//   tn_small  the same with M*N <= 160 and K huge                        staged row streams, one thread per output (conv filter gradient)
// Plain FFMA in f32; summation order is fixed (deterministic).
#include "agrs_internal.h"

namespace {

// ------------------------------------------------------------------ rows: one thread per output row
// Block = 256 rows.  The A tile [256, K] is staged through shared memory with coalesced loads (row pitch K|1: odd, so thread t
// reading row t is bank-conflict free); B (K x N, padded to NT columns) and the bias sit in shared memory and are read as
// broadcasts; the C tile goes back through shared memory so the store is coalesced too.
template <int NT>
__global__ void __launch_bounds__(256) k_gemm_rows(int64_t M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                   int64_t ldb, int transB, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias) {
  extern __shared__ float sm[];
  const int pitch = K | 1, cp = N | 1;
  float* As = sm;                                                  // [256][pitch]; reused as Cs [256][cp]
  float* Bs = sm + 256 * (pitch > cp ? pitch : cp);                // [K][NT]
  float* bs = Bs + K * NT;                                         // [NT]
  const int tid = threadIdx.x;
  for (int e = tid; e < K * NT; e += 256) {
    const int k = e / NT, n = e - k * NT;
    Bs[e] = n < N ? (transB ? B[int64_t(n) * ldb + k] : B[int64_t(k) * ldb + n]) : 0.f;
  }
  for (int n = tid; n < NT; n += 256) bs[n] = (bias && n < N) ? bias[n] : 0.f;
  const int64_t tiles = (M + 255) / 256;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t m0 = tile * 256;
    const int rows = int(M - m0 < 256 ? M - m0 : 256);
    __syncthreads();  // Bs ready / previous tile's Cs drained
    if (lda == K) {   // the tile is one contiguous run of rows*K floats
      const float* src = A + m0 * K;
      const int total = rows * K;
      if (pitch == K && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {  // odd K: shared layout == global layout, plain 16-byte copy
        for (int e4 = tid; e4 < total / 4; e4 += 256) reinterpret_cast<float4*>(As)[e4] = __ldcs(reinterpret_cast<const float4*>(src) + e4);
        for (int e = (total / 4) * 4 + tid; e < total; e += 256) As[e] = src[e];
      } else if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        for (int e4 = tid; e4 < total / 4; e4 += 256) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + e4);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = 4 * e4 + j, r = e / K;
            As[r * pitch + (e - r * K)] = vv[j];
          }
        }
        for (int e = (total / 4) * 4 + tid; e < total; e += 256) {
          const int r = e / K;
          As[r * pitch + (e - r * K)] = src[e];
        }
      } else {
        for (int e = tid; e < total; e += 256) {
          const int r = e / K;
          As[r * pitch + (e - r * K)] = __ldcs(src + e);
        }
      }
    } else {
      for (int e = tid; e < rows * K; e += 256) {
        const int r = e / K, k = e - r * K;
        As[r * pitch + k] = A[(m0 + r) * lda + k];
      }
    }
    __syncthreads();
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
    if (tid < rows) {
      const float* a = As + tid * pitch;
      for (int k = 0; k < K; ++k) {
        const float av = a[k];
        const float4* b4 = reinterpret_cast<const float4*>(Bs + k * NT);
#pragma unroll
        for (int n4 = 0; n4 < NT / 4; ++n4) {
          const float4 b = b4[n4];
          acc[4 * n4] = fmaf(av, b.x, acc[4 * n4]);
          acc[4 * n4 + 1] = fmaf(av, b.y, acc[4 * n4 + 1]);
          acc[4 * n4 + 2] = fmaf(av, b.z, acc[4 * n4 + 2]);
          acc[4 * n4 + 3] = fmaf(av, b.w, acc[4 * n4 + 3]);
        }
      }
    }
    __syncthreads();  // everybody is done reading As
    float* Cs = As;
    if (tid < rows) {
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (n < N) Cs[tid * cp + n] = acc[n] + bs[n];
    }
    __syncthreads();
    if (ldc == N && cp == N && (reinterpret_cast<uintptr_t>(C + m0 * N) & 15u) == 0) {  // odd N: plain 16-byte copy out
      float* dst = C + m0 * N;
      const int total = rows * N;
      for (int e4 = tid; e4 < total / 4; e4 += 256) __stcs(reinterpret_cast<float4*>(dst) + e4, reinterpret_cast<const float4*>(Cs)[e4]);
      for (int e = (total / 4) * 4 + tid; e < total; e += 256) dst[e] = Cs[e];
    } else if (ldc == N) {
      float* dst = C + m0 * N;
      for (int e = tid; e < rows * N; e += 256) {
        const int r = e / N;
        __stcs(dst + e, Cs[r * cp + (e - r * N)]);
      }
    } else {
      for (int e = tid; e < rows * N; e += 256) {
        const int r = e / N, n = e - r * N;
        C[(m0 + r) * ldc + n] = Cs[r * cp + n];
      }
    }
  }
}

// ------------------------------------------------------------------ dot: one warp per output row, K long
// Block = 4 warps = 4 rows.  B is walked in chunks of 512 k staged in shared memory ([512][NT+1], so lane k reading row k is
// conflict free); every lane strides its row of A by 32 (coalesced) and keeps NT partial sums; a shuffle tree ends the row.
constexpr int kDotChunk = 512;
template <int NT>
__global__ void __launch_bounds__(128) k_gemm_dot(int64_t M, int N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                  int64_t ldb, int transB, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias) {
  __shared__ float Bs[kDotChunk][NT + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t groups = (M + 3) / 4;
  for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
    const int64_t m = g * 4 + w;
    const float* a = A + (m < M ? m : M - 1) * lda;
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
    for (int64_t k0 = 0; k0 < K; k0 += kDotChunk) {
      const int kc = int(K - k0 < kDotChunk ? K - k0 : kDotChunk);
      __syncthreads();
      if (transB) {  // B stored [N, K]: coalesced along k
        for (int e = threadIdx.x; e < kc * N; e += 128) {
          const int n = e / kc, k = e - n * kc;
          Bs[k][n] = __ldg(B + int64_t(n) * ldb + k0 + k);
        }
      } else {       // B stored [K, N]: rows of N floats
        for (int e = threadIdx.x; e < kc * N; e += 128) {
          const int k = e / N, n = e - k * N;
          Bs[k][n] = __ldg(B + (k0 + k) * ldb + n);
        }
      }
      __syncthreads();
      for (int k = lane; k < kc; k += 32) {
        const float av = __ldcs(a + k0 + k);
#pragma unroll
        for (int n = 0; n < NT; ++n)
          if (n < N) acc[n] = fmaf(av, Bs[k][n], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float v = acc[n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == n && n < N && m < M) C[m * ldc + n] = v + (bias ? bias[n] : 0.f);
    }
  }
}

// ------------------------------------------------------------------ dot, K sliced over CTAs, two output rows per warp
// The kernel above re-stages B in 512-row chunks for every four rows of C and feeds every FMA from its own shared-memory word:
// Linear(3456, 10) forward on 1024 rows took 47 us for 14 MB of A (profiles/r2_c3_timeline.md).  Here a CTA owns a slice of K
// (a few 128-wide steps) and a group of rows: it stages its slice of B once, transposed -- Bt[n][k], k contiguous -- and each warp
// walks row pairs of A with 16-byte coalesced loads: lane l holds k = 4l .. 4l+3 of a step, reads the matching four k of column n
// as one conflict-free vector and feeds 8 multiply-adds with it.  Lanes are folded by a shuffle tree, slices by the last CTA of
// the row group to finish, in slice order (ticket counter): one launch, no float atomics, deterministic.
template <int NT>
__global__ void __launch_bounds__(256) k_gemm_dot_sliced(int64_t M, int N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                         int64_t ldb, int transB, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                                         int steps_per_slice, int rows_per_cta, float* __restrict__ part, unsigned int* __restrict__ tickets) {
  extern __shared__ __align__(16) float Bt[];
  __shared__ unsigned int s_ticket;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int slice = blockIdx.x, slices = gridDim.x, rg = blockIdx.y;
  const int ks = steps_per_slice * 128, pitch = ks + 4;  // + 4: rows of Bt start in different bank groups
  const int64_t k0 = int64_t(slice) * ks;
  for (int kk = tid; kk < ks; kk += 256) {  // thread = one k of the slice: N consecutive floats of B in, N conflict-free stores out
    const int64_t k = k0 + kk;
#pragma unroll
    for (int n = 0; n < NT; ++n)
      if (n < N) Bt[n * pitch + kk] = k < K ? (transB ? __ldg(B + int64_t(n) * ldb + k) : __ldg(B + k * ldb + n)) : 0.f;
  }
  __syncthreads();
  const int64_t r0 = int64_t(rg) * rows_per_cta;
  const int rows = int(M - r0 < rows_per_cta ? M - r0 : rows_per_cta);
  for (int pr = w; 2 * pr < rows; pr += 8) {
    const int64_t m0 = r0 + 2 * pr, m1 = 2 * pr + 1 < rows ? m0 + 1 : m0;
    const float *a0 = A + m0 * lda + k0, *a1 = A + m1 * lda + k0;
    float acc0[NT], acc1[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc0[n] = acc1[n] = 0.f;
#pragma unroll 4
    for (int st = 0; st < steps_per_slice; ++st) {
      const int kk = st * 128 + 4 * lane;
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (k0 + kk + 3 < K) {  // K % 4 == 0 (checked by the host): a vector is whole or absent
        x0 = __ldcs(reinterpret_cast<const float4*>(a0 + kk));
        x1 = __ldcs(reinterpret_cast<const float4*>(a1 + kk));
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (n < N) {
          const float4 b = *reinterpret_cast<const float4*>(Bt + n * pitch + kk);
          acc0[n] = fmaf(x0.x, b.x, acc0[n]); acc0[n] = fmaf(x0.y, b.y, acc0[n]); acc0[n] = fmaf(x0.z, b.z, acc0[n]); acc0[n] = fmaf(x0.w, b.w, acc0[n]);
          acc1[n] = fmaf(x1.x, b.x, acc1[n]); acc1[n] = fmaf(x1.y, b.y, acc1[n]); acc1[n] = fmaf(x1.z, b.z, acc1[n]); acc1[n] = fmaf(x1.w, b.w, acc1[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float v0 = acc0[n], v1 = acc1[n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      }
      if (lane == n && n < N) {
        if (slices == 1) {
          const float bv = bias ? bias[n] : 0.f;
          C[m0 * ldc + n] = v0 + bv;
          if (m1 != m0) C[m1 * ldc + n] = v1 + bv;
        } else {
          part[(int64_t(slice) * M + m0) * N + n] = v0;
          if (m1 != m0) part[(int64_t(slice) * M + m1) * N + n] = v1;
        }
      }
    }
  }
  if (slices == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(tickets + rg, 1u);
  __syncthreads();
  if (s_ticket != unsigned(slices) - 1u) return;
  __threadfence();
  for (int e = tid; e < rows * N; e += 256) {  // the group's last CTA: slices added in order
    const int r = e / N, n = e - r * N;
    float t = bias ? bias[n] : 0.f;
    for (int sl = 0; sl < slices; ++sl) t += __ldcg(part + (int64_t(sl) * M + r0 + r) * N + n);
    C[(r0 + r) * ldc + n] = t;
  }
  if (tid == 0) tickets[rg] = 0;
}

// ------------------------------------------------------------------ tn_small: C[M,N] = A[K,M]^T . B[K,N] with M*N <= 160, K huge
// (the conv filter gradient: 25 x 6 outputs over K = B*P rows).  A and B are contiguous streams of short rows; a block stages
// 256 rows of each with coalesced loads, and output (m, n) is owned by ONE thread (two row-parity groups of 160 threads) that
// walks the staged rows in order.  Block partials go to a workspace; the last block folds them in block order.
constexpr int kTnRows = 256;
__global__ void __launch_bounds__(320) k_gemm_tn_small(int M, int N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                       int64_t ldb, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                                       int64_t rows_per_block, float* __restrict__ part, unsigned int* __restrict__ ticket) {
  extern __shared__ float sm[];
  float* As = sm;                       // [kTnRows][M]
  float* Bs = sm + kTnRows * M;         // [kTnRows][N]
  __shared__ float half[160];
  __shared__ unsigned int s_ticket;
  const int tid = threadIdx.x, grp = tid / 160, t = tid - grp * 160, MN = M * N;
  const int m = t / N, n = t - m * N;
  const int64_t k0 = int64_t(blockIdx.x) * rows_per_block, k1 = k0 + rows_per_block < K ? k0 + rows_per_block : K;
  float acc = 0.f;
  for (int64_t kb = k0; kb < k1; kb += kTnRows) {
    const int rows = int(k1 - kb < kTnRows ? k1 - kb : kTnRows);
    __syncthreads();
    if (lda == M) {
      const float* src = A + kb * M;
      for (int e = tid; e < rows * M; e += 320) As[e] = __ldcs(src + e);
    } else {
      for (int e = tid; e < rows * M; e += 320) As[e] = A[(kb + e / M) * lda + e % M];
    }
    if (ldb == N) {
      const float* src = B + kb * N;
      for (int e = tid; e < rows * N; e += 320) Bs[e] = __ldcs(src + e);
    } else {
      for (int e = tid; e < rows * N; e += 320) Bs[e] = B[(kb + e / N) * ldb + e % N];
    }
    __syncthreads();
    if (t < MN)
      for (int r = grp; r < rows; r += 2) acc = fmaf(As[r * M + m], Bs[r * N + n], acc);
  }
  __syncthreads();
  if (grp == 1 && t < MN) half[t] = acc;
  __syncthreads();
  const int splits = gridDim.x;
  if (grp == 0 && t < MN) {
    acc += half[t];
    if (splits == 1) C[int64_t(m) * ldc + n] = acc + (bias ? bias[n] : 0.f);
    else part[int64_t(blockIdx.x) * MN + t] = acc;
  }
  if (splits == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != unsigned(splits - 1)) return;
  __threadfence();
  // fold: thread (grp, t) sums the blocks z = grp, grp+2, ..; then the two halves in order
  float s = 0.f;
  if (t < MN)
    for (int z = grp; z < splits; z += 2) s += __ldcg(part + int64_t(z) * MN + t);
  if (grp == 1 && t < MN) half[t] = s;
  __syncthreads();
  if (grp == 0 && t < MN) C[int64_t(m) * ldc + n] = s + half[t] + (bias ? bias[n] : 0.f);
  if (tid == 0) *ticket = 0;
}

// ------------------------------------------------------------------ tn: C = A^T . B with A stored [K, M], split over K
// Block = 32 columns of A (m) x 8 k-lanes; grid.y tiles M, grid.x splits K.  Reads of A are coalesced along m, the matching row
// of B (N <= 16 floats) is a broadcast.  Partials [splits][M*N] are folded in split order by the last block of each m-tile.
template <int NT>
__global__ void __launch_bounds__(256) k_gemm_tn(int64_t M, int N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                 int64_t ldb, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int64_t k_per_split,
                                                 float* __restrict__ part, unsigned int* __restrict__ tickets) {
  __shared__ float red[8][32][NT + 1];
  __shared__ unsigned int s_ticket;
  const int mx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int64_t m = int64_t(blockIdx.y) * 32 + mx;
  const int64_t k0 = int64_t(blockIdx.x) * k_per_split, k1 = k0 + k_per_split < K ? k0 + k_per_split : K;
  float acc[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n] = 0.f;
  if (m < M) {
    for (int64_t k = k0 + ky; k < k1; k += 8) {
      const float av = __ldcs(A + k * lda + m);
      const float* b = B + k * ldb;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (n < N) acc[n] = fmaf(av, __ldg(b + n), acc[n]);
    }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) red[ky][mx][n] = acc[n];
  __syncthreads();
  const int splits = gridDim.x;
  // 32 x N outputs of this block: thread (mx, n = ky, ky+8, ..) folds the 8 k-lanes in order
  for (int n = ky; n < N; n += 8) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][mx][n];
    if (m < M) {
      if (splits == 1) C[m * ldc + n] = s + (bias ? bias[n] : 0.f);
      else part[(int64_t(blockIdx.x) * M + m) * N + n] = s;
    }
  }
  if (splits == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[blockIdx.y], 1u);
  __syncthreads();
  if (s_ticket != unsigned(splits - 1)) return;
  __threadfence();
  // fold the split partials: k-lane ky sums the splits z = ky, ky+8, .. (fixed assignment), then the 8 lanes fold in order
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    float s = 0.f;
    if (n < N && m < M)
      for (int z = ky; z < splits; z += 8) s += __ldcg(part + (int64_t(z) * M + m) * N + n);
    red[ky][mx][n] = s;
  }
  __syncthreads();
  for (int n = ky; n < N; n += 8) {
    if (m >= M) break;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][mx][n];
    C[m * ldc + n] = s + (bias ? bias[n] : 0.f);
  }
  if (threadIdx.x == 0) tickets[blockIdx.y] = 0;
}

// ------------------------------------------------------------------ tn, four columns of A per thread
// The same product with 16-byte loads: block = 32 lanes x 4 k-lanes, a lane owns 4 consecutive m (one float4 of a row of A, 512
// coalesced bytes per warp) and keeps 4 x N sums; the block's slice of B sits in shared memory, rows padded to NP = 4, 8, 12 or 16
// floats, and is read as broadcast vectors: 4 N multiply-adds per 1 + NP/4 loads where k_gemm_tn issues N per 1 + N.
// dW of Linear(3456, 10) on 1024 rows: 19 -> see profiles/r2_c3_timeline.md.  Needs M % 4 == 0, lda % 4 == 0, A 16-byte aligned.
template <int NT>
__global__ void __launch_bounds__(128) k_gemm_tn_v4(int64_t M, int N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                    int64_t ldb, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int64_t k_per_split,
                                                    float* __restrict__ part, unsigned int* __restrict__ tickets) {
  extern __shared__ __align__(16) float smv[];
  __shared__ unsigned int s_ticket;
  float* Bs = smv;                               // [k_per_split][NT]
  float* red = smv + k_per_split * NT;           // [4][32][4 * NT + 4]
  constexpr int RP = 4 * NT + 4;
  const int mx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int64_t m = (int64_t(blockIdx.y) * 32 + mx) * 4;
  const int64_t k0 = int64_t(blockIdx.x) * k_per_split, k1 = k0 + k_per_split < K ? k0 + k_per_split : K;
  const int kn = int(k1 - k0);
  for (int e = threadIdx.x; e < kn * NT; e += 128) {
    const int kk = e / NT, n = e - kk * NT;
    Bs[e] = n < N ? __ldg(B + (k0 + kk) * ldb + n) : 0.f;
  }
  __syncthreads();
  float acc[4][NT];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[q][n] = 0.f;
  if (m < M) {
#pragma unroll 4
    for (int kk = ky; kk < kn; kk += 4) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(A + (k0 + kk) * lda + m));
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int n4 = 0; n4 < NT / 4; ++n4) {
        const float4 b = *reinterpret_cast<const float4*>(Bs + kk * NT + 4 * n4);  // one address per warp: broadcast
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[q][4 * n4 + j] = fmaf(av[q], bv[j], acc[q][4 * n4 + j]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int n = 0; n < NT; ++n) red[(ky * 32 + mx) * RP + q * NT + n] = acc[q][n];
  __syncthreads();
  const int splits = gridDim.x;
  // 128 x N outputs of this block, element e = (column of the block, n): the 4 k-lanes folded in order
  for (int e = threadIdx.x; e < 128 * N; e += 128) {
    const int mc = e / N, n = e - mc * N, lane = mc >> 2, q = mc & 3;
    const int64_t mm = int64_t(blockIdx.y) * 128 + mc;
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) sum += red[(r * 32 + lane) * RP + q * NT + n];
    if (mm < M) {
      if (splits == 1) C[mm * ldc + n] = sum + (bias ? bias[n] : 0.f);
      else part[(int64_t(blockIdx.x) * M + mm) * N + n] = sum;
    }
  }
  if (splits == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[blockIdx.y], 1u);
  __syncthreads();
  if (s_ticket != unsigned(splits - 1)) return;
  __threadfence();
  // the m-tile's last block: thread t owns outputs e = t, t + 128, .. (N of them); splits are added in order, and the loads of one
  // split for all of a thread's outputs (and of the next split) are in flight together
  {
    float t[NT];
    const float* src[NT];
    const int64_t plane = M * N;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      t[i] = 0.f;
      const int e = threadIdx.x + 128 * i, mc = e / N, n = e - mc * N;
      const int64_t mm = int64_t(blockIdx.y) * 128 + mc;
      src[i] = (i < N && mm < M) ? part + mm * N + n : nullptr;
    }
#pragma unroll 2
    for (int z = 0; z < splits; ++z) {
      float v[NT];
#pragma unroll
      for (int i = 0; i < NT; ++i) v[i] = src[i] ? __ldcg(src[i] + z * plane) : 0.f;
#pragma unroll
      for (int i = 0; i < NT; ++i) t[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (!src[i]) continue;
      const int e = threadIdx.x + 128 * i, mc = e / N, n = e - mc * N;
      C[(int64_t(blockIdx.y) * 128 + mc) * ldc + n] = t[i] + (bias ? bias[n] : 0.f);
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.y] = 0;
}

int ensure_tickets(agrs_ctx* ctx, size_t need) {
  if (need <= ctx->tickets_n) return AGRS_OK;
  AGRS_NOT_CAPTURING(ctx, "growing the ticket array (run one ordinary step before capturing)");
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->tickets) AGRS_CUDA(ctx, cudaFree(ctx->tickets));
  ctx->tickets = nullptr;
  ctx->tickets_n = 0;
  size_t n = need < 4096 ? 4096 : need;
  AGRS_CUDA(ctx, cudaMalloc(&ctx->tickets, n * sizeof(unsigned int)));
  AGRS_CUDA(ctx, cudaMemsetAsync(ctx->tickets, 0, n * sizeof(unsigned int), ctx->stream));
  ctx->tickets_n = n;
  return AGRS_OK;
}

}  // namespace

bool agrs_gemm_skinny_eligible(int transA, int transB, int64_t M, int64_t N, int64_t K) {
  if (M < 1 || N < 1 || K < 1) return false;
  if (transA) return !transB && N <= 16;                       // tn
  if (K <= 64 && N <= 32 && M >= 1024) return true;            // rows
  return N <= 16 && K > 64;                                    // dot
}

int agrs_gemm_skinny(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float* C, int64_t ldc, const float* bias) {
  const int n = int(N);
  // tn_small stages kTnRows rows of A and of B: 1 KB of dynamic shared memory per (M + N).  It is only taken while that fits the
  // 48 KB every kernel gets without an opt-in (next to its 0.7 KB of static arrays); wider shapes (Linear(100, 1): M + N = 101)
  // go to k_gemm_tn below.
  if (transA && M * N <= 160 && K >= 4096 && M + N <= 46) {  // tn_small
    const int64_t blocks_wanted = int64_t(ctx->sm_count) * 4, max_blocks = (K + kTnRows - 1) / kTnRows;
    int64_t blocks = blocks_wanted < max_blocks ? blocks_wanted : max_blocks;
    int64_t rpb = ((K + blocks - 1) / blocks + kTnRows - 1) / kTnRows * kTnRows;
    blocks = (K + rpb - 1) / rpb;
    float* part = nullptr;
    if (blocks > 1) {
      int rc = agrs_ensure_workspace(ctx, size_t(blocks) * size_t(M * N) * sizeof(float));
      if (rc) return rc;
      rc = ensure_tickets(ctx, 1);
      if (rc) return rc;
      part = ctx->workspace;
    }
    const size_t smem = size_t(kTnRows) * size_t(M + N) * sizeof(float);
    k_gemm_tn_small<<<unsigned(blocks), 320, smem, ctx->stream>>>(int(M), n, K, A, lda, B, ldb, C, ldc, bias, rpb, part, ctx->tickets);
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  if (transA && M % 4 == 0 && lda % 4 == 0 && agrs_aligned16(A) && M >= 512 && K >= 64) {  // tn with 16-byte loads
    const int64_t my = (M + 127) / 128;
    if (my <= 65535) {
      // about two blocks per SM, at least 32 rows of K per block
      int64_t want = (2 * int64_t(ctx->sm_count) + my - 1) / my, maxs = (K + 31) / 32;
      int64_t splits = want < maxs ? want : maxs;
      if (splits < 1) splits = 1;
      int64_t kps = (K + splits - 1) / splits;
      kps = (kps + 3) / 4 * 4;
      splits = (K + kps - 1) / kps;
      const int nt = n <= 4 ? 4 : (n <= 8 ? 8 : (n <= 12 ? 12 : 16));
      const size_t smem = (size_t(kps) * nt + size_t(128) * (4 * nt + 4)) * sizeof(float);
      if (smem <= 48 * 1024) {
        float* part = nullptr;
        if (splits > 1) {
          int rc = agrs_ensure_workspace(ctx, size_t(splits) * size_t(M) * size_t(N) * sizeof(float));
          if (rc) return rc;
          rc = ensure_tickets(ctx, size_t(my));
          if (rc) return rc;
          part = ctx->workspace;
        }
        dim3 grid((unsigned)splits, (unsigned)my);
        if (nt == 4)       k_gemm_tn_v4<4><<<grid, 128, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, C, ldc, bias, kps, part, ctx->tickets);
        else if (nt == 8)  k_gemm_tn_v4<8><<<grid, 128, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, C, ldc, bias, kps, part, ctx->tickets);
        else if (nt == 12) k_gemm_tn_v4<12><<<grid, 128, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, C, ldc, bias, kps, part, ctx->tickets);
        else               k_gemm_tn_v4<16><<<grid, 128, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, C, ldc, bias, kps, part, ctx->tickets);
        AGRS_LAUNCHED(ctx);
        return AGRS_OK;
      }
    }
  }
  if (transA) {
    const int64_t my = (M + 31) / 32;
    AGRS_REQUIRE(ctx, my <= 65535, AGRS_ERR_UNSUPPORTED, "agrs_gemm_skinny(tn): M too large");
    // split K so the grid fills the machine (~4 blocks per SM), at least 64 rows of K per block
    int64_t want = (int64_t(ctx->sm_count) * 4 + my - 1) / my, maxs = (K + 63) / 64;
    int64_t splits = want < maxs ? want : maxs;
    if (splits < 1) splits = 1;
    if (splits > 2048) splits = 2048;
    int64_t kps = (K + splits - 1) / splits;
    kps = (kps + 7) / 8 * 8;
    splits = (K + kps - 1) / kps;
    float* part = nullptr;
    if (splits > 1) {
      int rc = agrs_ensure_workspace(ctx, size_t(splits) * size_t(M) * size_t(N) * sizeof(float));
      if (rc) return rc;
      rc = ensure_tickets(ctx, size_t(my));
      if (rc) return rc;
      part = ctx->workspace;
    }
    dim3 grid((unsigned)splits, (unsigned)my);
    if (n <= 8) k_gemm_tn<8><<<grid, 256, 0, ctx->stream>>>(M, n, K, A, lda, B, ldb, C, ldc, bias, kps, part, ctx->tickets);
    else        k_gemm_tn<16><<<grid, 256, 0, ctx->stream>>>(M, n, K, A, lda, B, ldb, C, ldc, bias, kps, part, ctx->tickets);
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  if (K <= 64 && N <= 32 && M >= 1024) {
    const int k = int(K);
    const int nt = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    const int pitch = k | 1, cp = n | 1;
    const size_t smem = size_t(256 * (pitch > cp ? pitch : cp) + k * nt + nt) * sizeof(float);
    const int64_t tiles = (M + 255) / 256, cap = int64_t(ctx->sm_count) * 8;
    const unsigned grid = unsigned(tiles < cap ? tiles : cap);
#define AGRS_ROWS(NT_)                                                                                                             \
  do {                                                                                                                              \
    AGRS_CUDA(ctx, cudaFuncSetAttribute(k_gemm_rows<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));                \
    k_gemm_rows<NT_><<<grid, 256, smem, ctx->stream>>>(M, n, k, A, lda, B, ldb, transB, C, ldc, bias);                               \
  } while (0)
    if (nt == 8) AGRS_ROWS(8);
    else if (nt == 16) AGRS_ROWS(16);
    else AGRS_ROWS(32);
#undef AGRS_ROWS
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  // dot with K sliced over CTAs: long K, 16-byte loadable rows of A
  if (K >= 512 && K % 4 == 0 && lda % 4 == 0 && agrs_aligned16(A) && M >= 64) {
    const int64_t steps = (K + 127) / 128;
    // about two CTAs per SM: row groups of 32 rows (two row pairs per warp), then as many K slices as that takes
    const int rows_per_cta = 32;
    const int64_t rgs = (M + rows_per_cta - 1) / rows_per_cta;
    int64_t slices = (2 * int64_t(ctx->sm_count) + rgs - 1) / rgs;
    if (slices > steps) slices = steps;
    if (slices < 1) slices = 1;
    const int sps = int((steps + slices - 1) / slices);
    slices = (steps + sps - 1) / sps;
    if (rgs <= 65535 && size_t(n) * size_t(sps * 128 + 4) * sizeof(float) <= 48 * 1024) {
      float* part = nullptr;
      if (slices > 1) {
        int rc = agrs_ensure_workspace(ctx, size_t(slices) * size_t(M) * size_t(n) * sizeof(float));
        if (rc) return rc;
        rc = ensure_tickets(ctx, size_t(rgs));
        if (rc) return rc;
        part = ctx->workspace;
      }
      const size_t smem = size_t(n) * size_t(sps * 128 + 4) * sizeof(float);
      dim3 grid((unsigned)slices, (unsigned)rgs);
      if (n <= 4)       k_gemm_dot_sliced<4><<<grid, 256, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, transB, C, ldc, bias, sps, rows_per_cta, part, ctx->tickets);
      else if (n <= 8)  k_gemm_dot_sliced<8><<<grid, 256, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, transB, C, ldc, bias, sps, rows_per_cta, part, ctx->tickets);
      else if (n <= 12) k_gemm_dot_sliced<12><<<grid, 256, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, transB, C, ldc, bias, sps, rows_per_cta, part, ctx->tickets);
      else              k_gemm_dot_sliced<16><<<grid, 256, smem, ctx->stream>>>(M, n, K, A, lda, B, ldb, transB, C, ldc, bias, sps, rows_per_cta, part, ctx->tickets);
      AGRS_LAUNCHED(ctx);
      return AGRS_OK;
    }
  }
  {
    const int64_t groups = (M + 3) / 4, cap = int64_t(ctx->sm_count) * 8;
    const unsigned grid = unsigned(groups < cap ? groups : cap);
    if (n <= 8) k_gemm_dot<8><<<grid, 128, 0, ctx->stream>>>(M, n, K, A, lda, B, ldb, transB, C, ldc, bias);
    else        k_gemm_dot<16><<<grid, 128, 0, ctx->stream>>>(M, n, K, A, lda, B, ldb, transB, C, ldc, bias);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}
