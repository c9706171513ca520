// Conv2D lowering kernels: im2col, col2im (gather form, no atomics), filter broadcast.
// Index maps follow src/operators/functional_ndarray.rs:40-138 of the reference exactly;
// both are pure data movement (HBM-bound).  Samples that fit in shared memory take the staged kernels (k_*_tile);
// the one-thread-per-output-element kernels remain for samples that do not.
#include <stdlib.h>

#include "agrs_internal.h"
#include "act.cuh"
#include "mag.cuh"

namespace {

// col[b, y*Wo + x, c*k*k + i*k + j] = im[b, c, y*s - p + i, x*s - p + j]  (0 outside the image)
__global__ void __launch_bounds__(256) k_im2col(const float* __restrict__ im, int64_t total, int C, int H, int W, int k, int s, int p,
                                                int Ho, int Wo, float* __restrict__ col) {
  const int kk = k * k, K = C * kk;
  const int64_t P = int64_t(Ho) * Wo;
  int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int q = int(idx % K);
    int64_t bp = idx / K;
    int pp = int(bp % P);
    int64_t b = bp / P;
    int c = q / kk, r = q % kk, i = r / k, j = r % k;
    int y = pp / Wo, x = pp % Wo;
    int h = y * s - p + i, w = x * s - p + j;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(im + ((b * C + c) * H + h) * int64_t(W) + w);
    col[idx] = v;
  }
}

// im[b,c,h,w] = sum over patches (y,x) and taps (i,j) with y*s+i == h, x*s+j == w of col[b, y*Wo+x, c*kk+i*k+j].
// Contributions are added in increasing patch index, the order the reference's overlap-add uses
// (functional_ndarray.rs:121-131), so the f32 result is reproducible and atomics-free.
__global__ void __launch_bounds__(256) k_col2im(const float* __restrict__ col, int64_t total, int C, int Hi, int Wi, int k, int s,
                                                int Ho, int Wo, float* __restrict__ im) {
  const int kk = k * k, K = C * kk;
  const int64_t P = int64_t(Ho) * Wo;
  int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int w = int(idx % Wi);
    int64_t t = idx / Wi;
    int h = int(t % Hi);
    t /= Hi;
    int c = int(t % C);
    int64_t b = t / C;
    float acc = 0.f;
    for (int i = k - 1; i >= 0; --i) {
      int hy = h - i;
      if (hy < 0 || hy % s) continue;
      int y = hy / s;
      if (y >= Ho) continue;
      for (int j = k - 1; j >= 0; --j) {
        int wx = w - j;
        if (wx < 0 || wx % s) continue;
        int x = wx / s;
        if (x >= Wo) continue;
        acc += __ldg(col + (b * P + int64_t(y) * Wo + x) * K + c * kk + i * k + j);
      }
    }
    im[idx] = acc;
  }
}

// ---- shared-memory staged versions: one CTA per image (grid-stride over the batch) --------------------------------
// im2col writes 18x (k*k x) more than it reads and col2im reads 18x more than it writes; what matters is that the big side
// moves as full 16-byte coalesced vectors with no per-element 64-bit div/mod.  The image (im2col) or the col block
// (col2im) of ONE sample is staged in shared memory; the scattered side of the index map is then a shared-memory gather
// (row pitch K = C*k*k floats: odd for the LeNet shape, so bank-conflict free).
// PAD = false: every tap is inside the image (pad == 0), no bounds checks.  Index arithmetic is incremental: a thread's
// consecutive float4 are kTileThreads*4 output elements apart, so (patch row y, patch column x, tap q) advance by constant
// steps with carries -- no division in the loop.
constexpr int kTileThreads = 512;

template <bool PAD>
__global__ void __launch_bounds__(kTileThreads) k_im2col_tile(const float* __restrict__ im, int B, int C, int H, int W, int k, int s, int p,
                                                              int Ho, int Wo, float* __restrict__ col, int vec_in, int vec_out) {
  extern __shared__ float sm[];
  const int n_im = C * H * W, kk = k * k, K = C * kk, P = Ho * Wo, per = P * K;
  int* trel = reinterpret_cast<int*>(sm + ((n_im + 3) & ~3));  // per tap q: offset of the tap relative to the patch origin
  int* tij = trel + K;                                         // (i << 16) | j, for the bounds checks of padded convolutions
  for (int q = threadIdx.x; q < K; q += blockDim.x) {
    const int c = q / kk, r = q - c * kk, i = r / k, j = r - i * k;
    trel[q] = c * H * W + i * W + j;
    tij[q] = (i << 16) | j;
  }
  // per-iteration step of a thread's first element: blockDim*4 elements = dP patches + dq taps; dP patches = dY rows + dX columns
  const int step = blockDim.x * 4;
  const int dP = step / K, dq = step - dP * K, dY = dP / Wo, dX = dP - dY * Wo;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();  // table built / previous sample fully consumed
    const float* src = im + int64_t(b) * n_im;
    if (vec_in) {
      for (int t = threadIdx.x; t < n_im / 4; t += blockDim.x) reinterpret_cast<float4*>(sm)[t] = __ldg(reinterpret_cast<const float4*>(src) + t);
    } else {
      for (int t = threadIdx.x; t < n_im; t += blockDim.x) sm[t] = __ldg(src + t);
    }
    __syncthreads();
    float* dst = col + int64_t(b) * per;
    if (vec_out) {
      const int o0 = 4 * threadIdx.x;
      int pp = o0 / K, q0 = o0 - pp * K;
      int y0 = pp / Wo, x0 = pp - y0 * Wo;
      for (int o4 = threadIdx.x; o4 < per / 4; o4 += blockDim.x) {
        int q = q0, x = x0, y = y0;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (PAD) {
            const int ij = tij[q], h = y * s - p + (ij >> 16), w = x * s - p + (ij & 0xFFFF);
            v[e] = (h >= 0 && h < H && w >= 0 && w < W) ? sm[(y * s - p) * W + (x * s - p) + trel[q]] : 0.f;
          } else {
            v[e] = sm[(y * W + x) * s + trel[q]];
          }
          if (++q == K) {
            q = 0;
            if (++x == Wo) { x = 0; ++y; }
          }
        }
        __stcs(reinterpret_cast<float4*>(dst) + o4, make_float4(v[0], v[1], v[2], v[3]));
        q0 += dq;
        int carry = q0 >= K;
        q0 -= carry ? K : 0;
        x0 += dX + carry;
        y0 += dY;
        while (x0 >= Wo) { x0 -= Wo; ++y0; }
      }
    } else {
      for (int o = threadIdx.x; o < per; o += blockDim.x) {
        const int pp = o / K, q = o - pp * K, y = pp / Wo, x = pp - y * Wo;
        const int ij = tij[q], h = y * s - p + (ij >> 16), w = x * s - p + (ij & 0xFFFF);
        dst[o] = (h >= 0 && h < H && w >= 0 && w < W) ? sm[(y * s - p) * W + (x * s - p) + trel[q]] : 0.f;
      }
    }
  }
}

// same summation order as k_col2im (increasing patch index), reading the sample's col block from shared memory.
// The block is staged with cp.async (16-byte LDGSTS, no register round trip).  NBUF = 2: while the CTA gathers sample i from
// one buffer, the cp.async copies of sample i+1 are already in flight into the other (one CTA of 1024 threads per SM);
// NBUF = 1 for col blocks too large to double-buffer.
template <int NBUF, int KS>  // KS: kernel size known at compile time (0 = run time): the stride-1 gather is then fully unrolled
__global__ void __launch_bounds__(NBUF >= 2 ? 1024 : kTileThreads) k_col2im_tile(const float* __restrict__ col, int B, int C, int Hi, int Wi, int k,
                                                                               int s, int Ho, int Wo, float* __restrict__ im, int vec_in) {
  extern __shared__ float sm_all[];
  const int kk = k * k, K = C * kk, P = Ho * Wo, per = P * K, n_im = C * Hi * Wi;
  const int per_pad = (per + 3) & ~3;
  const uint32_t sm_addr = static_cast<uint32_t>(__cvta_generic_to_shared(sm_all));
  auto stage = [&](int b, int buf) {  // start copying sample b's col block into buffer buf
    const float* src = col + int64_t(b) * per;
    if (vec_in) {
      for (int t = threadIdx.x; t < per / 4; t += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm_addr + 4u * (buf * per_pad) + 16u * t), "l"(src + 4 * t) : "memory");
    } else {
      float* dstb = sm_all + buf * per_pad;
      for (int t = threadIdx.x; t < per; t += blockDim.x) dstb[t] = __ldg(src + t);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto decode_taps = [&](int idx, int& base, uint32_t& mask) {
    const int w = idx % Wi, t = idx / Wi, h = t % Hi, c = t / Hi;
    base = (h * Wo + w) * K + c * kk;
    mask = 0;
    for (int i = 0; i < KS; ++i)
      for (int j = 0; j < KS; ++j)
        if (h - i >= 0 && h - i < Ho && w - j >= 0 && w - j < Wo) mask |= 1u << (i * KS + j);
  };
  int tap_base0 = 0;
  uint32_t tap_mask0 = 0;
  if (KS > 0 && s == 1 && int(threadIdx.x) < n_im) decode_taps(threadIdx.x, tap_base0, tap_mask0);
  int buf = 0;
  if (NBUF >= 2) {  // prologue: NBUF-1 samples in flight (empty groups keep the group count uniform)
#pragma unroll
    for (int j = 0; j < NBUF - 1; ++j) {
      const int64_t pb = int64_t(blockIdx.x) + int64_t(j) * gridDim.x;
      if (pb < B) stage(int(pb), j);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (NBUF >= 2) {
      const int64_t nb = int64_t(b) + int64_t(NBUF - 1) * gridDim.x;
      int nbuf = buf + NBUF - 1;
      nbuf -= nbuf >= NBUF ? NBUF : 0;
      if (nb < B) stage(int(nb), nbuf);  // NBUF-1 samples ahead: in flight during this gather
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(NBUF >= 2 ? NBUF - 1 : 0) : "memory");  // this sample has landed
    } else {
      __syncthreads();  // previous sample fully consumed
      stage(b, 0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* sm = sm_all + buf * per_pad;
    float* dst = im + int64_t(b) * n_im;
    if (KS > 0 && s == 1) {
      // Unit stride, kernel size known: which taps of an output pixel exist and where its first patch sits depend on (c, h, w)
      // only, so they were decoded once per thread (tap_mask0 / tap_base0) for the pixel the thread owns in the first round.
      // Every tap is then an independent predicated shared-memory load, all in flight at once, added in the reference's order.
      for (int idx = threadIdx.x, round = 0; idx < n_im; idx += blockDim.x, ++round) {
        int base = tap_base0;
        uint32_t mask = tap_mask0;
        if (round > 0) decode_taps(idx, base, mask);
        float acc = 0.f;
#pragma unroll
        for (int i = KS - 1; i >= 0; --i) {
#pragma unroll
          for (int j = KS - 1; j >= 0; --j) {
            const bool ok = (mask >> (i * KS + j)) & 1u;
            const float v = ok ? sm[base + i * KS + j - (i * Wo + j) * K] : 0.f;  // patch (h-i, w-j), tap (i, j)
            acc = ok ? acc + v : acc;
          }
        }
        __stcs(dst + idx, acc);
      }
    } else
    for (int idx = threadIdx.x; idx < n_im; idx += blockDim.x) {
      const int w = idx % Wi;
      int t = idx / Wi;
      const int h = t % Hi, c = t / Hi;
      float acc = 0.f;
      if (KS > 0 && s == 1) {  // handled by the hoisted path above
      } else if (s == 1) {  // unit stride: taps i in [h-Ho+1, h] and j in [w-Wo+1, w], clipped to the kernel
        const int i_hi = h < k - 1 ? h : k - 1, i_lo = h - (Ho - 1) > 0 ? h - (Ho - 1) : 0;
        const int j_hi = w < k - 1 ? w : k - 1, j_lo = w - (Wo - 1) > 0 ? w - (Wo - 1) : 0;
        for (int i = i_hi; i >= i_lo; --i) {
          const float* row = sm + ((h - i) * Wo + w) * K + c * kk + i * k;
          for (int j = j_hi; j >= j_lo; --j) acc += row[j - j * K];  // patch (h-i, w-j), tap (i, j)
        }
      } else {
        for (int i = k - 1; i >= 0; --i) {
          const int hy = h - i;
          if (hy < 0 || hy % s) continue;
          const int y = hy / s;
          if (y >= Ho) continue;
          for (int j = k - 1; j >= 0; --j) {
            const int wx = w - j;
            if (wx < 0 || wx % s) continue;
            const int x = wx / s;
            if (x >= Wo) continue;
            acc += sm[(y * Wo + x) * K + c * kk + i * k + j];
          }
        }
      }
      __stcs(dst + idx, acc);
    }
    if (NBUF >= 2) {
      __syncthreads();  // everybody is done with this buffer before a later iteration's copies overwrite it
      buf = buf + 1 == NBUF ? 0 : buf + 1;
    }
  }
}

constexpr size_t kTileSmemMax = 200 * 1024;  // leave room for a second CTA's static needs; larger samples take the generic kernels

__global__ void k_broadcast_filter(const float* __restrict__ filt, int64_t per, int64_t total, float* __restrict__ out) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) out[i] = filt[i % per];
}

unsigned grid1d(agrs_ctx* ctx, int64_t total) {
  int64_t blocks = (total + 255) / 256, cap = int64_t(ctx->sm_count) * 16;
  return unsigned(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}


// ---------------------------------------------------------------------------------------------------------------------------
// Implicit-GEMM Conv2D for single-channel inputs (the LeNet layer of BASELINE configs[2]): the col matrix of
// operators.rs:350-390 -- 18x the image for k = 5 -- is never materialised.  With C_in = 1 the lowered GEMMs are
//   forward   out[b, p, o] = sum_t col[b, p, t] * filt[t, o] + bias[o]          K = k*k (25), N = C_out (6)
//   dxcol     g[b, p, :] . filt[t, :]  followed by the overlap-add of col2im      (functional_ndarray.rs:107-138)
//   dw        sum_{b, p} col[b, p, t] * g[b, p, o],   db = sum_{b, p} g[b, p, o]  (operators.rs:405, :445)
// and col[b, p, t] is im[b, 0, y*s - pad + i, x*s - pad + j]: every kernel below reads it straight from the image of one sample
// staged (zero-padded) in shared memory.  HBM traffic: image + output forward; image + gradient + dx backward.
// One CTA walks samples b = blockIdx.x, blockIdx.x + gridDim.x, ..; all sums run in a fixed order (deterministic).
constexpr int kConvCo = 8;  // C_out handled per thread (register accumulators); channel vectors are padded to 8 in shared memory

// 8 floats (one position's C_out values, or one tap's filter row) from shared memory: two 16-byte loads
struct F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ld8(const float* p) { return {*reinterpret_cast<const float4*>(p), *reinterpret_cast<const float4*>(p + 4)}; }
__device__ __forceinline__ float dot8(const F8& a, const F8& b) {  // sum over C_out in channel order (padding channels are zero)
  float t = a.lo.x * b.lo.x;
  t = fmaf(a.lo.y, b.lo.y, t); t = fmaf(a.lo.z, b.lo.z, t); t = fmaf(a.lo.w, b.lo.w, t);
  t = fmaf(a.hi.x, b.hi.x, t); t = fmaf(a.hi.y, b.hi.y, t); t = fmaf(a.hi.z, b.hi.z, t); t = fmaf(a.hi.w, b.hi.w, t);
  return t;
}

// The forward kernels stage a group's output [p][co] in shared memory and stream it out; with an activation requested
// (agrs_conv2d_fwd_act_f32: Conv2D feeding ReLU / Sigmoid) the same stream also writes y = act(out) -- the
// pre-activation stays a tensor of its own, the activation's graph node keeps it -- and tracks the magnitude of y.
struct ConvAct {
  int kind;      // AGRS_ACT_NONE: plain forward
  float* y;      // same layout as out
  uint32_t* mx;  // magnitude word of y, or null
};
__device__ __forceinline__ float conv_act1(int kind, float v) { return kind == AGRS_ACT_RELU ? agrs_act_relu(v) : agrs_act_sigmoid(v); }
__device__ __forceinline__ void conv_stream_out(const float* stage, float* dst, int64_t off, int n, const ConvAct& act, uint32_t& mag, int tid, int nthreads) {
  float* const d = dst + off;
  float* const y = act.kind ? act.y + off : nullptr;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) && (!y || (reinterpret_cast<uintptr_t>(y) & 15u) == 0)) {
    for (int e = tid; e < n >> 2; e += nthreads) {
      const float4 v = reinterpret_cast<const float4*>(stage)[e];
      __stcs(reinterpret_cast<float4*>(d) + e, v);
      if (y) {
        const float4 r = make_float4(conv_act1(act.kind, v.x), conv_act1(act.kind, v.y), conv_act1(act.kind, v.z), conv_act1(act.kind, v.w));
        __stcs(reinterpret_cast<float4*>(y) + e, r);
        mag = mag4(mag, r);
      }
    }
  } else {
    for (int e = tid; e < n; e += nthreads) {
      const float v = stage[e];
      d[e] = v;
      if (y) {
        const float r = conv_act1(act.kind, v);
        y[e] = r;
        mag = max(mag, mag_bits(r));
      }
    }
  }
}

// forward: thread per output position, the CTA's results staged in shared memory and written as one coalesced stream
__global__ void __launch_bounds__(256) k_conv1_fwd(const float* __restrict__ im, const float* __restrict__ filt, const float* __restrict__ bias,
                                                   int B, int H, int W, int k, int co, int s, int pad, int Ho, int Wo, float* __restrict__ out,
                                                   const ConvAct act) {
  extern __shared__ __align__(16) float sm[];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, kk = k * k, P = Ho * Wo;
  uint32_t y_mag = 0;
  float* fl = sm;                          // [kk][8], zero padded
  float* bs = fl + kk * 8;                 // [8]
  float* img = bs + 8;                     // [Hp][Wp], zero border
  float* stage = img + ((Hp * Wp + 3) & ~3);  // [P][co]: the sample's output exactly as it lies in global memory
  for (int e = threadIdx.x; e < kk * 8 + 8; e += blockDim.x) {
    const int t = e >> 3, o = e & 7;
    fl[e] = o < co ? (t < kk ? __ldg(filt + t * co + o) : __ldg(bias + o)) : 0.f;
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float* src = im + int64_t(b) * H * W;
    if (pad == 0) {
      for (int e = threadIdx.x; e < H * W; e += blockDim.x) img[e] = __ldcs(src + e);
    } else {
      for (int e = threadIdx.x; e < Hp * Wp; e += blockDim.x) {
        const int h = e / Wp - pad, w = e % Wp - pad;
        img[e] = (h >= 0 && h < H && w >= 0 && w < W) ? __ldcs(src + h * W + w) : 0.f;
      }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      const int y = p / Wo, x = p - y * Wo;
      float acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = 0.f;
      const float* base = img + (y * s) * Wp + x * s;
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {  // t = i*k + j ascending: the GEMM's order over K
          const float v = base[i * Wp + j];
          const F8 f = ld8(fl + (i * k + j) * 8);  // the same address for the whole warp: a broadcast
          acc[0] = fmaf(v, f.lo.x, acc[0]); acc[1] = fmaf(v, f.lo.y, acc[1]); acc[2] = fmaf(v, f.lo.z, acc[2]); acc[3] = fmaf(v, f.lo.w, acc[3]);
          acc[4] = fmaf(v, f.hi.x, acc[4]); acc[5] = fmaf(v, f.hi.y, acc[5]); acc[6] = fmaf(v, f.hi.z, acc[6]); acc[7] = fmaf(v, f.hi.w, acc[7]);
        }
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o < co) stage[p * co + o] = acc[o] + bs[o];
    }
    __syncthreads();
    // NHWC (operators.rs:383): [P][co] contiguous per sample
    conv_stream_out(stage, out, int64_t(b) * P * co, P * co, act, y_mag, threadIdx.x, blockDim.x);
  }
  if (act.kind && act.mx) mag_commit(y_mag, act.mx);
}

// backward of the same layer: dx (padded size, per sample), and this CTA's partial sums of dw [kk][co] and db [co]; the last CTA
// to finish folds the partials in CTA order.  dw / db: thread = (tap, half of the channels, slice of the positions), accumulators
// live in registers across all the samples of the CTA; dx: thread per pixel, channel vectors of 8 from shared memory.  The
// gradient tile is staged with a zero border of k-1 positions, so for stride 1 the dx loops carry no bounds checks (profiling the
// first version showed 7 integer / branch instructions per FFMA: profiles/r2_c3_timeline.md).  KT = compile-time k (0: run-time).
template <bool S1, int KT>
__global__ void __launch_bounds__(256) k_conv1_bwd(const float* __restrict__ im, const float* __restrict__ filt, const float* __restrict__ g,
                                                   int B, int H, int W, int k_rt, int co, int s, int pad, int Ho, int Wo, float* __restrict__ dx,
                                                   float* __restrict__ dw, float* __restrict__ db, float* __restrict__ part,
                                                   unsigned int* __restrict__ ticket) {
  extern __shared__ __align__(16) float sm[];
  __shared__ unsigned int s_ticket;
  const int k = KT ? KT : k_rt;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, kk = k * k, P = Ho * Wo;
  const int hx = (Ho - 1) * s + k, wx = (Wo - 1) * s + k;  // dx keeps the padded size (TODO at operators.rs:427)
  const int bd = k - 1, Wg = Wo + 2 * bd, Hg = Ho + 2 * bd;  // gradient tile with its zero border
  float* fl = sm;                           // [kk][8]
  float* gt = fl + kk * 8;                  // [Hg][Wg][8]
  float* img = gt + Hg * Wg * 8;            // [Hp][Wp]
  float* red = img + ((Hp * Wp + 3) & ~3);  // [slices][kk + 1][8]: the CTA's dw / db partials, folded over the position slices at the end
  for (int e = threadIdx.x; e < kk * 8; e += blockDim.x) fl[e] = (e & 7) < co ? __ldg(filt + (e >> 3) * co + (e & 7)) : 0.f;
  for (int e = threadIdx.x; e < Hg * Wg * 8; e += blockDim.x) gt[e] = 0.f;  // the border (and the padding channels) stay zero
  // dw / db work split: unit u = threadIdx.x covers tap (or the db row) r = u % (kk + 1), channel half h = (u / (kk + 1)) & 1, and
  // position slice sl = u / (2 * (kk + 1)); slices = blockDim.x / (2 * (kk + 1)) >= 1 (k*k*Co + Co <= 256 is checked by the host)
  const int rows = kk + 1, slices = int(blockDim.x) / (2 * rows);
  const int u = threadIdx.x, r = u % rows, hf = (u / rows) & 1, sl = u / (2 * rows);
  const bool worker = sl < slices;
  const int ti = r < kk ? r / k : 0, tj = r < kk ? r % k : 0;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float* src = im + int64_t(b) * H * W;
    if (pad == 0) {
      for (int e = threadIdx.x; e < H * W; e += blockDim.x) img[e] = __ldcs(src + e);
    } else {
      for (int e = threadIdx.x; e < Hp * Wp; e += blockDim.x) {
        const int h = e / Wp - pad, w = e % Wp - pad;
        img[e] = (h >= 0 && h < H && w >= 0 && w < W) ? __ldcs(src + h * W + w) : 0.f;
      }
    }
    const float* gb = g + int64_t(b) * P * co;
    {
      int p = int(threadIdx.x) / co, o = int(threadIdx.x) - p * co;   // element e = p*co + o of the sample's gradient, no division per element
      int y = p / Wo, x = p - y * Wo;
      const int dp = int(blockDim.x) / co, d_o = int(blockDim.x) - dp * co;
      for (int e = threadIdx.x; e < P * co; e += blockDim.x) {
        gt[((y + bd) * Wg + x + bd) * 8 + o] = __ldcs(gb + e);
        o += d_o; x += dp;
        if (o >= co) { o -= co; ++x; }
        while (x >= Wo) { x -= Wo; ++y; }
      }
    }
    __syncthreads();
    if (worker) {  // positions p = sl, sl + slices, ..: ascending inside the slice
      int y = sl / Wo, x = sl - y * Wo;  // (y, x) of position p, advanced without a division per position
      const float* tap = img + ti * Wp + tj;
      const float* gp = gt + (bd * Wg + bd) * 8 + hf * 4;
      for (int p = sl; p < P; p += slices) {
        const float v = r < kk ? tap[y * s * Wp + x * s] : 1.0f;  // the db row sums g itself
        const float4 gv = *reinterpret_cast<const float4*>(gp + (y * Wg + x) * 8);
        a0 = fmaf(v, gv.x, a0); a1 = fmaf(v, gv.y, a1); a2 = fmaf(v, gv.z, a2); a3 = fmaf(v, gv.w, a3);
        x += slices;
        while (x >= Wo) { x -= Wo; ++y; }
      }
    }
    // dx[b, 0, u, v] = sum over positions (y, x) ascending and taps with y*s + i == u, x*s + j == v of g[y, x, :] . filt[i*k + j, :]
    int uu = int(threadIdx.x) / wx, vv = int(threadIdx.x) - uu * wx;
    for (int e = threadIdx.x; e < hx * wx; e += blockDim.x) {
      float a = 0.f;
      if (S1) {
        // stride 1: y = u - i, x = v - j; in the bordered tile every (i, j) is in range, the border contributes exact zeros
        const float* gq = gt + (uu * Wg + vv) * 8;  // tile position (u - i + bd, v - j + bd) for i = j = bd
#pragma unroll
        for (int i = k - 1; i >= 0; --i) {          // i descending = y ascending
#pragma unroll
          for (int j = k - 1; j >= 0; --j)
            a += dot8(ld8(gq + ((bd - i) * Wg + (bd - j)) * 8), ld8(fl + (i * k + j) * 8));  // dxcol[p, t], then col2im's overlap-add
        }
      } else {
        for (int i = k - 1; i >= 0; --i) {
          int y = uu - i;
          if (y < 0 || y % s) continue;
          y /= s;
          if (y >= Ho) continue;
          for (int j = k - 1; j >= 0; --j) {
            int x = vv - j;
            if (x < 0 || x % s) continue;
            x /= s;
            if (x >= Wo) continue;
            a += dot8(ld8(gt + ((y + bd) * Wg + x + bd) * 8), ld8(fl + (i * k + j) * 8));
          }
        }
      }
      __stcs(dx + int64_t(b) * hx * wx + e, a);
      vv += blockDim.x;
      while (vv >= wx) { vv -= wx; ++uu; }
    }
  }
  // fold the position slices (slice order), then hand this CTA's kk*co + co sums to the cross-CTA fold
  __syncthreads();
  if (worker) *reinterpret_cast<float4*>(red + (sl * rows + r) * 8 + hf * 4) = make_float4(a0, a1, a2, a3);
  __syncthreads();
  const int nacc = kk * co + co;
  if (int(threadIdx.x) < nacc) {
    const int rr = threadIdx.x / co, o = threadIdx.x % co;  // rows 0 .. kk-1: dw, row kk: db
    float t = 0.f;
    for (int q = 0; q < slices; ++q) t += red[(q * rows + rr) * 8 + o];
    part[int64_t(blockIdx.x) * nacc + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  // Cross-CTA fold by the last CTA: thread o adds the partials of output o in CTA order; the loads go out in independent
  // batches of 8 (a one-at-a-time chain of L2 round trips costs tens of microseconds for ~600 CTAs), the adds stay in order.
  if (int(threadIdx.x) < nacc) {
    const float* src = part + threadIdx.x;
    float t = 0.f;
    unsigned c = 0;
    for (; c + 8 <= gridDim.x; c += 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __ldcg(src + int64_t(c + q) * nacc);
#pragma unroll
      for (int q = 0; q < 8; ++q) t += v[q];
    }
    for (; c < gridDim.x; ++c) t += __ldcg(src + int64_t(c) * nacc);
    if (int(threadIdx.x) < kk * co) dw[threadIdx.x] = t;  // [k][k][co] == [kk][co]
    else                            db[threadIdx.x - kk * co] = t;
  }
  if (threadIdx.x == 0) *ticket = 0;
}

// ---- register-tiled backward for the common small-image case (stride 1, no padding, k <= 5, H, W <= 32: the LeNet layer) ----
// The kernel above feeds every FMA from shared memory (one or two loads per multiply-add): at the C3 size it runs at 119 us for
// 0.2 GFMA, shared-memory bound (profiles/r2_c3_timeline.md).  Here every thread keeps a tile of results in registers and each
// shared-memory vector feeds 10-17 multiply-adds:
//   * the gradient of a sample is staged channel-PLANAR with a zero border, gb[o][y + K-1][x + 4], rows `gs` floats apart;
//   * dx: thread = 4 x 4 tile of the input gradient.  For one channel it holds the K*K filter in registers and walks the
//     (4 + K-1) x 8 window of gb above its tile: 2 vector loads per window row, 4 * K * K multiply-adds per output row;
//   * dw / db: warp = (channel, filter row i), lane = output row y.  A lane slides along its row four positions at a time
//     (one vector of g, one new vector of the image row y + i) and keeps the K taps dw[i][0..K-1][o] in registers, across all
//     the samples of the CTA; the lanes are summed once, at the end.
// Row strides are multiples of 4 floats whose quarter is odd: 8 lanes that differ in y (or in the tile column) then hit 8 distinct
// 16-byte bank groups.  Four samples are resident per round (64 threads each for dx).  Sums run in a fixed order: deterministic.
constexpr int kCtSamples = 4, kCtThreads = 256, kCtColBorder = 4;
__host__ __device__ inline int ct_stride(int need) {  // smallest multiple of 4 >= need whose quarter is odd
  int s = (need + 3) & ~3;
  return ((s >> 2) & 1) ? s : s + 4;
}
struct CtGeom {
  int gs, gr, gcs, is, img_sz;  // gb row stride / rows / channel stride; image row stride / floats per sample
  int raw_g, raw_i;             // floats of the raw landing buffers
  size_t smem;
};
__host__ __device__ inline CtGeom ct_geom(int H, int W, int K, int co) {
  CtGeom g;
  const int hx = H, wx = W;  // stride 1, no padding: the input gradient has the size of the image
  g.gs = ct_stride(((wx + 3) / 4) * 4 + kCtColBorder);
  g.gr = ((hx + 3) / 4) * 4 + (K - 1);
  g.gcs = g.gr * g.gs + 4;   // + 4: the channels of one position spread over different banks while staging
  g.is = ct_stride(W + 4);    // 4 zero columns after every image row: a window that slides past the row's end multiplies zeros by
                              // the gradient tile's zero border, never a neighbouring row's value (which could be inf) by zero
  g.img_sz = (H + 1) * g.is;
  const int Ho = H - K + 1, Wo = W - K + 1;
  g.raw_g = kCtSamples * Ho * Wo * co;  // the group's gradients and images exactly as they lie in global memory (bulk-copied)
  g.raw_i = kCtSamples * H * W;
  g.smem = (size_t(8) * 28 + size_t(kCtSamples) * (size_t(co) * g.gcs + g.img_sz) + size_t(g.raw_g) + size_t(g.raw_i)) * sizeof(float) + 16;
  return g;
}

__device__ __forceinline__ uint32_t ct_smem(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ct_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {  // contiguous global -> shared, async
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ct_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <int K>
__global__ void __launch_bounds__(kCtThreads, 1) k_conv1_bwd_tiled(const float* __restrict__ im, const float* __restrict__ filt, const float* __restrict__ g,
                                                                   int B, int H, int W, int co, float* __restrict__ dx, float* __restrict__ dw,
                                                                   float* __restrict__ db, float* __restrict__ part, unsigned int* __restrict__ ticket) {
  extern __shared__ __align__(16) float sm[];
  __shared__ unsigned int s_ticket;
  constexpr int BD = K - 1, KK = K * K, FP = 28;  // filter of one channel padded to 28 floats (K <= 5)
  const int Ho = H - K + 1, Wo = W - K + 1, P = Ho * Wo;
  const CtGeom ge = ct_geom(H, W, K, co);
  float* fl = sm;                                   // [8][FP]
  float* gb = fl + 8 * FP;                          // [S][co][gr][gs] (+4 per channel)
  float* img = gb + kCtSamples * co * ge.gcs;       // [S][H + 1][is]
  float* rawg = img + kCtSamples * ge.img_sz;       // [S][P][co]: landing buffer of the bulk copies (the next group's data arrives
  float* rawi = rawg + ge.raw_g;                    // [S][H][W]    while this group is being computed)
  const uint32_t bar = ct_smem(rawi + ge.raw_i);    // one mbarrier, 8-byte aligned
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int groups = (B + kCtSamples - 1) / kCtSamples;
  auto fetch = [&](int grp) {  // thread 0: both blocks of group grp, completion counted in bytes on the barrier
    const int b0 = grp * kCtSamples, ns = B - b0 < kCtSamples ? B - b0 : kCtSamples;
    const uint32_t bg = uint32_t(ns) * uint32_t(P * co) * 4u, bi = uint32_t(ns) * uint32_t(H * W) * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bg + bi) : "memory");
    ct_bulk_load(ct_smem(rawg), g + int64_t(b0) * P * co, bg, bar);
    ct_bulk_load(ct_smem(rawi), im + int64_t(b0) * H * W, bi, bar);
  };
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (int(blockIdx.x) < groups) fetch(blockIdx.x);
  }
  uint32_t phase = 0;
  for (int e = tid; e < 8 * FP; e += kCtThreads) {
    const int o = e / FP, t = e - o * FP;
    fl[e] = (o < co && t < KK) ? __ldg(filt + t * co + o) : 0.f;
  }
  for (int e = tid; e < kCtSamples * (co * ge.gcs + ge.img_sz); e += kCtThreads) gb[e] = 0.f;  // borders, spare rows: zero for good

  // dw / db: this warp's (channel, filter row) pairs c = warp, warp + 8, ..  (o = c / K, i = c % K); accumulators live across samples
  constexpr int kMaxPairs = (8 * K + 7) / 8;  // co <= 8
  float acc[kMaxPairs][K], accb[kMaxPairs];
#pragma unroll
  for (int q = 0; q < kMaxPairs; ++q) {
    accb[q] = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc[q][j] = 0.f;
  }
  // dx: slot = one of the resident samples, tile (tu, tv) of 4 x 4 input-gradient pixels
  const int slot = tid >> 6, tt = tid & 63, tu = tt >> 3, tv = tt & 7;
  const bool dx_thread = 4 * tu < H && 4 * tv < W;

  __syncthreads();  // the barrier exists before anybody polls it
  // the walk of a sample's gradient, e = p * co + o -> (o, y, x), starts at the same place for every sample: found once
  const int p_first = tid / co, o_first = tid - p_first * co, y_first = p_first / Wo, x_first = p_first - y_first * Wo;
  const int dp = kCtThreads / co, d_o = kCtThreads - dp * co;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int b0 = grp * kCtSamples;
    const int ns = B - b0 < kCtSamples ? B - b0 : kCtSamples;
    ct_wait(bar, phase);  // this group's raw blocks have landed
    phase ^= 1u;
    __syncthreads();      // the previous round's readers of gb / img are done (and the zero fill, the first time)
    // ---- shared -> shared: image rows to their padded pitch, the gradient NHWC -> planar with border ----
    for (int sidx = 0; sidx < ns; ++sidx) {
      const float* src = rawi + sidx * H * W;
      float* dsti = img + sidx * ge.img_sz;
      for (int e = tid; e < (H * W) >> 2; e += kCtThreads) {  // W % 4 == 0: a vector never straddles two rows
        const int row = (4 * e) / W, col = 4 * e - row * W;
        *reinterpret_cast<float4*>(dsti + row * ge.is + col) = reinterpret_cast<const float4*>(src)[e];
      }
      const float* gsrc = rawg + sidx * P * co;
      float* dstg = gb + sidx * co * ge.gcs;
      int o = o_first, y = y_first, x = x_first;
      for (int e = tid; e < P * co; e += kCtThreads) {
        dstg[o * ge.gcs + (y + BD) * ge.gs + x + kCtColBorder] = gsrc[e];
        o += d_o; x += dp;
        if (o >= co) { o -= co; ++x; }
        while (x >= Wo) { x -= Wo; ++y; }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my reads of the landing buffers come before the next bulk write
    __syncthreads();
    if (tid == 0 && grp + int(gridDim.x) < groups) fetch(grp + gridDim.x);  // lands while this group is computed
    // ---- dx ----
    if (slot < ns && dx_thread) {
      float a[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[r][c] = 0.f;
      const float* gp0 = gb + slot * co * ge.gcs + (4 * tu) * ge.gs + 4 * tv;
      for (int o = 0; o < co; ++o) {
        float f[FP];
#pragma unroll
        for (int q = 0; q < FP / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(fl + o * FP + 4 * q);  // one address per warp: broadcast
          f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
        }
        const float* gp = gp0 + o * ge.gcs;
#pragma unroll
        for (int wr = 0; wr < 4 + BD; ++wr) {  // window row wr holds gradient row 4*tu + wr - BD
          const float4 w0 = *reinterpret_cast<const float4*>(gp + wr * ge.gs), w1 = *reinterpret_cast<const float4*>(gp + wr * ge.gs + 4);
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = r + BD - wr;  // filter row that maps window row wr onto output row r
            if (i < 0 || i >= K) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int j = 0; j < K; ++j) a[r][c] = fmaf(w[c + kCtColBorder - j], f[i * K + j], a[r][c]);
          }
        }
      }
      float* dst = dx + int64_t(b0 + slot) * H * W;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int u = 4 * tu + r;
        if (u < H) __stcs(reinterpret_cast<float4*>(dst + u * W + 4 * tv), make_float4(a[r][0], a[r][1], a[r][2], a[r][3]));  // W % 4 == 0
      }
    }
    // ---- dw / db ----
    if (lane < Ho) {
#pragma unroll
      for (int q = 0; q < kMaxPairs; ++q) {
        const int c = warp + 8 * q;
        if (c >= co * K) break;
        const int o = c / K, i = c - o * K;
        for (int sidx = 0; sidx < ns; ++sidx) {
          const float* ip = img + sidx * ge.img_sz + (lane + i) * ge.is;
          const float* gp = gb + sidx * co * ge.gcs + o * ge.gcs + (lane + BD) * ge.gs + kCtColBorder;
          float4 prev = *reinterpret_cast<const float4*>(ip);
          for (int x = 0; x < Wo; x += 4) {  // past Wo the gradient vector holds border zeros, the image vector padding zeros
            const float4 gv = *reinterpret_cast<const float4*>(gp + x);
            const float4 nxt = *reinterpret_cast<const float4*>(ip + x + 4);
            const float w[8] = {prev.x, prev.y, prev.z, prev.w, nxt.x, nxt.y, nxt.z, nxt.w};
            const float gq[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int xx = 0; xx < 4; ++xx)
#pragma unroll
              for (int j = 0; j < K; ++j) acc[q][j] = fmaf(w[xx + j], gq[xx], acc[q][j]);
            if (i == 0) accb[q] += (gq[0] + gq[1]) + (gq[2] + gq[3]);
            prev = nxt;
          }
        }
      }
    }
  }
  // ---- this CTA's sums: lanes (output rows) folded in a fixed tree, then [kk][co] dw followed by [co] db ----
  const int nacc = KK * co + co;
#pragma unroll
  for (int q = 0; q < kMaxPairs; ++q) {
    const int c = warp + 8 * q;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float v = acc[q][j];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, sft);
      acc[q][j] = v;
    }
    float vb = accb[q];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) vb += __shfl_xor_sync(0xFFFFFFFFu, vb, sft);
    if (lane == 0 && c < co * K) {
      const int o = c / K, i = c - o * K;
#pragma unroll
      for (int j = 0; j < K; ++j) part[int64_t(blockIdx.x) * nacc + (i * K + j) * co + o] = acc[q][j];
      if (i == 0) part[int64_t(blockIdx.x) * nacc + KK * co + o] = vb;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  if (tid < nacc) {  // last CTA: partials added in CTA order, loads batched 8 at a time
    const float* src = part + tid;
    float t = 0.f;
    unsigned c = 0;
    for (; c + 8 <= gridDim.x; c += 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __ldcg(src + int64_t(c + q) * nacc);
#pragma unroll
      for (int q = 0; q < 8; ++q) t += v[q];
    }
    for (; c < gridDim.x; ++c) t += __ldcg(src + int64_t(c) * nacc);
    if (tid < KK * co) dw[tid] = t;
    else               db[tid - KK * co] = t;
  }
  if (tid == 0) *ticket = 0;
}

// ---- register-tiled forward for the same small-image case ----
// Thread = four neighbouring output positions of one row x all (<= 8, zero-padded) channels: 32 accumulators.  Per filter row it
// reads 8 image values of row y + i with two 16-byte loads and each tap's 8 channel weights as two broadcast loads: 160 multiply-adds
// per 12 loads (k = 5), where k_conv1_fwd does 8 per 3.  Four samples per CTA round land by one bulk copy (the next group's under
// this group's arithmetic); the results are staged [p][co] in shared memory and leave as one coalesced 16-byte stream (NHWC).
constexpr int kCfSamples = 4, kCfThreads = 256;
__host__ __device__ inline size_t cf_smem(int H, int W, int K, int co) {
  const int Ho = H - K + 1, Wo = W - K + 1;
  return (size_t(K * K + 1) * 8 + size_t(kCfSamples) * H * W + 8 + size_t(kCfSamples) * Ho * Wo * co) * sizeof(float) + 16;
}
template <int K>
__global__ void __launch_bounds__(kCfThreads) k_conv1_fwd_tiled(const float* __restrict__ im, const float* __restrict__ filt, const float* __restrict__ bias,
                                                                int B, int H, int W, int co, float* __restrict__ out, const ConvAct act) {
  extern __shared__ __align__(16) float sm[];
  constexpr int KK = K * K;
  uint32_t y_mag = 0;
  const int Ho = H - K + 1, Wo = W - K + 1, P = Ho * Wo, strips_x = (Wo + 3) >> 2, strips = Ho * strips_x;
  float* fl = sm;                                   // [KK + 1][8]: taps, then the bias row
  float* raw = fl + (KK + 1) * 8;                   // [S][H][W] as in global memory (+ 8 floats a window may look past the end)
  float* stage = raw + kCfSamples * H * W + 8;      // [S][P][co]
  const uint32_t bar = ct_smem(stage + kCfSamples * P * co);
  const int tid = threadIdx.x;
  const int groups = (B + kCfSamples - 1) / kCfSamples;
  auto fetch = [&](int grp) {
    const int b0 = grp * kCfSamples, ns = B - b0 < kCfSamples ? B - b0 : kCfSamples;
    const uint32_t bytes = uint32_t(ns) * uint32_t(H * W) * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    ct_bulk_load(ct_smem(raw), im + int64_t(b0) * H * W, bytes, bar);
  };
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (int(blockIdx.x) < groups) fetch(blockIdx.x);
  }
  for (int e = tid; e < (KK + 1) * 8; e += kCfThreads) {
    const int t = e >> 3, o = e & 7;
    fl[e] = o < co ? (t < KK ? __ldg(filt + t * co + o) : __ldg(bias + o)) : 0.f;
  }
  if (tid < 8) raw[kCfSamples * H * W + tid] = 0.f;
  __syncthreads();
  uint32_t phase = 0;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int b0 = grp * kCfSamples, ns = B - b0 < kCfSamples ? B - b0 : kCfSamples;
    ct_wait(bar, phase);
    phase ^= 1u;
    for (int item = tid; item < ns * strips; item += kCfThreads) {
      const int sidx = item / strips, rem = item - sidx * strips, y = rem / strips_x, x0 = (rem - y * strips_x) << 2;
      float acc[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[q][o] = 0.f;
      const float* ip = raw + sidx * H * W + y * W + x0;   // W % 4 == 0: 16-byte aligned
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float4 w0 = *reinterpret_cast<const float4*>(ip + i * W), w1 = *reinterpret_cast<const float4*>(ip + i * W + 4);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < K; ++j) {  // t = i*K + j ascending: the GEMM's order over K
          const float4 f0 = *reinterpret_cast<const float4*>(fl + (i * K + j) * 8), f1 = *reinterpret_cast<const float4*>(fl + (i * K + j) * 8 + 4);
          const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[q][o] = fmaf(w[q + j], f[o], acc[q][o]);
        }
      }
      float* st = stage + (sidx * P + y * Wo + x0) * co;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (x0 + q < Wo) {
#pragma unroll
          for (int o = 0; o < 8; ++o)
            if (o < co) st[q * co + o] = acc[q][o] + fl[KK * 8 + o];  // taps in ascending order, then the bias: k_conv1_fwd's bits
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();  // results staged; everybody is done reading the landing buffer
    if (tid == 0 && grp + int(gridDim.x) < groups) fetch(grp + gridDim.x);
    // the group's samples are contiguous in NHWC
    conv_stream_out(stage, out, int64_t(b0) * P * co, ns * P * co, act, y_mag, tid, kCfThreads);
    __syncthreads();  // the staging area is free for the next group
  }
  if (act.kind && act.mx) mag_commit(y_mag, act.mx);
}

static bool conv1_fwd_tiled_ok(int64_t B, int64_t H, int64_t W, int64_t k, int64_t co, int64_t stride, int64_t pad, const float* im) {
  if (stride != 1 || pad != 0 || (k != 3 && k != 5) || co < 1 || co > 8 || H < k || W < k || B < 1 || W % 4 != 0 || H > 64 || W > 64) return false;
  if (!agrs_aligned16(im)) return false;
  return cf_smem(int(H), int(W), int(k), int(co)) <= 100 * 1024;
}

static bool conv1_tiled_ok(int64_t B, int64_t H, int64_t W, int64_t k, int64_t co, int64_t stride, int64_t pad, const float* im, const float* g,
                           const float* dx) {
  if (stride != 1 || pad != 0 || (k != 3 && k != 5) || co < 1 || co > 8 || H > 32 || W > 32 || H < k || W < k || B < 1) return false;
  if (W % 4 != 0) return false;  // image rows staged and dx rows stored as whole 16-byte vectors
  if (k * k * co + co > kCtThreads) return false;
  if (!agrs_aligned16(im) || !agrs_aligned16(g) || !agrs_aligned16(dx)) return false;
  if (((H - k + 1) * (W - k + 1) * co) % 4 != 0) return false;  // every sample's gradient block starts 16-byte aligned (bulk copies)
  return ct_geom(int(H), int(W), int(k), int(co)).smem <= 220 * 1024;
}

}  // namespace

extern "C" {

int agrs_im2col_f32(agrs_ctx* ctx, const float* im, int64_t B, int64_t C, int64_t H, int64_t W, int64_t k, int64_t stride,
                    int64_t pad, float* col) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_im2col_f32");
  AGRS_REQUIRE(ctx, B >= 0 && C >= 0 && H >= 0 && W >= 0 && k > 0 && stride > 0 && pad >= 0, AGRS_ERR_INVALID, "agrs_im2col_f32: bad geometry");
  AGRS_REQUIRE(ctx, H + 2 * pad >= k && W + 2 * pad >= k, AGRS_ERR_SHAPE, "agrs_im2col_f32: kernel %lld larger than padded image %lldx%lld",
               (long long)k, (long long)(H + 2 * pad), (long long)(W + 2 * pad));
  int64_t Ho = (H - k + 2 * pad) / stride + 1, Wo = (W - k + 2 * pad) / stride + 1;
  int64_t total = B * Ho * Wo * C * k * k;
  if (total == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, im && col, AGRS_ERR_INVALID, "agrs_im2col_f32: null pointer");
  AGRS_REQUIRE(ctx, C * k * k < (1LL << 31) && Ho * Wo < (1LL << 31), AGRS_ERR_UNSUPPORTED, "agrs_im2col_f32: extent overflow");
  const int64_t n_im = C * H * W, per = Ho * Wo * C * k * k;
  const size_t smem = size_t(((n_im + 3) & ~int64_t(3)) + 2 * C * k * k) * sizeof(float);
  if (smem <= kTileSmemMax && per < (1LL << 30) && B < (1LL << 31)) {
    const int vec_in = (n_im % 4 == 0) && agrs_aligned16(im), vec_out = (per % 4 == 0) && agrs_aligned16(col);
    const int64_t cap = int64_t(ctx->sm_count) * 4;
    const unsigned grid = unsigned(B < cap ? B : cap);
    if (pad > 0) {
      AGRS_CUDA(ctx, cudaFuncSetAttribute(k_im2col_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTileSmemMax)));
      k_im2col_tile<true><<<grid, kTileThreads, smem, ctx->stream>>>(im, int(B), int(C), int(H), int(W), int(k), int(stride), int(pad), int(Ho),
                                                                    int(Wo), col, vec_in, vec_out);
    } else {
      AGRS_CUDA(ctx, cudaFuncSetAttribute(k_im2col_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTileSmemMax)));
      k_im2col_tile<false><<<grid, kTileThreads, smem, ctx->stream>>>(im, int(B), int(C), int(H), int(W), int(k), int(stride), int(pad), int(Ho),
                                                                     int(Wo), col, vec_in, vec_out);
    }
  } else {
    k_im2col<<<grid1d(ctx, total), 256, 0, ctx->stream>>>(im, total, int(C), int(H), int(W), int(k), int(stride), int(pad), int(Ho), int(Wo), col);
  }
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_col2im_f32(agrs_ctx* ctx, const float* col, int64_t B, int64_t C, int64_t Ho, int64_t Wo, int64_t k, int64_t stride,
                    float* im_out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_col2im_f32");
  AGRS_REQUIRE(ctx, B >= 0 && C >= 0 && Ho > 0 && Wo > 0 && k > 0 && stride > 0, AGRS_ERR_INVALID, "agrs_col2im_f32: bad geometry");
  int64_t Hi = (Ho - 1) * stride + k, Wi = (Wo - 1) * stride + k;
  int64_t total = B * C * Hi * Wi;
  if (total == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, col && im_out, AGRS_ERR_INVALID, "agrs_col2im_f32: null pointer");
  const int64_t per = Ho * Wo * C * k * k;
  const size_t smem = size_t(per) * sizeof(float);
  if (smem <= kTileSmemMax && C * Hi * Wi < (1LL << 30) && B < (1LL << 31)) {
    const int vec_in = (per % 4 == 0) && agrs_aligned16(col);
    const size_t per_pad = size_t((per + 3) & ~int64_t(3)) * sizeof(float);
    static const int nbuf_env = getenv("AGRS_COL2IM_NBUF") ? atoi(getenv("AGRS_COL2IM_NBUF")) : 3;
    int nbuf = 1;  // samples resident per CTA: as many as fit (up to 3) when there are enough samples to pipeline
    if (B >= 2 * int64_t(ctx->sm_count) && nbuf_env >= 2) nbuf = (nbuf_env >= 3 && 3 * per_pad <= kTileSmemMax) ? 3 : (2 * per_pad <= kTileSmemMax ? 2 : 1);
    const int64_t per_sm = nbuf == 1 ? 4 : (int64_t(kTileSmemMax) / int64_t(nbuf * per_pad) >= 2 ? 2 : 1);
    const int64_t cap = int64_t(ctx->sm_count) * per_sm;
    const unsigned grid = unsigned(B < cap ? B : cap);
#define AGRS_COL2IM(NB_, KS_)                                                                                                        \
  do {                                                                                                                                \
    AGRS_CUDA(ctx, cudaFuncSetAttribute(k_col2im_tile<NB_, KS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTileSmemMax)));    \
    k_col2im_tile<NB_, KS_><<<grid, NB_ >= 2 ? 1024 : kTileThreads, NB_ * per_pad, ctx->stream>>>(col, int(B), int(C), int(Hi), int(Wi), int(k),    \
                                                                                                 int(stride), int(Ho), int(Wo), im_out, vec_in); \
  } while (0)
#define AGRS_COL2IM_K(NB_)                          \
  do {                                              \
    if (k == 5) AGRS_COL2IM(NB_, 5);                \
    else if (k == 3) AGRS_COL2IM(NB_, 3);           \
    else if (k == 2) AGRS_COL2IM(NB_, 2);           \
    else AGRS_COL2IM(NB_, 0);                       \
  } while (0)
    if (nbuf == 3) AGRS_COL2IM_K(3);
    else if (nbuf == 2) AGRS_COL2IM_K(2);
    else AGRS_COL2IM_K(1);
#undef AGRS_COL2IM_K
#undef AGRS_COL2IM
  } else {
    k_col2im<<<grid1d(ctx, total), 256, 0, ctx->stream>>>(col, total, int(C), int(Hi), int(Wi), int(k), int(stride), int(Ho), int(Wo), im_out);
  }
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_broadcast_filter_f32(agrs_ctx* ctx, const float* filt, int64_t kk, int64_t c_out, int64_t c_in, float* out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_broadcast_filter_f32");
  AGRS_REQUIRE(ctx, kk >= 0 && c_out >= 0 && c_in >= 0, AGRS_ERR_INVALID, "agrs_broadcast_filter_f32: negative extent");
  int64_t per = kk * c_out, total = per * c_in;
  if (total == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, filt && out, AGRS_ERR_INVALID, "agrs_broadcast_filter_f32: null pointer");
  k_broadcast_filter<<<unsigned((total + 255) / 256), 256, 0, ctx->stream>>>(filt, per, total, out);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}


// shapes the implicit-GEMM kernels take (anything else goes through im2col + GEMM)
static bool conv1_shape_ok(int64_t B, int64_t H, int64_t W, int64_t k, int64_t co, int64_t stride, int64_t pad, size_t* smem_fwd, size_t* smem_bwd) {
  if (B < 1 || k < 1 || k > 11 || co < 1 || co > kConvCo || stride < 1 || pad < 0 || H + 2 * pad < k || W + 2 * pad < k) return false;
  const int64_t Ho = (H - k + 2 * pad) / stride + 1, Wo = (W - k + 2 * pad) / stride + 1;
  if (k * k * co + co > 256) return false;  // one dw / db accumulator per thread
  if (2 * (k * k + 1) > 256) return false;   // at least one position slice of (tap, channel half) threads
  const size_t img = (size_t(H + 2 * pad) * size_t(W + 2 * pad) + 3) & ~size_t(3), fl = size_t(k * k * 8), P = size_t(Ho * Wo);
  const size_t slices = 256 / (2 * size_t(k * k + 1));
  *smem_fwd = (fl + 8 + img + P * size_t(co)) * sizeof(float);
  const size_t gtile = size_t(Ho + 2 * (k - 1)) * size_t(Wo + 2 * (k - 1)) * 8;  // gradient tile with its zero border
  *smem_bwd = (fl + gtile + img + slices * size_t(k * k + 1) * 8) * sizeof(float);
  return *smem_bwd <= 160 * 1024 && *smem_fwd <= 160 * 1024 && B < (int64_t(1) << 31) && H < 32768 && W < 32768;
}

int agrs_conv2d_implicit_ok(int64_t B, int64_t C, int64_t H, int64_t W, int64_t k, int64_t c_out, int64_t stride, int64_t pad) {
  size_t a, b;
  return C == 1 && conv1_shape_ok(B, H, W, k, c_out, stride, pad, &a, &b) ? 1 : 0;
}

static int conv2d_fwd(agrs_ctx* ctx, const float* im, const float* filt, const float* bias, int64_t B, int64_t H, int64_t W, int64_t k,
                      int64_t c_out, int64_t stride, int64_t pad, float* out, int act_kind, float* y);

int agrs_conv2d_fwd_f32(agrs_ctx* ctx, const float* im, const float* filt, const float* bias, int64_t B, int64_t H, int64_t W, int64_t k,
                        int64_t c_out, int64_t stride, int64_t pad, float* out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_conv2d_fwd_f32");
  return conv2d_fwd(ctx, im, filt, bias, B, H, W, k, c_out, stride, pad, out, AGRS_ACT_NONE, nullptr);
}

int agrs_conv2d_fwd_act_f32(agrs_ctx* ctx, const float* im, const float* filt, const float* bias, int64_t B, int64_t H, int64_t W, int64_t k,
                            int64_t c_out, int64_t stride, int64_t pad, float* out, int act, float* y) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_conv2d_fwd_act_f32");
  AGRS_REQUIRE(ctx, act == AGRS_ACT_RELU || act == AGRS_ACT_SIGMOID, AGRS_ERR_INVALID, "agrs_conv2d_fwd_act_f32: unknown activation %d", act);
  AGRS_REQUIRE(ctx, y && y != out, AGRS_ERR_INVALID, "agrs_conv2d_fwd_act_f32: null y, or y aliases out (the pre-activation stays a tensor of its own)");
  return conv2d_fwd(ctx, im, filt, bias, B, H, W, k, c_out, stride, pad, out, act, y);
}

static int conv2d_fwd(agrs_ctx* ctx, const float* im, const float* filt, const float* bias, int64_t B, int64_t H, int64_t W, int64_t k,
                      int64_t c_out, int64_t stride, int64_t pad, float* out, int act_kind, float* y) {
  size_t smem = 0, smem_b = 0;
  AGRS_REQUIRE(ctx, conv1_shape_ok(B, H, W, k, c_out, stride, pad, &smem, &smem_b), AGRS_ERR_UNSUPPORTED,
               "agrs_conv2d_fwd_f32: shape outside the implicit-GEMM kernels (single input channel, C_out <= %d, k*k*C_out + C_out <= 256): use im2col + GEMM", kConvCo);
  AGRS_REQUIRE(ctx, im && filt && bias && out, AGRS_ERR_INVALID, "agrs_conv2d_fwd_f32: null pointer");
  const int Ho = int((H - k + 2 * pad) / stride + 1), Wo = int((W - k + 2 * pad) / stride + 1);
  ConvAct act{act_kind, y, nullptr};
  if (act_kind) {
    if (int rc = agrs_take_mx(ctx, &act.mx)) return rc;  // agrs_next_output_absmax reports on y
  }
  if (ctx->conv_tiled && conv1_fwd_tiled_ok(B, H, W, k, c_out, stride, pad, im)) {
    const size_t tsmem = cf_smem(int(H), int(W), int(k), int(c_out));
    const int64_t groups = (B + kCfSamples - 1) / kCfSamples, tcap = int64_t(ctx->sm_count) * 2;
    const unsigned tgrid = unsigned(groups < tcap ? groups : tcap);
    if (k == 5) {
      if (!(ctx->func_attr_done & 1u)) { AGRS_CUDA(ctx, cudaFuncSetAttribute(k_conv1_fwd_tiled<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); ctx->func_attr_done |= 1u; }
      k_conv1_fwd_tiled<5><<<tgrid, kCfThreads, tsmem, ctx->stream>>>(im, filt, bias, int(B), int(H), int(W), int(c_out), out, act);
    } else {
      if (!(ctx->func_attr_done & 2u)) { AGRS_CUDA(ctx, cudaFuncSetAttribute(k_conv1_fwd_tiled<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); ctx->func_attr_done |= 2u; }
      k_conv1_fwd_tiled<3><<<tgrid, kCfThreads, tsmem, ctx->stream>>>(im, filt, bias, int(B), int(H), int(W), int(c_out), out, act);
    }
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  const int64_t cap = int64_t(ctx->sm_count) * 4;
  if (!(ctx->func_attr_done & 4u)) {  // per context: the attribute belongs to the device the context is on
    AGRS_CUDA(ctx, cudaFuncSetAttribute(k_conv1_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    ctx->func_attr_done |= 4u;
  }
  k_conv1_fwd<<<unsigned(B < cap ? B : cap), 256, smem, ctx->stream>>>(im, filt, bias, int(B), int(H), int(W), int(k), int(c_out), int(stride), int(pad), Ho, Wo, out, act);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_conv2d_bwd_f32(agrs_ctx* ctx, const float* im, const float* filt, const float* grad, int64_t B, int64_t H, int64_t W, int64_t k,
                        int64_t c_out, int64_t stride, int64_t pad, float* dx, float* dw, float* db) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_conv2d_bwd_f32");
  size_t smem_f = 0, smem = 0;
  AGRS_REQUIRE(ctx, conv1_shape_ok(B, H, W, k, c_out, stride, pad, &smem_f, &smem), AGRS_ERR_UNSUPPORTED,
               "agrs_conv2d_bwd_f32: shape outside the implicit-GEMM kernels: use the im2col path");
  AGRS_REQUIRE(ctx, im && filt && grad && dx && dw && db, AGRS_ERR_INVALID, "agrs_conv2d_bwd_f32: null pointer");
  const int Ho = int((H - k + 2 * pad) / stride + 1), Wo = int((W - k + 2 * pad) / stride + 1);
  const size_t nacc = size_t(k * k * c_out + c_out);
  if (ctx->conv_tiled && conv1_tiled_ok(B, H, W, k, c_out, stride, pad, im, grad, dx)) {
    const CtGeom ge = ct_geom(int(H), int(W), int(k), int(c_out));
    const int64_t groups = (B + kCtSamples - 1) / kCtSamples;
    const int64_t tgrid = groups < ctx->sm_count ? groups : ctx->sm_count;
    int rc = agrs_ensure_workspace(ctx, size_t(tgrid) * nacc * sizeof(float));
    if (rc) return rc;
    if (!ctx->tickets) {
      AGRS_NOT_CAPTURING(ctx, "allocating the ticket array (run one ordinary step before capturing)");
      AGRS_CUDA(ctx, cudaMalloc(&ctx->tickets, 4096 * sizeof(unsigned int)));
      AGRS_CUDA(ctx, cudaMemsetAsync(ctx->tickets, 0, 4096 * sizeof(unsigned int), ctx->stream));
      ctx->tickets_n = 4096;
    }
    if (k == 5) {
      if (!(ctx->func_attr_done & 8u)) { AGRS_CUDA(ctx, cudaFuncSetAttribute(k_conv1_bwd_tiled<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); ctx->func_attr_done |= 8u; }
      k_conv1_bwd_tiled<5><<<unsigned(tgrid), kCtThreads, ge.smem, ctx->stream>>>(im, filt, grad, int(B), int(H), int(W), int(c_out), dx, dw, db, ctx->workspace, ctx->tickets);
    } else {
      if (!(ctx->func_attr_done & 16u)) { AGRS_CUDA(ctx, cudaFuncSetAttribute(k_conv1_bwd_tiled<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); ctx->func_attr_done |= 16u; }
      k_conv1_bwd_tiled<3><<<unsigned(tgrid), kCtThreads, ge.smem, ctx->stream>>>(im, filt, grad, int(B), int(H), int(W), int(c_out), dx, dw, db, ctx->workspace, ctx->tickets);
    }
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  const int64_t cap = int64_t(ctx->sm_count) * 4;
  const int64_t grid = B < cap ? B : cap;
  int rc = agrs_ensure_workspace(ctx, size_t(grid) * nacc * sizeof(float));
  if (rc) return rc;
  if (ctx->tickets_n < 1) {
    AGRS_NOT_CAPTURING(ctx, "allocating the ticket array (run one ordinary step before capturing)");
    AGRS_CUDA(ctx, cudaMalloc(&ctx->tickets, 4096 * sizeof(unsigned int)));
    AGRS_CUDA(ctx, cudaMemsetAsync(ctx->tickets, 0, 4096 * sizeof(unsigned int), ctx->stream));
    ctx->tickets_n = 4096;
  }
#define AGRS_CONV_BWD(S1_, KT_)                                                                                                            \
  do {                                                                                                                                   \
    const uint32_t bit = 32u << ((S1_ ? 1 : 0) + (KT_ == 5 ? 2 : KT_ == 3 ? 4 : 0));                                                      \
    if (!(ctx->func_attr_done & bit)) {                                                                                                  \
      AGRS_CUDA(ctx, cudaFuncSetAttribute(k_conv1_bwd<S1_, KT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));               \
      ctx->func_attr_done |= bit;                                                                                                        \
    }                                                                                                                                    \
    k_conv1_bwd<S1_, KT_><<<unsigned(grid), 256, smem, ctx->stream>>>(im, filt, grad, int(B), int(H), int(W), int(k), int(c_out), int(stride), \
                                                                      int(pad), Ho, Wo, dx, dw, db, ctx->workspace, ctx->tickets);        \
  } while (0)
  if (stride != 1) AGRS_CONV_BWD(false, 0);
  else if (k == 5) AGRS_CONV_BWD(true, 5);
  else if (k == 3) AGRS_CONV_BWD(true, 3);
  else             AGRS_CONV_BWD(true, 0);
#undef AGRS_CONV_BWD
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

}  // extern "C"
