// HBM-bound elementwise kernels: activations, MSE backward, tensor arithmetic, SGD.
// Every kernel touches each tensor exactly once, 16 bytes per thread per access, four
// independent accesses in flight per thread, grid = a multiple of the SM count (persistent
// grid-stride).  Streaming cache hints: nothing here is re-read by the same kernel.
#include "agrs_internal.h"
#include "mag.cuh"
#include "act.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kBlocksPerSM = 8;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }


// ---- functors (semantics cited from the reference) ----
struct ReluFwd {  // functional_ndarray.rs:14-22: x > 0 ? x : 0  (-0.0, NaN -> +0.0); act.cuh
  __device__ float operator()(float x) const { return agrs_act_relu(x); }
};
struct ReluBwd {  // functional_ndarray.rs:23-34: x <= 0 ? 0 : g  (NaN x passes g)
  __device__ float operator()(float x, float g) const { return x <= 0.f ? 0.f : g; }
};
struct SigmoidFwd {  // operators.rs:190-199: 1/(1+exp(-x)) in f32 (full-precision expf, IEEE divide); act.cuh
  __device__ float operator()(float x) const { return agrs_act_sigmoid(x); }
};
struct SigmoidBwd {  // operators.rs:219-223: ((1 - s) * s) * g
  __device__ float operator()(float x, float g) const {
    float s = 1.0f / (1.0f + expf(-x));
    return ((1.0f - s) * s) * g;
  }
};
struct MseBwd {  // operators.rs:249-253: ((p - t) * 2.0 / B) * grad
  float batch;
  const float* up;
  __device__ float operator()(float p, float t) const { return (((p - t) * 2.0f) / batch) * __ldg(up); }
};
struct AddOp { __device__ float operator()(float a, float b) const { return a + b; } };
struct SubOp { __device__ float operator()(float a, float b) const { return a - b; } };
struct MulOp { __device__ float operator()(float a, float b) const { return a * b; } };
struct SMul { float s; __device__ float operator()(float a) const { return a * s; } };
struct SDiv { float s; __device__ float operator()(float a) const { return a / s; } };
struct SRsub { float s; __device__ float operator()(float a) const { return s - a; } };
struct SAdd { float s; __device__ float operator()(float a) const { return a + s; } };
struct NegOp { __device__ float operator()(float a) const { return -a; } };
struct PowiOp {  // f32::powi: repeated multiplication (exp >= 0) or reciprocal of it
  int e;
  __device__ float operator()(float a) const {
    int n = e < 0 ? -e : e;
    float r = 1.0f, b = a;
    while (n) {
      if (n & 1) r *= b;
      b *= b;
      n >>= 1;
    }
    return e < 0 ? 1.0f / r : r;
  }
};
struct Sq { __device__ float operator()(float a) const { return a * a; } };
struct FillOp { float v; };

// ---- unary: out[i] = f(a[i]) ----  (out may BE a: powi_inplace, *= scalar, /= scalar run in place -- no __restrict__ here)
template <class F, bool MX>
__global__ void __launch_bounds__(kThreads) k_unary_v4(F f, float* out, const float* a, size_t n4, uint32_t* mx) {
  size_t stride = size_t(gridDim.x) * kThreads;
  size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x;
  uint32_t m = 0;
  for (; i + (kUnroll - 1) * stride < n4; i += kUnroll * stride) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = ld4(a + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float4 r = make_float4(f(v[u].x), f(v[u].y), f(v[u].z), f(v[u].w));
      if (MX) m = mag4(m, r);
      st4(out + 4 * (i + u * stride), r);
    }
  }
  for (; i < n4; i += stride) {
    float4 v = ld4(a + 4 * i);
    float4 r = make_float4(f(v.x), f(v.y), f(v.z), f(v.w));
    if (MX) m = mag4(m, r);
    st4(out + 4 * i, r);
  }
  if (MX) mag_commit(m, mx);
}
template <class F, bool MX>
__global__ void k_unary_s(F f, float* out, const float* a, size_t begin, size_t n, uint32_t* mx) {
  size_t i = begin + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  uint32_t m = 0;
  for (; i < n; i += stride) {
    const float r = f(a[i]);
    if (MX) m = max(m, mag_bits(r));
    out[i] = r;
  }
  if (MX) mag_commit(m, mx);
}

// ---- binary, same length: out[i] = f(a[i], b[i]) ----
template <class F, bool MX>
__global__ void __launch_bounds__(kThreads) k_binary_v4(F f, float* out, const float* a, const float* __restrict__ b, size_t n4, uint32_t* mx) {
  size_t stride = size_t(gridDim.x) * kThreads;
  size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x;
  uint32_t m = 0;
  for (; i + (kUnroll - 1) * stride < n4; i += kUnroll * stride) {
    float4 x[kUnroll], y[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      x[u] = ld4(a + 4 * (i + u * stride));
      y[u] = ld4(b + 4 * (i + u * stride));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float4 r = make_float4(f(x[u].x, y[u].x), f(x[u].y, y[u].y), f(x[u].z, y[u].z), f(x[u].w, y[u].w));
      if (MX) m = mag4(m, r);
      st4(out + 4 * (i + u * stride), r);
    }
  }
  for (; i < n4; i += stride) {
    float4 x = ld4(a + 4 * i), y = ld4(b + 4 * i);
    float4 r = make_float4(f(x.x, y.x), f(x.y, y.y), f(x.z, y.z), f(x.w, y.w));
    if (MX) m = mag4(m, r);
    st4(out + 4 * i, r);
  }
  if (MX) mag_commit(m, mx);
}
template <class F, bool MX>
__global__ void k_binary_s(F f, float* out, const float* a, const float* b, size_t begin, size_t n, uint32_t* mx) {
  size_t i = begin + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  uint32_t m = 0;
  for (; i < n; i += stride) {
    const float r = f(a[i], b[i]);
    if (MX) m = max(m, mag_bits(r));
    out[i] = r;
  }
  if (MX) mag_commit(m, mx);
}

// ---- binary with broadcast rhs ----
// trailing: b index = i % nb.  When nb % 4 == 0 rows are float4-tiled: (row, c4) decomposition keeps
// the rhs load a float4 that stays in L1/L2 (nb*4 bytes, re-read by every row).
template <class F>
__global__ void __launch_bounds__(kThreads) k_bcast_trailing_v4(F f, float* out, const float* a, const float* __restrict__ b, size_t n4,
                                                               unsigned nb4) {
  size_t stride = size_t(gridDim.x) * kThreads;
  for (size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    unsigned c = unsigned(i % nb4);
    float4 x = ld4(a + 4 * i);
    float4 y = __ldg(reinterpret_cast<const float4*>(b) + c);
    st4(out + 4 * i, make_float4(f(x.x, y.x), f(x.y, y.y), f(x.z, y.z), f(x.w, y.w)));
  }
}
template <class F>
__global__ void k_bcast_s(F f, float* out, const float* a, const float* __restrict__ b, size_t n, size_t nb, size_t rep, int leading) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    size_t j = leading ? i / rep : i % nb;
    out[i] = f(a[i], __ldg(b + j));
  }
}

__global__ void __launch_bounds__(kThreads) k_fill_v4(float* out, float v, size_t n4) {
  size_t stride = size_t(gridDim.x) * kThreads;
  float4 r = make_float4(v, v, v, v);
  for (size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) st4(out + 4 * i, r);
}
__global__ void k_fill_s(float* out, float v, size_t begin, size_t n) {
  size_t i = begin + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = v;
}

// SGD: w -= lr * g  (optim.rs:31-33: tmp = g * lr; w -= tmp -- two roundings, kept: no FMA contraction)
struct SgdOp {
  float lr;
  __device__ float operator()(float w, float g) const { return __fsub_rn(w, __fmul_rn(g, lr)); }
};

int grid_for(agrs_ctx* ctx, size_t work_items, int per_thread) {
  size_t blocks = (work_items + size_t(kThreads) * per_thread - 1) / (size_t(kThreads) * per_thread);
  size_t cap = size_t(ctx->sm_count) * kBlocksPerSM;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

template <class F>
int launch_unary(agrs_ctx* ctx, F f, float* out, const float* a, size_t n) {
  uint32_t* mx = nullptr;
  if (int rc = agrs_take_mx(ctx, &mx)) return rc;
  if (n == 0) return AGRS_OK;
  size_t n4 = (agrs_aligned16(out) && agrs_aligned16(a)) ? n / 4 : 0;
  if (n4) {
    if (mx) k_unary_v4<F, true><<<grid_for(ctx, n4, kUnroll), kThreads, 0, ctx->stream>>>(f, out, a, n4, mx);
    else    k_unary_v4<F, false><<<grid_for(ctx, n4, kUnroll), kThreads, 0, ctx->stream>>>(f, out, a, n4, nullptr);
    AGRS_LAUNCHED(ctx);
  }
  if (n4 * 4 < n) {
    if (mx) k_unary_s<F, true><<<grid_for(ctx, n - n4 * 4, 1), kThreads, 0, ctx->stream>>>(f, out, a, n4 * 4, n, mx);
    else    k_unary_s<F, false><<<grid_for(ctx, n - n4 * 4, 1), kThreads, 0, ctx->stream>>>(f, out, a, n4 * 4, n, nullptr);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}

template <class F>
int launch_binary(agrs_ctx* ctx, F f, float* out, const float* a, const float* b, size_t n) {
  uint32_t* mx = nullptr;
  if (int rc = agrs_take_mx(ctx, &mx)) return rc;
  if (n == 0) return AGRS_OK;
  size_t n4 = (agrs_aligned16(out) && agrs_aligned16(a) && agrs_aligned16(b)) ? n / 4 : 0;
  if (n4) {
    if (mx) k_binary_v4<F, true><<<grid_for(ctx, n4, kUnroll), kThreads, 0, ctx->stream>>>(f, out, a, b, n4, mx);
    else    k_binary_v4<F, false><<<grid_for(ctx, n4, kUnroll), kThreads, 0, ctx->stream>>>(f, out, a, b, n4, nullptr);
    AGRS_LAUNCHED(ctx);
  }
  if (n4 * 4 < n) {
    if (mx) k_binary_s<F, true><<<grid_for(ctx, n - n4 * 4, 1), kThreads, 0, ctx->stream>>>(f, out, a, b, n4 * 4, n, mx);
    else    k_binary_s<F, false><<<grid_for(ctx, n - n4 * 4, 1), kThreads, 0, ctx->stream>>>(f, out, a, b, n4 * 4, n, nullptr);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}

// out[0] = bits of the largest finite |x[i]| : the stand-alone pass, for tensors no producer reported on (uploads, foreign writers)
template <bool V4>
__global__ void __launch_bounds__(kThreads) k_absmax(const float* __restrict__ x, size_t begin, size_t n, uint32_t* mx) {
  const size_t stride = size_t(gridDim.x) * kThreads;
  uint32_t m = 0;
  if (V4) {
    size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
      float4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) v[u] = ld4(x + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) m = mag4(m, v[u]);
    }
    for (; i < n; i += stride) m = mag4(m, ld4(x + 4 * i));
  } else {
    for (size_t i = begin + size_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) m = max(m, mag_bits(x[i]));
  }
  mag_commit(m, mx);
}

// ---- SGD on up to 16 parameters in ONE launch (SGD::step walks the parameter list, optim.rs:25-34) ----
// blockIdx.y = parameter; the blocks of a parameter stride over it like k_binary_v4 does.  The launch-bound configurations pay one
// launch instead of one per tensor (LeNet: 4 parameters, fit_sine: 4, each a ~1.3 us kernel otherwise); the arithmetic is SgdOp's.
constexpr int kMultiMax = 16;
struct SgdMulti {
  float* w[kMultiMax];
  const float* g[kMultiMax];
  size_t n[kMultiMax];
  uint32_t* mx[kMultiMax];  // magnitude word of the updated parameter, or null
  int count;
  float lr;
};
__global__ void k_zero_words(SgdMulti a) {
  if (int(threadIdx.x) < a.count && a.mx[threadIdx.x]) *a.mx[threadIdx.x] = 0u;
}
__global__ void __launch_bounds__(kThreads) k_sgd_multi(const __grid_constant__ SgdMulti a) {
  const int t = blockIdx.y;
  float* w = a.w[t];
  const float* g = a.g[t];
  const size_t n = a.n[t];
  uint32_t* mx = a.mx[t];
  const SgdOp f{a.lr};
  const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0;
  const size_t n4 = vec ? n / 4 : 0, stride = size_t(gridDim.x) * kThreads;
  if (size_t(blockIdx.x) * kThreads >= (n4 ? n4 : n)) return;  // more blocks than this (small) parameter needs; warp-uniform
  uint32_t m = 0;
  size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x;
  for (; i + (kUnroll - 1) * stride < n4; i += kUnroll * stride) {
    float4 x[kUnroll], y[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      x[u] = ld4(w + 4 * (i + u * stride));
      y[u] = ld4(g + 4 * (i + u * stride));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const float4 r = make_float4(f(x[u].x, y[u].x), f(x[u].y, y[u].y), f(x[u].z, y[u].z), f(x[u].w, y[u].w));
      m = mag4(m, r);
      st4(w + 4 * (i + u * stride), r);
    }
  }
  for (; i < n4; i += stride) {
    const float4 x = ld4(w + 4 * i), y = ld4(g + 4 * i);
    const float4 r = make_float4(f(x.x, y.x), f(x.y, y.y), f(x.z, y.z), f(x.w, y.w));
    m = mag4(m, r);
    st4(w + 4 * i, r);
  }
  for (size_t e = n4 * 4 + size_t(blockIdx.x) * kThreads + threadIdx.x; e < n; e += stride) {  // ragged tail / unaligned parameter
    const float r = f(w[e], g[e]);
    m = max(m, mag_bits(r));
    w[e] = r;
  }
  if (mx) mag_commit(m, mx);  // block-uniform branch: every lane of every warp that got here arrives
}

template <class F>
int launch_bcast(agrs_ctx* ctx, F f, float* out, const float* a, const float* b, size_t n, size_t nb, int bmode) {
  if (n == 0) return AGRS_OK;
  if (nb == n) return launch_binary(ctx, f, out, a, b, n);
  AGRS_REQUIRE(ctx, !ctx->next_out_mx, AGRS_ERR_UNSUPPORTED, "agrs_next_output_absmax: broadcast forms do not report magnitudes");
  AGRS_REQUIRE(ctx, nb > 0 && n % nb == 0, AGRS_ERR_SHAPE, "cannot broadcast %zu elements onto %zu", nb, n);
  if (bmode == AGRS_BCAST_TRAILING && nb % 4 == 0 && nb / 4 < (size_t(1) << 31) && agrs_aligned16(out) && agrs_aligned16(a) &&
      agrs_aligned16(b)) {
    k_bcast_trailing_v4<F><<<grid_for(ctx, n / 4, 1), kThreads, 0, ctx->stream>>>(f, out, a, b, n / 4, unsigned(nb / 4));
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  k_bcast_s<F><<<grid_for(ctx, n, 1), kThreads, 0, ctx->stream>>>(f, out, a, b, n, nb, n / nb, bmode == AGRS_BCAST_LEADING);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

}  // namespace

extern "C" {

int agrs_next_output_absmax(agrs_ctx* ctx, uint32_t* out_bits) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_next_output_absmax");
  ctx->next_out_mx = out_bits;
  return AGRS_OK;
}

int agrs_absmax_f32(agrs_ctx* ctx, const float* x, size_t n, uint32_t* out_bits) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_absmax_f32");
  AGRS_REQUIRE(ctx, out_bits && (n == 0 || x), AGRS_ERR_INVALID, "agrs_absmax_f32: null pointer");
  if (int rc = agrs_zero_word(ctx, out_bits)) return rc;
  if (n == 0) return AGRS_OK;
  const size_t n4 = agrs_aligned16(x) ? n / 4 : 0;
  if (n4) {
    k_absmax<true><<<grid_for(ctx, n4, kUnroll), kThreads, 0, ctx->stream>>>(x, 0, n4, out_bits);
    AGRS_LAUNCHED(ctx);
  }
  if (n4 * 4 < n) {
    k_absmax<false><<<grid_for(ctx, n - n4 * 4, 1), kThreads, 0, ctx->stream>>>(x, n4 * 4, n, out_bits);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}

int agrs_fill_f32(agrs_ctx* ctx, float* x, float value, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_fill_f32");
  if (n == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, x, AGRS_ERR_INVALID, "agrs_fill_f32: null pointer");
  if (value == 0.0f) {  // zero_grad / zeros: the copy engine's memset
    AGRS_CUDA(ctx, cudaMemsetAsync(x, 0, n * sizeof(float), ctx->stream));
    return AGRS_OK;
  }
  size_t n4 = agrs_aligned16(x) ? n / 4 : 0;
  if (n4) {
    k_fill_v4<<<grid_for(ctx, n4, 1), kThreads, 0, ctx->stream>>>(x, value, n4);
    AGRS_LAUNCHED(ctx);
  }
  if (n4 * 4 < n) {
    k_fill_s<<<grid_for(ctx, n - n4 * 4, 1), kThreads, 0, ctx->stream>>>(x, value, n4 * 4, n);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}

int agrs_relu_fwd_f32(agrs_ctx* ctx, const float* x, float* y, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_relu_fwd_f32");
  AGRS_REQUIRE(ctx, n == 0 || (x && y), AGRS_ERR_INVALID, "agrs_relu_fwd_f32: null pointer");
  return launch_unary(ctx, ReluFwd{}, y, x, n);
}
int agrs_relu_bwd_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_relu_bwd_f32");
  AGRS_REQUIRE(ctx, n == 0 || (x && g && dx), AGRS_ERR_INVALID, "agrs_relu_bwd_f32: null pointer");
  return launch_binary(ctx, ReluBwd{}, dx, x, g, n);
}
int agrs_sigmoid_fwd_f32(agrs_ctx* ctx, const float* x, float* y, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sigmoid_fwd_f32");
  AGRS_REQUIRE(ctx, n == 0 || (x && y), AGRS_ERR_INVALID, "agrs_sigmoid_fwd_f32: null pointer");
  return launch_unary(ctx, SigmoidFwd{}, y, x, n);
}
int agrs_sigmoid_bwd_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sigmoid_bwd_f32");
  AGRS_REQUIRE(ctx, n == 0 || (x && g && dx), AGRS_ERR_INVALID, "agrs_sigmoid_bwd_f32: null pointer");
  return launch_binary(ctx, SigmoidBwd{}, dx, x, g, n);
}
int agrs_mse_bwd_f32(agrs_ctx* ctx, const float* pred, const float* target, const float* upstream1, float batch, size_t n,
                     float* out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_mse_bwd_f32");
  AGRS_REQUIRE(ctx, n == 0 || (pred && target && upstream1 && out), AGRS_ERR_INVALID, "agrs_mse_bwd_f32: null pointer");
  return launch_binary(ctx, MseBwd{batch, upstream1}, out, pred, target, n);
}

int agrs_binary_f32(agrs_ctx* ctx, int op, float* out, const float* a, const float* b, size_t n, size_t nb, int bmode) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_binary_f32");
  AGRS_REQUIRE(ctx, n == 0 || (out && a && b), AGRS_ERR_INVALID, "agrs_binary_f32: null pointer");
  AGRS_REQUIRE(ctx, bmode == AGRS_BCAST_TRAILING || bmode == AGRS_BCAST_LEADING, AGRS_ERR_INVALID, "unknown broadcast mode %d", bmode);
  switch (op) {
    case AGRS_ADD: return launch_bcast(ctx, AddOp{}, out, a, b, n, nb, bmode);
    case AGRS_SUB: return launch_bcast(ctx, SubOp{}, out, a, b, n, nb, bmode);
    case AGRS_MUL: return launch_bcast(ctx, MulOp{}, out, a, b, n, nb, bmode);
  }
  return agrs_fail(ctx, AGRS_ERR_INVALID, "unknown binary op %d", op);
}

int agrs_scalar_f32(agrs_ctx* ctx, int op, float* out, const float* a, float s, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_scalar_f32");
  AGRS_REQUIRE(ctx, n == 0 || (out && a), AGRS_ERR_INVALID, "agrs_scalar_f32: null pointer");
  switch (op) {
    case AGRS_SMUL: return launch_unary(ctx, SMul{s}, out, a, n);
    case AGRS_SDIV: return launch_unary(ctx, SDiv{s}, out, a, n);
    case AGRS_SRSUB: return launch_unary(ctx, SRsub{s}, out, a, n);
    case AGRS_SADD: return launch_unary(ctx, SAdd{s}, out, a, n);
  }
  return agrs_fail(ctx, AGRS_ERR_INVALID, "unknown scalar op %d", op);
}

int agrs_neg_f32(agrs_ctx* ctx, float* out, const float* a, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_neg_f32");
  AGRS_REQUIRE(ctx, n == 0 || (out && a), AGRS_ERR_INVALID, "agrs_neg_f32: null pointer");
  return launch_unary(ctx, NegOp{}, out, a, n);
}

int agrs_powi_f32(agrs_ctx* ctx, float* out, const float* a, int exp, size_t n) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_powi_f32");
  AGRS_REQUIRE(ctx, n == 0 || (out && a), AGRS_ERR_INVALID, "agrs_powi_f32: null pointer");
  if (exp == 2) return launch_unary(ctx, Sq{}, out, a, n);
  return launch_unary(ctx, PowiOp{exp}, out, a, n);
}

int agrs_sgd_step_multi_f32(agrs_ctx* ctx, int count, float* const* w, const float* const* g, const size_t* n, uint32_t* const* out_absmax, float lr) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sgd_step_multi_f32");
  AGRS_REQUIRE(ctx, count >= 0 && (count == 0 || (w && g && n)), AGRS_ERR_INVALID, "agrs_sgd_step_multi_f32: null pointer");
  for (int base = 0; base < count; base += kMultiMax) {
    SgdMulti a{};
    a.count = count - base < kMultiMax ? count - base : kMultiMax;
    a.lr = lr;
    size_t largest = 0;
    bool any_mx = false;
    for (int i = 0; i < a.count; ++i) {
      AGRS_REQUIRE(ctx, n[base + i] == 0 || (w[base + i] && g[base + i]), AGRS_ERR_INVALID, "agrs_sgd_step_multi_f32: null parameter %d", base + i);
      a.w[i] = w[base + i];
      a.g[i] = g[base + i];
      a.n[i] = n[base + i];
      a.mx[i] = out_absmax ? out_absmax[base + i] : nullptr;
      any_mx = any_mx || a.mx[i];
      if (a.n[i] > largest) largest = a.n[i];
    }
    if (largest == 0) continue;
    if (any_mx) {
      k_zero_words<<<1, kMultiMax, 0, ctx->stream>>>(a);
      AGRS_LAUNCHED(ctx);
    }
    dim3 grid(unsigned(grid_for(ctx, (largest + 3) / 4, kUnroll)), unsigned(a.count));
    k_sgd_multi<<<grid, kThreads, 0, ctx->stream>>>(a);
    AGRS_LAUNCHED(ctx);
  }
  return AGRS_OK;
}

int agrs_sgd_step_f32(agrs_ctx* ctx, float* w, const float* g, float lr, size_t n, size_t ng) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sgd_step_f32");
  AGRS_REQUIRE(ctx, n == 0 || (w && g), AGRS_ERR_INVALID, "agrs_sgd_step_f32: null pointer");
  return launch_bcast(ctx, SgdOp{lr}, w, w, g, n, ng, AGRS_BCAST_TRAILING);
}

}  // extern "C"
