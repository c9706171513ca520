// Reductions: sum / mean / fused MSE forward (storage.rs:91-112, operators.rs:237),
// sum_axis (db = g.sum_axis(0), operators.rs:161; col.sum_axis(-1), operators.rs:443), equality, transpose.
// All reductions are deterministic (fixed summation tree, no float atomics): replicas of a
// data-parallel job stay bit-identical.
#include "agrs_internal.h"
#include "mag.cuh"

namespace {

constexpr int kThreads = 256;

// ---- sources of the column-sum kernel: a plain tensor, or an activation / loss backward computed on the fly ----
// The fused sources write the elementwise result AND feed it to the column sum in the same pass: the gradient that ReLU /
// Sigmoid / MSE backward hand to the Linear layer below is exactly the tensor whose sum over rows that layer needs for db
// (operators.rs:161), so the separate pass that re-read it disappears.  Same arithmetic as the functors of elementwise.cu.
// Every source splits a 16-byte access into ld4 (the loads) and fin4 (arithmetic + the store of a writing source).  Written as one
// function per row, the store of row l stands between the loads of rows l and l + 8 (the pointers may alias as far as the compiler
// knows): one row in flight per thread.  Behind ReLU's two instructions and Sigmoid's expf that costs nothing measurable, and issuing
// the loads of all four rows first costs THEM 10 us of an [8192, 4096] tensor's 70 - 78; behind MSE's four IEEE divisions per vector it
// was 110 us instead of 82 (tools/gpu_r3_i.sh, gpu_r3_j.sh): kLoadsFirst chooses per source.  MSE's upstream scalar is read once per
// thread (prep) instead of after every store.
struct Raw2 {
  float4 a, b;
};
struct NoPre {};
struct SrcPlain {
  const float* x;
  static constexpr bool kWrites = false;
  static constexpr bool kLoadsFirst = false;
  using Raw = float4;
  __device__ NoPre prep() const { return NoPre{}; }
  __device__ Raw ld4(int64_t e) const { return __ldcs(reinterpret_cast<const float4*>(x + e)); }
  __device__ float4 fin4(int64_t, const Raw& r, NoPre) const { return r; }
  __device__ float at1(int64_t e) const { return __ldcs(x + e); }
};
struct SrcReluBwd {  // functional_ndarray.rs:23-34: x <= 0 ? 0 : g
  const float* x;
  const float* g;
  float* out;
  static constexpr bool kWrites = true;
  static constexpr bool kLoadsFirst = false;
  using Raw = Raw2;
  __device__ NoPre prep() const { return NoPre{}; }
  __device__ Raw ld4(int64_t e) const { return Raw{__ldcs(reinterpret_cast<const float4*>(x + e)), __ldcs(reinterpret_cast<const float4*>(g + e))}; }
  __device__ float4 fin4(int64_t e, const Raw& v, NoPre) const {
    const float4 a = v.a, b = v.b;
    const float4 r = make_float4(a.x <= 0.f ? 0.f : b.x, a.y <= 0.f ? 0.f : b.y, a.z <= 0.f ? 0.f : b.z, a.w <= 0.f ? 0.f : b.w);
    __stcs(reinterpret_cast<float4*>(out + e), r);
    return r;
  }
  __device__ float at1(int64_t) const { return 0.f; }
};
struct SrcSigmoidBwd {  // operators.rs:219-223: ((1 - s) * s) * g, s recomputed from x
  const float* x;
  const float* g;
  float* out;
  static constexpr bool kWrites = true;
  static constexpr bool kLoadsFirst = false;
  using Raw = Raw2;
  __device__ static float f(float x, float g) {
    const float s = 1.0f / (1.0f + expf(-x));
    return ((1.0f - s) * s) * g;
  }
  __device__ NoPre prep() const { return NoPre{}; }
  __device__ Raw ld4(int64_t e) const { return Raw{__ldcs(reinterpret_cast<const float4*>(x + e)), __ldcs(reinterpret_cast<const float4*>(g + e))}; }
  __device__ float4 fin4(int64_t e, const Raw& v, NoPre) const {
    const float4 a = v.a, b = v.b;
    const float4 r = make_float4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w));
    __stcs(reinterpret_cast<float4*>(out + e), r);
    return r;
  }
  __device__ float at1(int64_t) const { return 0.f; }
};
struct SrcMseBwd {  // operators.rs:249-253: ((p - t) * 2.0 / B) * grad
  const float* p;
  const float* t;
  const float* up;
  float batch;
  float* out;
  float inv_batch;  // 1 / batch when batch is a power of two (dividing by it and multiplying by this then round identically), else 0
  static constexpr bool kWrites = true;
  static constexpr bool kLoadsFirst = true;  // see above: the divisions sit between a row's loads and its store
  using Raw = Raw2;
  __device__ float f(float a, float b, float u) const {
    const float d2 = (a - b) * 2.0f;
    return (inv_batch != 0.f ? d2 * inv_batch : d2 / batch) * u;
  }
  __device__ float prep() const { return __ldg(up); }  // the upstream [1, 1] gradient: read once per thread
  __device__ Raw ld4(int64_t e) const { return Raw{__ldcs(reinterpret_cast<const float4*>(p + e)), __ldcs(reinterpret_cast<const float4*>(t + e))}; }
  __device__ float4 fin4(int64_t e, const Raw& v, float u) const {
    const float4 a = v.a, b = v.b;
    const float4 r = make_float4(f(a.x, b.x, u), f(a.y, b.y, u), f(a.z, b.z, u), f(a.w, b.w, u));
    __stcs(reinterpret_cast<float4*>(out + e), r);
    return r;
  }
  __device__ float at1(int64_t) const { return 0.f; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* sm) {  // sm: 32 floats
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sm[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

struct LoadId {
  const float* a;
  __device__ float4 v4(size_t i) const { return __ldcs(reinterpret_cast<const float4*>(a) + i); }
  __device__ float s(size_t i) const { return a[i]; }
  __device__ static float f(float x) { return x; }
};
struct LoadSqDiff {
  const float* p;
  const float* t;
  __device__ float4 v4(size_t i) const {
    float4 a = __ldcs(reinterpret_cast<const float4*>(p) + i), b = __ldcs(reinterpret_cast<const float4*>(t) + i);
    float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    return make_float4(d.x * d.x, d.y * d.y, d.z * d.z, d.w * d.w);
  }
  __device__ float s(size_t i) const {
    float d = p[i] - t[i];
    return d * d;
  }
};

// One launch: every block reduces a grid-stride slice to partial[blockIdx]; the last block to
// finish (atomic ticket) folds the partials in index order and writes out[0] * scale.
template <class L>
__global__ void __launch_bounds__(kThreads) k_reduce_all(L ld, size_t n4, size_t n, float scale, float* partial, unsigned* ticket,
                                                        float* out) {
  __shared__ float sm[32];
  __shared__ bool last;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  size_t stride = size_t(gridDim.x) * kThreads;
  size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v0 = ld.v4(i), v1 = ld.v4(i + stride), v2 = ld.v4(i + 2 * stride), v3 = ld.v4(i + 3 * stride);
    acc[0] += (v0.x + v0.y) + (v0.z + v0.w);
    acc[1] += (v1.x + v1.y) + (v1.z + v1.w);
    acc[2] += (v2.x + v2.y) + (v2.z + v2.w);
    acc[3] += (v3.x + v3.y) + (v3.z + v3.w);
  }
  for (; i < n4; i += stride) {
    float4 v = ld.v4(i);
    acc[0] += (v.x + v.y) + (v.z + v.w);
  }
  for (size_t j = n4 * 4 + size_t(blockIdx.x) * kThreads + threadIdx.x; j < n; j += stride) acc[1] += ld.s(j);
  float v = block_sum((acc[0] + acc[1]) + (acc[2] + acc[3]), sm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = v;
    __threadfence();
    unsigned t = atomicInc(ticket, gridDim.x - 1);  // wraps back to 0 for the next launch
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float s = 0.f;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += kThreads) s += __ldcg(partial + k);
    s = block_sum(s, sm);
    if (threadIdx.x == 0) out[0] = s * scale;
  }
}

template <class L>
int launch_reduce_all(agrs_ctx* ctx, L ld, size_t n, bool vec_ok, float scale, float* out) {
  size_t n4 = vec_ok ? n / 4 : 0;
  size_t work = n4 ? n4 : n;
  size_t blocks = (work + size_t(kThreads) * 4 - 1) / (size_t(kThreads) * 4);
  size_t cap = size_t(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_reduce_all<L><<<int(blocks), kThreads, 0, ctx->stream>>>(ld, n4, n, scale, ctx->scratch, ctx->counter, out);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

// ---- sum over the middle axis of [outer, len, inner], inner >= 2: coalesced along inner ----
// Block = 32 x 8 threads: x covers 128 consecutive inner elements as float4 (or 32 scalars),
// y strides the reduced axis; grid.y splits `len` into chunks whose partials go to a
// [chunks, outer*inner] workspace folded by k_fold_chunks (fixed order).
template <int VEC, class S>
__global__ void __launch_bounds__(256) k_sum_axis_mid(const S src, int64_t len, int64_t inner, int64_t chunk_len,
                                                      float* __restrict__ part /*[chunks][outer*inner]*/, int64_t out_elems,
                                                      unsigned int* __restrict__ tickets, float* __restrict__ out, uint32_t* mx) {
  __shared__ float sm[8][32 * VEC + 1];
  __shared__ unsigned int s_ticket;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int64_t col = (int64_t(blockIdx.x) * 32 + tx) * VEC;  // position inside inner
  int64_t o = blockIdx.z;
  int64_t l0 = int64_t(blockIdx.y) * chunk_len;
  int64_t l1 = l0 + chunk_len < len ? l0 + chunk_len : len;
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  uint32_t m = 0;  // magnitude of what a writing source produces (agrs_next_output_absmax), when asked for
  const auto pre = src.prep();
  auto at4 = [&](int64_t e) { return src.fin4(e, src.ld4(e), pre); };
  if (col < inner) {
    const int64_t base = o * len * inner + col;
    int64_t l = l0 + ty;
    if (VEC == 4) {  // four independent 16-byte loads in flight per thread (rows l, l+8, l+16, l+24)
      for (; l + 24 < l1; l += 32) {
        float4 v0, v1, v2, v3;
        if constexpr (S::kLoadsFirst) {
          const int64_t e0 = base + l * inner, e1 = base + (l + 8) * inner, e2 = base + (l + 16) * inner, e3 = base + (l + 24) * inner;
          const typename S::Raw r0 = src.ld4(e0), r1 = src.ld4(e1), r2 = src.ld4(e2), r3 = src.ld4(e3);
          v0 = src.fin4(e0, r0, pre);
          v1 = src.fin4(e1, r1, pre);
          v2 = src.fin4(e2, r2, pre);
          v3 = src.fin4(e3, r3, pre);
        } else {
          v0 = at4(base + l * inner);
          v1 = at4(base + (l + 8) * inner);
          v2 = at4(base + (l + 16) * inner);
          v3 = at4(base + (l + 24) * inner);
        }
        acc[0] += v0.x; acc[1 % VEC] += v0.y; acc[2 % VEC] += v0.z; acc[3 % VEC] += v0.w;
        acc[0] += v1.x; acc[1 % VEC] += v1.y; acc[2 % VEC] += v1.z; acc[3 % VEC] += v1.w;
        acc[0] += v2.x; acc[1 % VEC] += v2.y; acc[2 % VEC] += v2.z; acc[3 % VEC] += v2.w;
        acc[0] += v3.x; acc[1 % VEC] += v3.y; acc[2 % VEC] += v3.z; acc[3 % VEC] += v3.w;
        if (S::kWrites && mx) m = mag4(mag4(mag4(mag4(m, v0), v1), v2), v3);
      }
    }
    for (; l < l1; l += 8) {
      if (VEC == 4) {
        float4 v = at4(base + l * inner);
        acc[0] += v.x;
        acc[1 % VEC] += v.y;
        acc[2 % VEC] += v.z;
        acc[3 % VEC] += v.w;
        if (S::kWrites && mx) m = mag4(m, v);
      } else {
        acc[0] += src.at1(base + l * inner);
      }
    }
  }
  if (S::kWrites && mx) mag_commit(m, mx);  // warp-uniform branch: every lane arrives
#pragma unroll
  for (int v = 0; v < VEC; ++v) sm[ty][tx * VEC + v] = acc[v];
  __syncthreads();
  if (ty == 0 && col < inner) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) s += sm[r][tx * VEC + v];
      part[int64_t(blockIdx.y) * out_elems + o * inner + col + v] = s;
    }
  }
  if (gridDim.y == 1) return;  // single chunk: `part` is the output
  // One launch: the LAST chunk-block of this column group to finish folds the partials, always in chunk order
  // (deterministic), and resets the ticket for the next launch.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[o * gridDim.x + blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  __threadfence();
  const int64_t c0 = int64_t(blockIdx.x) * 32 * VEC;
  for (int t = threadIdx.x; t < 32 * VEC; t += blockDim.x) {
    if (c0 + t >= inner) break;
    float s = 0.f;
    for (unsigned c = 0; c < gridDim.y; ++c) s += __ldcg(part + int64_t(c) * out_elems + o * inner + c0 + t);
    out[o * inner + c0 + t] = s;
  }
  if (threadIdx.x == 0) tickets[o * gridDim.x + blockIdx.x] = 0;
}

// ---- outer == 1 and a SHORT inner axis (db of a narrow layer: [B*P, 6] -> [6], [1024, 10] -> [10]) ----
// The array is read as one flat, fully coalesced stream: T = 256 - 256 % inner threads are active and thread t owns the flat
// indices t, t+T, t+2T, ..; T is a multiple of inner, so every one of them lies in column t % inner.  Columns are then folded
// across threads, and across blocks by the last block to finish, always in index order (deterministic).
__global__ void __launch_bounds__(256) k_colsum_narrow(const float* __restrict__ x, int64_t n, int inner, int64_t elems_per_block,
                                                       float* __restrict__ part, unsigned int* __restrict__ ticket, float* __restrict__ out) {
  __shared__ float s[256];
  __shared__ unsigned int s_ticket;
  const int tid = threadIdx.x, T = 256 - 256 % inner;
  const int64_t b0 = int64_t(blockIdx.x) * elems_per_block, b1 = b0 + elems_per_block < n ? b0 + elems_per_block : n;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (tid < T) {
    int64_t i = b0 + tid;
    for (; i + 3 * int64_t(T) < b1; i += 4 * int64_t(T)) {
      a0 += __ldcs(x + i);
      a1 += __ldcs(x + i + T);
      a2 += __ldcs(x + i + 2 * int64_t(T));
      a3 += __ldcs(x + i + 3 * int64_t(T));
    }
    for (; i < b1; i += T) a0 += __ldcs(x + i);
  }
  s[tid] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  const int blocks = gridDim.x;
  if (tid < inner) {
    float c = 0.f;
    for (int t = tid; t < T; t += inner) c += s[t];
    if (blocks == 1) out[tid] = c;
    else part[int64_t(blockIdx.x) * inner + tid] = c;
  }
  if (blocks == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != unsigned(blocks - 1)) return;
  __threadfence();
  float c = 0.f;
  if (tid < T)
    for (int z = tid / inner; z < blocks; z += T / inner) c += __ldcg(part + int64_t(z) * inner + tid % inner);
  s[tid] = c;
  __syncthreads();
  if (tid < inner) {
    float r = 0.f;
    for (int t = tid; t < T; t += inner) r += s[t];
    out[tid] = r;
  }
  if (tid == 0) *ticket = 0;
}

// ---- inner == 1 (sum over the last axis): one warp per output row for long rows, one thread for short ----
__global__ void k_sum_last_warp(const float* __restrict__ x, int64_t rows, int64_t len, float* __restrict__ out) {
  int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int64_t l = lane; l < len; l += 32) s += x[row * len + l];
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}
__global__ void k_sum_last_thread(const float* __restrict__ x, int64_t rows, int64_t len, float* __restrict__ out) {
  int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float s = 0.f;
  for (int64_t l = 0; l < len; ++l) s += x[row * len + l];
  out[row] = s;
}

__global__ void __launch_bounds__(kThreads) k_not_equal(const float* __restrict__ a, const float* __restrict__ b, size_t n, unsigned* flag) {
  size_t stride = size_t(gridDim.x) * kThreads;
  bool ne = false;
  for (size_t i = size_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) ne |= (a[i] != b[i]);  // NaN != NaN, as ndarray ==
  if (__syncthreads_or(ne) && threadIdx.x == 0) atomicOr(flag, 1u);
}

__global__ void k_transpose(const float* __restrict__ in, int64_t rows, int64_t cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  int64_t c = int64_t(blockIdx.x) * 32 + threadIdx.x, r0 = int64_t(blockIdx.y) * 32;
  for (int j = threadIdx.y; j < 32; j += 8)
    if (c < cols && r0 + j < rows) tile[j][threadIdx.x] = in[(r0 + j) * cols + c];
  __syncthreads();
  int64_t r = r0 + threadIdx.x, c0 = int64_t(blockIdx.x) * 32;
  for (int j = threadIdx.y; j < 32; j += 8)
    if (r < rows && c0 + j < cols) out[(c0 + j) * rows + r] = tile[threadIdx.x][j];
}

}  // namespace

// the middle-axis column sum over source S ([outer, len, inner], inner >= 2), one launch, deterministic (see k_sum_axis_mid)
template <class S>
static int launch_sum_axis_mid(agrs_ctx* ctx, const S& src, bool vec, int64_t outer, int64_t len, int64_t inner, float* out, uint32_t* mx) {
  const int64_t out_elems = outer * inner;
  AGRS_REQUIRE(ctx, outer <= 65535, AGRS_ERR_UNSUPPORTED, "agrs_sum_axis_f32: outer=%lld > 65535 with inner > 1", (long long)outer);
  int per_block = vec ? 128 : 32;
  int64_t gx = (inner + per_block - 1) / per_block;
  // split the reduced axis so the grid fills the machine (>= ~4 blocks per SM), chunks of >= 64 rows
  int64_t want = (int64_t(ctx->sm_count) * 4 + gx * outer - 1) / (gx * outer);
  int64_t max_chunks = (len + 63) / 64;
  int64_t chunks = want < 1 ? 1 : (want > max_chunks ? max_chunks : want);
  if (chunks > 1024) chunks = 1024;
  int64_t chunk_len = (len + chunks - 1) / chunks;
  chunks = (len + chunk_len - 1) / chunk_len;
  float* part = out;
  if (chunks > 1) {
    int rc = agrs_ensure_workspace(ctx, size_t(chunks) * size_t(out_elems) * sizeof(float));
    if (rc) return rc;
    part = ctx->workspace;
  }
  if (chunks > 1 && size_t(gx * outer) > ctx->tickets_n) {
    AGRS_NOT_CAPTURING(ctx, "growing the column-sum ticket array (run one ordinary step before capturing)");
    AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->tickets) AGRS_CUDA(ctx, cudaFree(ctx->tickets));
    ctx->tickets = nullptr;
    ctx->tickets_n = 0;
    size_t n = size_t(gx * outer) < 4096 ? 4096 : size_t(gx * outer);
    AGRS_CUDA(ctx, cudaMalloc(&ctx->tickets, n * sizeof(unsigned int)));
    AGRS_CUDA(ctx, cudaMemsetAsync(ctx->tickets, 0, n * sizeof(unsigned int), ctx->stream));
    ctx->tickets_n = n;
  }
  dim3 grid((unsigned)gx, (unsigned)chunks, (unsigned)outer);
  if (vec)
    k_sum_axis_mid<4, S><<<grid, 256, 0, ctx->stream>>>(src, len, inner, chunk_len, part, out_elems, ctx->tickets, out, mx);
  else
    k_sum_axis_mid<1, S><<<grid, 256, 0, ctx->stream>>>(src, len, inner, chunk_len, part, out_elems, ctx->tickets, out, mx);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

// activation / loss backward fused with the column sum of its result: rows x cols, cols % 4 == 0, 16-byte aligned pointers
template <class S>
static int fused_bwd_colsum(agrs_ctx* ctx, const S& src, const char* who, int64_t rows, int64_t cols, const void* p0, const void* p1, const void* p2,
                            float* colsum) {
  AGRS_REQUIRE(ctx, rows > 0 && cols > 0 && p0 && p1 && p2 && colsum, AGRS_ERR_INVALID, "%s: null pointer or empty shape", who);
  AGRS_REQUIRE(ctx, cols % 4 == 0 && agrs_aligned16(p0) && agrs_aligned16(p1) && agrs_aligned16(p2), AGRS_ERR_UNSUPPORTED,
               "%s: needs cols %% 4 == 0 and 16-byte aligned tensors (use the unfused pair otherwise)", who);
  uint32_t* mx = nullptr;
  if (int rc = agrs_take_mx(ctx, &mx)) return rc;
  return launch_sum_axis_mid(ctx, src, true, 1, rows, cols, colsum, mx);
}

extern "C" {

int agrs_sum_f32(agrs_ctx* ctx, const float* x, size_t n, float* out1) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sum_f32");
  AGRS_REQUIRE(ctx, out1 && (n == 0 || x), AGRS_ERR_INVALID, "agrs_sum_f32: null pointer");
  return launch_reduce_all(ctx, LoadId{x}, n, agrs_aligned16(x), 1.0f, out1);
}

int agrs_mean_f32(agrs_ctx* ctx, const float* x, size_t n, float* out1) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_mean_f32");
  AGRS_REQUIRE(ctx, out1 && x && n > 0, AGRS_ERR_INVALID, "agrs_mean_f32: empty or null input");
  return launch_reduce_all(ctx, LoadId{x}, n, agrs_aligned16(x), 1.0f / float(n), out1);
}

int agrs_mse_fwd_f32(agrs_ctx* ctx, const float* pred, const float* target, size_t n, float* out1) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_mse_fwd_f32");
  AGRS_REQUIRE(ctx, out1 && pred && target && n > 0, AGRS_ERR_INVALID, "agrs_mse_fwd_f32: empty or null input");
  return launch_reduce_all(ctx, LoadSqDiff{pred, target}, n, agrs_aligned16(pred) && agrs_aligned16(target), 1.0f / float(n), out1);
}

int agrs_sum_axis_f32(agrs_ctx* ctx, const float* x, int64_t outer, int64_t len, int64_t inner, float* out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sum_axis_f32");
  AGRS_REQUIRE(ctx, outer >= 0 && len >= 0 && inner >= 0, AGRS_ERR_INVALID, "agrs_sum_axis_f32: negative extent");
  int64_t out_elems = outer * inner;
  if (out_elems == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, out && (len == 0 || x), AGRS_ERR_INVALID, "agrs_sum_axis_f32: null pointer");
  if (len == 0) return agrs_fill_f32(ctx, out, 0.f, size_t(out_elems));
  if (inner == 1) {
    if (len >= 64) {
      int64_t threads = outer * 32;
      k_sum_last_warp<<<unsigned((threads + 255) / 256), 256, 0, ctx->stream>>>(x, outer, len, out);
    } else {
      k_sum_last_thread<<<unsigned((outer + 255) / 256), 256, 0, ctx->stream>>>(x, outer, len, out);
    }
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  if (outer == 1 && inner < 32 && len >= 512) {
    const int64_t n = len * inner;
    const int T = 256 - 256 % int(inner);
    int64_t blocks = int64_t(ctx->sm_count) * 4, maxb = (n + 8 * T - 1) / (8 * T);
    if (blocks > maxb) blocks = maxb;
    int64_t epb = ((n + blocks - 1) / blocks + T - 1) / T * T;  // a multiple of T (hence of inner): block starts are column-aligned
    blocks = (n + epb - 1) / epb;
    float* part = nullptr;
    if (blocks > 1) {
      int rc = agrs_ensure_workspace(ctx, size_t(blocks) * size_t(inner) * sizeof(float));
      if (rc) return rc;
      if (ctx->tickets_n < 1) {
        AGRS_NOT_CAPTURING(ctx, "growing the ticket array (run one ordinary step before capturing)");
        AGRS_CUDA(ctx, cudaMalloc(&ctx->tickets, 4096 * sizeof(unsigned int)));
        AGRS_CUDA(ctx, cudaMemsetAsync(ctx->tickets, 0, 4096 * sizeof(unsigned int), ctx->stream));
        ctx->tickets_n = 4096;
      }
      part = ctx->workspace;
    }
    k_colsum_narrow<<<unsigned(blocks), 256, 0, ctx->stream>>>(x, n, int(inner), epb, part, ctx->tickets, out);
    AGRS_LAUNCHED(ctx);
    return AGRS_OK;
  }
  const bool vec = (inner % 4 == 0) && agrs_aligned16(x);
  return launch_sum_axis_mid(ctx, SrcPlain{x}, vec, outer, len, inner, out, nullptr);
}

int agrs_relu_bwd_colsum_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, int64_t rows, int64_t cols, float* colsum) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_relu_bwd_colsum_f32");
  return fused_bwd_colsum(ctx, SrcReluBwd{x, g, dx}, "agrs_relu_bwd_colsum_f32", rows, cols, x, g, dx, colsum);
}
int agrs_sigmoid_bwd_colsum_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, int64_t rows, int64_t cols, float* colsum) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_sigmoid_bwd_colsum_f32");
  return fused_bwd_colsum(ctx, SrcSigmoidBwd{x, g, dx}, "agrs_sigmoid_bwd_colsum_f32", rows, cols, x, g, dx, colsum);
}
int agrs_mse_bwd_colsum_f32(agrs_ctx* ctx, const float* pred, const float* target, const float* upstream1, float batch, int64_t rows, int64_t cols,
                            float* out, float* colsum) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_mse_bwd_colsum_f32");
  AGRS_REQUIRE(ctx, upstream1, AGRS_ERR_INVALID, "agrs_mse_bwd_colsum_f32: null upstream");
  // batch = 2^k (8192 rows per GPU in C2 / C5): x / 2^k and x * 2^-k are the same single rounding of the same value, subnormal
  // results included, and the multiplication spares four IEEE divisions per 16-byte vector
  int ex = 0;
  const float inv_batch = (batch > 0.f && frexpf(batch, &ex) == 0.5f && ex > -100 && ex < 100) ? 1.0f / batch : 0.f;
  return fused_bwd_colsum(ctx, SrcMseBwd{pred, target, upstream1, batch, out, inv_batch}, "agrs_mse_bwd_colsum_f32", rows, cols, pred, target, out, colsum);
}

int agrs_equal_f32(agrs_ctx* ctx, const float* a, const float* b, size_t n, int* host_equal) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_equal_f32");
  AGRS_NOT_CAPTURING(ctx, "agrs_equal_f32");
  AGRS_REQUIRE(ctx, host_equal && (n == 0 || (a && b)), AGRS_ERR_INVALID, "agrs_equal_f32: null pointer");
  unsigned* flag = ctx->counter + 8;  // second cache line of the counter block
  AGRS_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(unsigned), ctx->stream));
  if (n) {
    size_t blocks = (n + kThreads - 1) / kThreads;
    size_t cap = size_t(ctx->sm_count) * 8;
    if (blocks > cap) blocks = cap;
    k_not_equal<<<int(blocks), kThreads, 0, ctx->stream>>>(a, b, n, flag);
    AGRS_LAUNCHED(ctx);
  }
  unsigned h = 0;
  AGRS_CUDA(ctx, cudaMemcpyAsync(&h, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *host_equal = h ? 0 : 1;
  return AGRS_OK;
}

int agrs_transpose_f32(agrs_ctx* ctx, const float* in, int64_t rows, int64_t cols, float* out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_transpose_f32");
  AGRS_REQUIRE(ctx, rows >= 0 && cols >= 0, AGRS_ERR_INVALID, "agrs_transpose_f32: negative extent");
  if (rows == 0 || cols == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, in && out, AGRS_ERR_INVALID, "agrs_transpose_f32: null pointer");
  dim3 grid(unsigned((cols + 31) / 32), unsigned((rows + 31) / 32));
  AGRS_REQUIRE(ctx, grid.y <= 65535, AGRS_ERR_UNSUPPORTED, "agrs_transpose_f32: too many rows");
  k_transpose<<<grid, dim3(32, 8), 0, ctx->stream>>>(in, rows, cols, out);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

}  // extern "C"
