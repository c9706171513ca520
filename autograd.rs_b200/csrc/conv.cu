// Conv2D lowering kernels: im2col, col2im (gather form, no atomics), filter broadcast.
// Index maps follow src/operators/functional_ndarray.rs:40-138 of the reference exactly;
// both kernels are pure data movement (HBM-bound): one thread per output element, writes coalesced.
#include "agrs_internal.h"

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

__global__ void k_broadcast_filter(const float* __restrict__ filt, int64_t per, int64_t total, float* __restrict__ out) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) out[i] = filt[i % per];
}

unsigned grid1d(agrs_ctx* ctx, int64_t total) {
  int64_t blocks = (total + 255) / 256, cap = int64_t(ctx->sm_count) * 16;
  return unsigned(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

extern "C" {

int agrs_im2col_f32(agrs_ctx* ctx, const float* im, int64_t B, int64_t C, int64_t H, int64_t W, int64_t k, int64_t stride,
                    int64_t pad, float* col) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, B >= 0 && C >= 0 && H >= 0 && W >= 0 && k > 0 && stride > 0 && pad >= 0, AGRS_ERR_INVALID, "agrs_im2col_f32: bad geometry");
  AGRS_REQUIRE(ctx, H + 2 * pad >= k && W + 2 * pad >= k, AGRS_ERR_SHAPE, "agrs_im2col_f32: kernel %lld larger than padded image %lldx%lld",
               (long long)k, (long long)(H + 2 * pad), (long long)(W + 2 * pad));
  int64_t Ho = (H - k + 2 * pad) / stride + 1, Wo = (W - k + 2 * pad) / stride + 1;
  int64_t total = B * Ho * Wo * C * k * k;
  if (total == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, im && col, AGRS_ERR_INVALID, "agrs_im2col_f32: null pointer");
  AGRS_REQUIRE(ctx, C * k * k < (1LL << 31) && Ho * Wo < (1LL << 31), AGRS_ERR_UNSUPPORTED, "agrs_im2col_f32: extent overflow");
  k_im2col<<<grid1d(ctx, total), 256, 0, ctx->stream>>>(im, total, int(C), int(H), int(W), int(k), int(stride), int(pad), int(Ho), int(Wo), col);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_col2im_f32(agrs_ctx* ctx, const float* col, int64_t B, int64_t C, int64_t Ho, int64_t Wo, int64_t k, int64_t stride,
                    float* im_out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, B >= 0 && C >= 0 && Ho > 0 && Wo > 0 && k > 0 && stride > 0, AGRS_ERR_INVALID, "agrs_col2im_f32: bad geometry");
  int64_t Hi = (Ho - 1) * stride + k, Wi = (Wo - 1) * stride + k;
  int64_t total = B * C * Hi * Wi;
  if (total == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, col && im_out, AGRS_ERR_INVALID, "agrs_col2im_f32: null pointer");
  k_col2im<<<grid1d(ctx, total), 256, 0, ctx->stream>>>(col, total, int(C), int(Hi), int(Wi), int(k), int(stride), int(Ho), int(Wo), im_out);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_broadcast_filter_f32(agrs_ctx* ctx, const float* filt, int64_t kk, int64_t c_out, int64_t c_in, float* out) {
  AGRS_CHECK_CTX(ctx);
  AGRS_REQUIRE(ctx, kk >= 0 && c_out >= 0 && c_in >= 0, AGRS_ERR_INVALID, "agrs_broadcast_filter_f32: negative extent");
  int64_t per = kk * c_out, total = per * c_in;
  if (total == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, filt && out, AGRS_ERR_INVALID, "agrs_broadcast_filter_f32: null pointer");
  k_broadcast_filter<<<unsigned((total + 255) / 256), 256, 0, ctx->stream>>>(filt, per, total, out);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

}  // extern "C"
