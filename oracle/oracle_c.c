/*
 * oracle_c.c -- single-threaded C restatement of the autograd.rs hot path, op for op as the reference decomposes it.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle_np.py): a second, independent restatement used (1) to cross-check the NumPy
 * oracle and (2) as the "reference CPU path on ONE core" figure of bench.py -- the reference is single-threaded
 * (Cargo.lock has no rayon / thread-tree; ndarray -> matrixmultiply::sgemm), the NumPy oracle borrows every BLAS thread.
 * Nothing under autograd.rs_b200/ may call into this file.
 *
 * What is kept from the reference's decomposition (file:line relative to the reference root):
 *   Linear::forward   x.dot(w), then `+ b` as a second in-place pass                       operators.rs:140-144
 *   Linear::backward  w.t_clone() and x.t_clone() MATERIALISED, two dots, sum_axis(0)     operators.rs:159-161
 *   ReLU / Sigmoid    separate passes; Sigmoid backward recomputes sigmoid(x)            operators.rs:109-130, :189-225
 *   MSE               sub, square, mean as separate passes; backward (p-t)*2/B*grad       operators.rs:230-254
 *   accumulate        first touch copies, later touches add                              tensor.rs:154-163
 *   SGD               tmp = g * lr ; w -= tmp (two passes, g broadcast onto w)            optim.rs:25-34
 *   zero_grad         grad = zeros(shape of data)                                        model.rs:14-21
 *   im2col / col2im   the index maps of functional_ndarray.rs:40-138
 * sgemm: the packed-panel, cache-blocked loop nest of matrixmultiply 0.3.7 (the crate ndarray 0.15's `dot` calls for f32; it is a
 * Cargo.lock dependency, not vendored under the reference tree -- restated from its published algorithm, see oc_sgemm).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OC_API __attribute__((visibility("default")))

OC_API void oc_relu_fwd(const float* x, float* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = x[i] > 0.0f ? x[i] : 0.0f; /* functional_ndarray.rs:14-22 */
}

OC_API void oc_relu_bwd(const float* x, const float* g, float* dx, long n) {
  for (long i = 0; i < n; ++i) dx[i] = x[i] <= 0.0f ? 0.0f : g[i]; /* functional_ndarray.rs:23-34 */
}

OC_API void oc_sigmoid_fwd(const float* x, float* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = 1.0f / (1.0f + expf(-x[i])); /* operators.rs:190-199 */
}

OC_API void oc_sigmoid_bwd(const float* x, const float* g, float* dx, long n) {
  for (long i = 0; i < n; ++i) { /* operators.rs:219-223: ((1 - s) * s) * g */
    float s = 1.0f / (1.0f + expf(-x[i]));
    dx[i] = ((1.0f - s) * s) * g[i];
  }
}

OC_API float oc_mse_fwd(const float* p, const float* t, float* scratch, long n) {
  for (long i = 0; i < n; ++i) scratch[i] = p[i] - t[i];          /* &pred - &target */
  for (long i = 0; i < n; ++i) scratch[i] = scratch[i] * scratch[i]; /* powi_inplace(2) */
  float s = 0.0f;
  for (long i = 0; i < n; ++i) s += scratch[i];                   /* mean(None) = sum / len */
  return s / (float)n;
}

OC_API void oc_mse_bwd(const float* p, const float* t, float upstream, long batch, long n, float* out) {
  for (long i = 0; i < n; ++i) out[i] = ((p[i] - t[i]) * 2.0f / (float)batch) * upstream; /* operators.rs:247-253 */
}

OC_API void oc_transpose(const float* in, long rows, long cols, float* out) { /* t_clone(): storage.rs:84-88 */
  for (long r = 0; r < rows; ++r)
    for (long c = 0; c < cols; ++c) out[c * rows + r] = in[r * cols + c];
}

/* C[M,N] = A[M,K] . B[K,N], row-major, plain operands (the reference always materialises its transposes first).
 * Loop nest of matrixmultiply 0.3.7's sgemm (gemm.rs: NC = 1024, KC = 256, MC = 64; x86 AVX/FMA micro-kernel 8 x 8): B is packed
 * into 8-column panels, A into 8-row panels, the micro-kernel runs k sequentially over one KC block with one multiply-add per
 * (i, j, k), and the block's partial sums are added into C -- so C(i,j) = ((blk0) + blk1) + ... in k order. */
enum { OC_KC = 256, OC_MC = 64, OC_NC = 1024, OC_MR = 8, OC_NR = 8 };
typedef float oc_v8 __attribute__((vector_size(32), aligned(4)));

static void oc_ukr(long kc, const float* ap, const float* bp, float* c, long ldc, long mr, long nr) {
  const oc_v8 z = {0, 0, 0, 0, 0, 0, 0, 0};
  oc_v8 c0 = z, c1 = z, c2 = z, c3 = z, c4 = z, c5 = z, c6 = z, c7 = z; /* one register per row of the 8 x 8 tile */
  for (long k = 0; k < kc; ++k) {
    const oc_v8 b = *(const oc_v8*)(bp + k * OC_NR);
    const float* a = ap + k * OC_MR;
    c0 += a[0] * b; c1 += a[1] * b; c2 += a[2] * b; c3 += a[3] * b;
    c4 += a[4] * b; c5 += a[5] * b; c6 += a[6] * b; c7 += a[7] * b;
  }
  const oc_v8 acc[OC_MR] = {c0, c1, c2, c3, c4, c5, c6, c7};
  if (mr == OC_MR && nr == OC_NR) {
    for (int i = 0; i < OC_MR; ++i) *(oc_v8*)(c + i * ldc) += acc[i];
  } else {
    for (long i = 0; i < mr; ++i)
      for (long j = 0; j < nr; ++j) c[i * ldc + j] += acc[i][j];
  }
}

OC_API void oc_sgemm(long M, long N, long K, const float* A, long lda, const float* B, long ldb, float* C, long ldc) {
  for (long i = 0; i < M; ++i) memset(C + i * ldc, 0, (size_t)N * sizeof(float));
  float* bpack = (float*)malloc((size_t)OC_KC * (OC_NC + OC_NR) * sizeof(float));
  float* apack = (float*)malloc((size_t)OC_KC * (OC_MC + OC_MR) * sizeof(float));
  for (long j0 = 0; j0 < N; j0 += OC_NC) {
    const long nc = N - j0 < OC_NC ? N - j0 : OC_NC;
    for (long k0 = 0; k0 < K; k0 += OC_KC) {
      const long kc = K - k0 < OC_KC ? K - k0 : OC_KC;
      for (long jp = 0; jp < nc; jp += OC_NR) { /* pack B: panel jp holds kc rows of 8 columns, zero padded */
        float* dst = bpack + jp * kc;
        const long nr = nc - jp < OC_NR ? nc - jp : OC_NR;
        for (long k = 0; k < kc; ++k)
          for (long j = 0; j < OC_NR; ++j) dst[k * OC_NR + j] = j < nr ? B[(k0 + k) * ldb + j0 + jp + j] : 0.0f;
      }
      for (long i0 = 0; i0 < M; i0 += OC_MC) {
        const long mc = M - i0 < OC_MC ? M - i0 : OC_MC;
        for (long ip = 0; ip < mc; ip += OC_MR) { /* pack A: panel ip holds kc columns of 8 rows */
          float* dst = apack + ip * kc;
          const long mr = mc - ip < OC_MR ? mc - ip : OC_MR;
          for (long k = 0; k < kc; ++k)
            for (long i = 0; i < OC_MR; ++i) dst[k * OC_MR + i] = i < mr ? A[(i0 + ip + i) * lda + k0 + k] : 0.0f;
        }
        for (long jp = 0; jp < nc; jp += OC_NR)
          for (long ip = 0; ip < mc; ip += OC_MR)
            oc_ukr(kc, apack + ip * kc, bpack + jp * kc, C + (i0 + ip) * ldc + j0 + jp, ldc,
                   mc - ip < OC_MR ? mc - ip : OC_MR, nc - jp < OC_NR ? nc - jp : OC_NR);
      }
    }
  }
  free(apack);
  free(bpack);
}

OC_API void oc_add_rowvec(float* y, const float* b, long rows, long cols) { /* Tensor + &Tensor in place, bias broadcast */
  for (long r = 0; r < rows; ++r)
    for (long c = 0; c < cols; ++c) y[r * cols + c] += b[c];
}

OC_API void oc_colsum(const float* g, long rows, long cols, float* out) { /* sum_axis(Axis(0)): res = res + row, row by row */
  for (long c = 0; c < cols; ++c) out[c] = 0.0f;
  for (long r = 0; r < rows; ++r)
    for (long c = 0; c < cols; ++c) out[c] += g[r * cols + c];
}

OC_API void oc_sgd(float* w, const float* g, float lr, long n, long ng, float* tmp) { /* optim.rs:31-33 */
  for (long i = 0; i < ng; ++i) tmp[i] = g[i] * lr;
  for (long i = 0; i < n; ++i) w[i] -= tmp[i % ng];
}

/* col[b, y*Wo + x, c*k*k + i*k + j] = im[b, c, y*s - p + i, x*s - p + j] or 0   (functional_ndarray.rs:40-104) */
OC_API void oc_im2col(const float* im, long B, long C, long H, long W, long k, long s, long p, float* col) {
  const long Ho = (H - k + 2 * p) / s + 1, Wo = (W - k + 2 * p) / s + 1, K = C * k * k;
  for (long b = 0; b < B; ++b)
    for (long y = 0; y < Ho; ++y)
      for (long x = 0; x < Wo; ++x)
        for (long c = 0; c < C; ++c)
          for (long i = 0; i < k; ++i)
            for (long j = 0; j < k; ++j) {
              const long h = y * s - p + i, w = x * s - p + j;
              col[(b * Ho * Wo + y * Wo + x) * K + c * k * k + i * k + j] =
                  (h >= 0 && h < H && w >= 0 && w < W) ? im[((b * C + c) * H + h) * W + w] : 0.0f;
            }
}

/* overlap-add in increasing patch order (functional_ndarray.rs:107-138), per sample */
OC_API void oc_col2im(const float* col, long B, long C, long Ho, long Wo, long k, long s, float* im) {
  const long Hi = (Ho - 1) * s + k, Wi = (Wo - 1) * s + k, K = C * k * k;
  memset(im, 0, (size_t)(B * C * Hi * Wi) * sizeof(float));
  for (long b = 0; b < B; ++b)
    for (long pp = 0; pp < Ho * Wo; ++pp) {
      const long y = (pp / Wo) * s, x = (pp % Wo) * s;
      for (long c = 0; c < C; ++c)
        for (long i = 0; i < k; ++i)
          for (long j = 0; j < k; ++j) im[((b * C + c) * Hi + y + i) * Wi + x + j] += col[(b * Ho * Wo + pp) * K + c * k * k + i * k + j];
    }
}

/* One training iteration of an L-layer MLP exactly as examples/fit_sine.rs:64-93 runs it through the operators above:
 * forward (Linear [+ act between layers]) -> MSE -> backward -> SGD -> zero_grad.  dims[0..L] are the layer widths,
 * W[l] is [dims[l], dims[l+1]], b[l] is [1, dims[l+1]].  act: 0 none, 1 ReLU, 2 Sigmoid.  Returns the loss.
 * Every temporary the reference allocates is allocated here too (the transposed copies, the per-op outputs, tmp = g*lr). */
OC_API float oc_mlp_step(int L, const long* dims, float** W, float** b, const float* x, const float* t, long B, int act, float lr) {
  float** z = (float**)calloc((size_t)L, sizeof(float*)); /* pre-activations */
  float** a = (float**)calloc((size_t)L + 1, sizeof(float*)); /* activations; a[0] = x */
  a[0] = (float*)x;
  for (int l = 0; l < L; ++l) {
    const long I = dims[l], O = dims[l + 1];
    z[l] = (float*)malloc((size_t)(B * O) * sizeof(float));
    oc_sgemm(B, O, I, a[l], I, W[l], O, z[l], O);
    oc_add_rowvec(z[l], b[l], B, O);
    if (act && l + 1 < L) {
      a[l + 1] = (float*)malloc((size_t)(B * O) * sizeof(float));
      if (act == 1) oc_relu_fwd(z[l], a[l + 1], B * O);
      else oc_sigmoid_fwd(z[l], a[l + 1], B * O);
    } else {
      a[l + 1] = z[l];
    }
  }
  const long D = dims[L];
  float* scratch = (float*)malloc((size_t)(B * D) * sizeof(float));
  const float loss = oc_mse_fwd(a[L], t, scratch, B * D);
  float* g = (float*)malloc((size_t)(B * D) * sizeof(float));
  oc_mse_bwd(a[L], t, 1.0f, B, B * D, g); /* seed = ones_like(loss) */
  free(scratch);
  for (int l = L - 1; l >= 0; --l) {
    const long I = dims[l], O = dims[l + 1];
    if (act && l + 1 < L) { /* activation backward into a fresh tensor */
      float* gz = (float*)malloc((size_t)(B * O) * sizeof(float));
      if (act == 1) oc_relu_bwd(z[l], g, gz, B * O);
      else oc_sigmoid_bwd(z[l], g, gz, B * O);
      free(g);
      g = gz;
    }
    float* wt = (float*)malloc((size_t)(I * O) * sizeof(float)); /* w.t_clone() */
    oc_transpose(W[l], I, O, wt);
    float* dx = (float*)malloc((size_t)(B * I) * sizeof(float));
    oc_sgemm(B, I, O, g, O, wt, I, dx, I);
    free(wt);
    float* xt = (float*)malloc((size_t)(B * I) * sizeof(float)); /* x.t_clone() */
    oc_transpose(a[l], B, I, xt);
    float* dw = (float*)malloc((size_t)(I * O) * sizeof(float));
    oc_sgemm(I, O, B, xt, B, g, O, dw, O);
    free(xt);
    float* db = (float*)malloc((size_t)O * sizeof(float));
    oc_colsum(g, B, O, db);
    /* accumulate into zero-initialised grads (zero_grad made them zeros(shape of data)), then SGD, then zero_grad again */
    float* gw = (float*)calloc((size_t)(I * O), sizeof(float));
    float* gb = (float*)calloc((size_t)O, sizeof(float));
    for (long i = 0; i < I * O; ++i) gw[i] += dw[i];
    for (long i = 0; i < O; ++i) gb[i] += db[i];
    float* tmp = (float*)malloc((size_t)(I * O) * sizeof(float));
    oc_sgd(W[l], gw, lr, I * O, I * O, tmp);
    oc_sgd(b[l], gb, lr, O, O, tmp);
    free(tmp);
    free(gw);
    free(gb);
    free(dw);
    free(db);
    free(g);
    g = dx; /* upstream of the previous layer; for l == 0 this is x.grad, computed as the reference does and dropped */
  }
  free(g);
  for (int l = 0; l < L; ++l) {
    if (a[l + 1] != z[l]) free(a[l + 1]);
    free(z[l]);
  }
  free(z);
  free(a);
  return loss;
}
