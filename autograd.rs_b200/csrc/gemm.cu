// agrs_gemm_f32: argument checking and path selection for the two GEMM back ends.
#include "agrs_internal.h"
#include "act.cuh"
#include "mag.cuh"

static cudaEvent_t take_event(agrs_ctx* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return e;
}

// live GEMM timing (agrs_profile_*): an event pair around one launch, shared with agrs_gemm_f32_scatter (gemm_tcgen05_2cta.cu)
int agrs_prof_begin(agrs_ctx* ctx, int path, double flops, agrs_ctx::ProfRec* rec) {
  *rec = agrs_ctx::ProfRec{};
  if (!ctx->profiling) return AGRS_OK;
  rec->path = path;
  rec->flops = flops;
  rec->start = take_event(ctx);
  rec->stop = take_event(ctx);
  AGRS_REQUIRE(ctx, rec->start && rec->stop, AGRS_ERR_CUDA, "agrs_gemm_f32: cannot create timing events");
  AGRS_CUDA(ctx, cudaEventRecord(rec->start, ctx->stream));
  return AGRS_OK;
}
int agrs_prof_end(agrs_ctx* ctx, const agrs_ctx::ProfRec& rec) {
  if (!rec.start) return AGRS_OK;
  AGRS_CUDA(ctx, cudaEventRecord(rec.stop, ctx->stream));
  ctx->prof.push_back(rec);
  return AGRS_OK;
}

// second pass of agrs_gemm_bias_act_f32 for outputs with padded rows (the contiguous case runs the elementwise kernels)
static __global__ void __launch_bounds__(256) k_act_rows(const float* __restrict__ Z, int64_t ldz, float* __restrict__ Y, int64_t ldy, int64_t M,
                                                         int64_t N, int kind, uint32_t* mx) {
  uint32_t mag = 0;
  const int64_t total = M * N, stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / N, n = i - m * N;
    const float z = Z[m * ldz + n];
    const float y = kind == AGRS_ACT_RELU ? agrs_act_relu(z) : agrs_act_sigmoid(z);
    Y[m * ldy + n] = y;
    mag = max(mag, mag_bits(y));
  }
  if (mx) mag_commit(mag, mx);
}

extern "C" {

int agrs_gemm_bias_act_f32(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                           const float* B, int64_t ldb, float* Z, int64_t ldz, const float* bias, int act, float* Y, int64_t ldy) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_gemm_bias_act_f32");
  AGRS_REQUIRE(ctx, act == AGRS_ACT_RELU || act == AGRS_ACT_SIGMOID, AGRS_ERR_INVALID, "agrs_gemm_bias_act_f32: unknown activation %d", act);
  AGRS_REQUIRE(ctx, M <= 0 || N <= 0 || (Y && ldy >= N), AGRS_ERR_INVALID, "agrs_gemm_bias_act_f32: null Y or ldy < N");
  AGRS_REQUIRE(ctx, Y != Z, AGRS_ERR_INVALID, "agrs_gemm_bias_act_f32: Y must not alias Z (the pre-activation stays a tensor of its own)");
  ctx->gemm_act.kind = act;
  ctx->gemm_act.y = Y;
  ctx->gemm_act.ldy = ldy;
  ctx->gemm_act.fused = false;
  const int rc = agrs_gemm_f32(ctx, transA, transB, M, N, K, A, lda, B, ldb, Z, ldz, bias);
  const bool fused = ctx->gemm_act.fused;
  ctx->gemm_act = agrs_ctx::GemmAct{};
  ctx->last_act_fused = fused ? 1 : 0;
  if (rc) {
    ctx->next_out_mx = nullptr;
    return rc;
  }
  if (fused || M <= 0 || N <= 0) {
    if (!fused) ctx->next_out_mx = nullptr;  // nothing was written
    return AGRS_OK;
  }
  // the back end that ran could not write Y from its epilogue: the activation as a second pass over Z
  if (ldz == N && ldy == N)
    return act == AGRS_ACT_RELU ? agrs_relu_fwd_f32(ctx, Z, Y, size_t(M) * size_t(N)) : agrs_sigmoid_fwd_f32(ctx, Z, Y, size_t(M) * size_t(N));
  uint32_t* mx = nullptr;
  if (int trc = agrs_take_mx(ctx, &mx)) return trc;
  const int64_t blocks = (M * N + 255) / 256, cap = int64_t(ctx->sm_count) * 8;
  k_act_rows<<<unsigned(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(Z, ldz, Y, ldy, M, N, act, mx);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}

int agrs_last_gemm_act_fused(agrs_ctx* ctx) { return ctx ? ctx->last_act_fused : 0; }

int agrs_set_tuning(agrs_ctx* ctx, int key, int64_t value) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_set_tuning");
  switch (key) {
    case AGRS_TUNE_TC_BN:
      AGRS_REQUIRE(ctx, value == 128 || value == 256, AGRS_ERR_INVALID, "AGRS_TUNE_TC_BN must be 128 or 256");
      ctx->tc_bn = int(value);
      return AGRS_OK;
    case AGRS_TUNE_TC_EXPLICIT_HI:
      ctx->tc_explicit_hi = value ? 1 : 0;
      return AGRS_OK;
    case AGRS_TUNE_TC_KCHUNK:
      AGRS_REQUIRE(ctx, value >= 0 && value < (1LL << 30), AGRS_ERR_INVALID, "AGRS_TUNE_TC_KCHUNK out of range");
      ctx->tc_kchunk = int(value);
      return AGRS_OK;
    case AGRS_TUNE_TC_SPLITK:
      AGRS_REQUIRE(ctx, value >= 0 && value <= 64, AGRS_ERR_INVALID, "AGRS_TUNE_TC_SPLITK out of range");
      ctx->tc_splitk = int(value);
      return AGRS_OK;
    case AGRS_TUNE_TC_2CTA:
      ctx->tc_2cta = value ? 1 : 0;
      return AGRS_OK;
    case AGRS_TUNE_TC_CROSS:
      AGRS_REQUIRE(ctx, value >= 0 && value <= 5, AGRS_ERR_INVALID, "AGRS_TUNE_TC_CROSS must be 0 .. 5");
      ctx->tc_cross = int(value);
      return AGRS_OK;
    case AGRS_TUNE_TC_EPI_TMA:
      ctx->tc_epi_tma = value ? 1 : 0;
      return AGRS_OK;
    case AGRS_TUNE_TC_REUSE_SCALES:
      ctx->tc_reuse_scales = value ? 1 : 0;
      return AGRS_OK;
    case AGRS_TUNE_TC_WAVE_SYNC:
      AGRS_REQUIRE(ctx, value >= 0 && value <= 2, AGRS_ERR_INVALID, "AGRS_TUNE_TC_WAVE_SYNC must be 0, 1 or 2");
      ctx->tc_wave_sync = int(value);
      return AGRS_OK;
    case AGRS_TUNE_TC_MIN_MACS:
      ctx->tc_min_flops = value;
      return AGRS_OK;
    case AGRS_TUNE_TC_ACT_WARPS:
      ctx->tc_act_warps = value ? 1 : 0;
      return AGRS_OK;
  }
  return agrs_fail(ctx, AGRS_ERR_INVALID, "unknown tuning key %d", key);
}

int agrs_get_tuning(agrs_ctx* ctx, int key, int64_t* value) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_get_tuning");
  AGRS_REQUIRE(ctx, value, AGRS_ERR_INVALID, "agrs_get_tuning: null out");
  switch (key) {
    case AGRS_TUNE_TC_BN: *value = ctx->tc_bn; return AGRS_OK;
    case AGRS_TUNE_TC_EXPLICIT_HI: *value = ctx->tc_explicit_hi; return AGRS_OK;
    case AGRS_TUNE_TC_KCHUNK: *value = ctx->tc_kchunk; return AGRS_OK;
    case AGRS_TUNE_TC_SPLITK: *value = ctx->tc_splitk; return AGRS_OK;
    case AGRS_TUNE_TC_2CTA: *value = ctx->tc_2cta; return AGRS_OK;
    case AGRS_TUNE_TC_CROSS: *value = ctx->tc_cross; return AGRS_OK;
    case AGRS_TUNE_TC_EPI_TMA: *value = ctx->tc_epi_tma; return AGRS_OK;
    case AGRS_TUNE_TC_REUSE_SCALES: *value = ctx->tc_reuse_scales; return AGRS_OK;
    case AGRS_TUNE_TC_WAVE_SYNC: *value = ctx->tc_wave_sync; return AGRS_OK;
    case AGRS_TUNE_TC_MIN_MACS: *value = ctx->tc_min_flops; return AGRS_OK;
    case AGRS_TUNE_TC_ACT_WARPS: *value = ctx->tc_act_warps; return AGRS_OK;
  }
  return agrs_fail(ctx, AGRS_ERR_INVALID, "unknown tuning key %d", key);
}

int agrs_gemm_f32(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                  const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_gemm_f32");
  // magnitudes announced for this call (agrs_gemm_operand_absmax): consumed here whatever path the call takes
  const uint32_t *a_mx = ctx->gemm_a_mx, *b_mx = ctx->gemm_b_mx;
  ctx->gemm_a_mx = ctx->gemm_b_mx = nullptr;
  AGRS_REQUIRE(ctx, M >= 0 && N >= 0 && K >= 0, AGRS_ERR_INVALID, "agrs_gemm_f32: negative extent");
  if (M == 0 || N == 0) return AGRS_OK;
  AGRS_REQUIRE(ctx, C && (K == 0 || (A && B)), AGRS_ERR_INVALID, "agrs_gemm_f32: null pointer");
  AGRS_REQUIRE(ctx, lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, AGRS_ERR_SHAPE,
               "agrs_gemm_f32: leading dimension too small (lda=%lld ldb=%lld ldc=%lld for M=%lld N=%lld K=%lld tA=%d tB=%d)",
               (long long)lda, (long long)ldb, (long long)ldc, (long long)M, (long long)N, (long long)K, transA, transB);
  const bool eligible = K > 0 && agrs_gemm_tc_eligible(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias);
  bool use_tc;
  if (ctx->gemm_path == AGRS_GEMM_TCGEN05) {
    AGRS_REQUIRE(ctx, eligible, AGRS_ERR_UNSUPPORTED,
                 "agrs_gemm_f32: tcgen05 path forced but operands are not TMA-mappable (16-byte aligned bases, leading dims %% 4 == 0)");
    use_tc = true;
  } else if (ctx->gemm_path == AGRS_GEMM_SIMT) {
    use_tc = false;
  } else {
    // a 128xBN tile per SM: tiny or very skinny problems leave the tensor cores idle and pay the pipeline fill
    use_tc = eligible && M >= 64 && N >= 32 && K >= 32 && (M * N) * K >= ctx->tc_min_flops;
  }
  const bool skinny_ok = K > 0 && agrs_gemm_skinny_eligible(transA, transB, M, N, K);
  bool use_skinny = false;
  if (ctx->gemm_path == AGRS_GEMM_SKINNY) {
    AGRS_REQUIRE(ctx, skinny_ok, AGRS_ERR_UNSUPPORTED, "agrs_gemm_f32: skinny path forced but the shape is not skinny (N <= 32, see gemm_skinny.cu)");
    use_skinny = true;
    use_tc = false;
  } else if (ctx->gemm_path == AGRS_GEMM_AUTO && !use_tc) {
    use_skinny = skinny_ok;
  }
  ctx->last_gemm_path = use_tc ? AGRS_GEMM_TCGEN05 : (use_skinny ? AGRS_GEMM_SKINNY : AGRS_GEMM_SIMT);
  agrs_ctx::ProfRec rec;
  if (int prc = agrs_prof_begin(ctx, ctx->last_gemm_path, 2.0 * double(M) * double(N) * double(K), &rec)) return prc;
  const bool pair = use_tc && ctx->tc_2cta && M > 128;
  if (pair) {
    ctx->gemm_a_mx = a_mx;
    ctx->gemm_b_mx = b_mx;
  }
  int rc = use_skinny ? agrs_gemm_skinny(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias)
         : pair   ? agrs_gemm_tc_2cta(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias)
         : use_tc ? agrs_gemm_tc(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias)
                  : agrs_gemm_simt(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias);
  ctx->gemm_a_mx = ctx->gemm_b_mx = nullptr;
  if (int prc = agrs_prof_end(ctx, rec)) return prc;
  return rc;
}

int agrs_gemm_operand_absmax(agrs_ctx* ctx, const uint32_t* a_bits, const uint32_t* b_bits) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_gemm_operand_absmax");
  ctx->gemm_a_mx = a_bits;
  ctx->gemm_b_mx = b_bits;
  return AGRS_OK;
}

int agrs_gemm_wants_absmax(agrs_ctx* ctx, int64_t M, int64_t N, int64_t K) {
  if (!ctx || ctx->sticky || (ctx->tc_cross != 4 && ctx->tc_cross != 5) || !ctx->tc_2cta) return 0;
  if (ctx->gemm_path == AGRS_GEMM_TCGEN05) return M > 128;
  return ctx->gemm_path == AGRS_GEMM_AUTO && M > 128 && N >= 32 && K >= 32 && (M * N) * K >= ctx->tc_min_flops;
}

int agrs_profile_enable(agrs_ctx* ctx, int on) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_profile_enable");
  AGRS_NOT_CAPTURING(ctx, "agrs_profile_enable");
  ctx->profiling = on != 0;
  return AGRS_OK;
}

int agrs_profile_gemm(agrs_ctx* ctx, int path, uint64_t* launches, double* total_ms, double* total_flops) {
  AGRS_CHECK_CTX(ctx);
  AGRS_RANGE("agrs_profile_gemm");
  AGRS_NOT_CAPTURING(ctx, "agrs_profile_gemm");
  AGRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t n = 0;
  double ms = 0.0, fl = 0.0;
  std::vector<agrs_ctx::ProfRec> keep;
  for (auto& r : ctx->prof) {
    if (r.path != path) {
      keep.push_back(r);
      continue;
    }
    float t = 0.f;
    AGRS_CUDA(ctx, cudaEventElapsedTime(&t, r.start, r.stop));
    ms += t;
    fl += r.flops;
    ++n;
    ctx->event_pool.push_back(r.start);
    ctx->event_pool.push_back(r.stop);
  }
  ctx->prof.swap(keep);
  if (launches) *launches = n;
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  return AGRS_OK;
}

}  // extern "C"
