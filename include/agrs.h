/*
 * agrs.h -- C ABI of libagrs_b200.so: the B200 (sm_100a) device arm of autograd.rs.
 *
 * This is the surface a thin Rust `extern "C"` (-sys) crate binds.  Every entry point replaces
 * one `StorageType::CudaData` arm that is `todo!()` in the reference (NickLucche/autograd.rs);
 * the reference call site is cited next to each declaration (paths relative to the reference
 * root).  Plain pointers and sizes only: no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - all tensors are f32, row-major ("standard layout"), device pointers unless named host_*;
 *   - every function returns 0 (AGRS_OK) or an error class; the message is kept per context
 *     (agrs_last_error).  Nothing unwinds or aborts across the boundary; the host wrapper turns
 *     a non-zero code into the reference's panic message;
 *   - every call is asynchronous on the context's stream; only agrs_download, agrs_sync,
 *     agrs_equal_f32 and agrs_ctx_destroy block;
 *   - one context per device, driven by one host thread (the reference's Tensor is Rc<RefCell>,
 *     i.e. !Send: src/utils/mod.rs:5, src/autograd/autograd.rs:9).
 *   - there is NO CPU fallback: on a machine without an sm_100 device agrs_ctx_create fails.
 */
#ifndef AGRS_H_
#define AGRS_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define AGRS_API __attribute__((visibility("default")))
#else
#define AGRS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agrs_ctx agrs_ctx;

enum agrs_status {
  AGRS_OK = 0,
  AGRS_ERR_INVALID = 1,     /* bad argument (null pointer, negative size, unknown enum) */
  AGRS_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed; sticky per context */
  AGRS_ERR_SHAPE = 3,       /* shapes do not conform / cannot broadcast */
  AGRS_ERR_UNSUPPORTED = 4, /* valid request this build cannot serve (e.g. forced tcgen05 path on unaligned operands) */
  AGRS_ERR_NODEVICE = 5     /* no CUDA device of compute capability 10.x */
};

/* ---- context, memory (replaces CudaData{ptr} + Drop, src/tensor/storage.rs:5-14) ---- */
AGRS_API int agrs_ctx_create(int device, agrs_ctx** out);
/* Same, but enqueue on an existing cudaStream_t (e.g. the one torch.distributed/NCCL uses). */
AGRS_API int agrs_ctx_create_on_stream(int device, void* cuda_stream, agrs_ctx** out);
AGRS_API int agrs_ctx_destroy(agrs_ctx* ctx);
AGRS_API const char* agrs_last_error(agrs_ctx* ctx);
AGRS_API int agrs_sync(agrs_ctx* ctx);
AGRS_API void* agrs_stream(agrs_ctx* ctx); /* the cudaStream_t launches go to */
AGRS_API int agrs_device_info(agrs_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
AGRS_API uint64_t agrs_launch_count(agrs_ctx* ctx);

AGRS_API int agrs_alloc(agrs_ctx* ctx, size_t nbytes, void** out); /* stream-ordered pool; 256-B aligned */
/* Back the pool with `nbytes` of device memory now (allocate + free once; the pool never returns memory to the driver), so that
 * the first training steps do not pay for pool growth -- a driver call that takes milliseconds per GB -- inside a timed region. */
AGRS_API int agrs_pool_reserve(agrs_ctx* ctx, size_t nbytes);
AGRS_API int agrs_pool_stats(agrs_ctx* ctx, size_t* reserved_bytes, size_t* used_bytes, size_t* reserved_high, size_t* used_high);
/* Hand the pool's unused blocks back to the driver (blocks: synchronises the stream first) and reset its high-water marks: between
 * two workloads of very different footprint, so that the second starts from the pool a fresh process would have. */
AGRS_API int agrs_pool_trim(agrs_ctx* ctx);
AGRS_API int agrs_free(agrs_ctx* ctx, void* ptr);                  /* stream-ordered: safe with work in flight */
AGRS_API int agrs_host_alloc(agrs_ctx* ctx, size_t nbytes, void** out); /* pinned host staging */
AGRS_API int agrs_host_free(agrs_ctx* ctx, void* ptr);
AGRS_API int agrs_upload(agrs_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes);   /* Tensor::from, tensor.rs:37-61 */
AGRS_API int agrs_download(agrs_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes); /* data()/Index/into_raw_vec, storage.rs:139-177; blocks */
/* Upload on a separate COPY stream so it overlaps kernels already running on the context's stream (input prefetch for the next
 * training step).  Ordered after everything enqueued on the context's stream at call time (earlier readers of dst are safe);
 * agrs_prefetch_join makes the context's stream wait for all prefetches issued so far.  src_host must be pinned and stay alive. */
AGRS_API int agrs_upload_prefetch(agrs_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes);
AGRS_API int agrs_prefetch_join(agrs_ctx* ctx);
typedef struct agrs_event agrs_event;
/* The same copy with its own completion event (agrs_event_create, below): agrs_wait_event makes the context's stream -- not the
 * host -- wait for THAT copy only, so a consumer of the first tensor starts while the second is still on the bus. */
AGRS_API int agrs_upload_prefetch_event(agrs_ctx* ctx, void* dst_dev, const void* src_host, size_t nbytes, agrs_event* landed);
AGRS_API int agrs_wait_event(agrs_ctx* ctx, agrs_event* ev);
/* Read-back that does not stall the pipeline: the copy is enqueued on the context's stream behind the work that produces src_dev
 * and the call returns; agrs_event_wait blocks the host until THAT copy has landed, not until the stream is empty.  A training
 * loop reads step i's loss after it has enqueued step i+1 (examples/fit_sine.rs:64-93 reads `loss` every epoch): the device never
 * idles while the host walks the next step's graph.  dst_host must be pinned (agrs_host_alloc) and stay alive until the wait. */
AGRS_API int agrs_event_create(agrs_ctx* ctx, agrs_event** out);
AGRS_API int agrs_event_destroy(agrs_ctx* ctx, agrs_event* ev);
AGRS_API int agrs_download_async(agrs_ctx* ctx, void* dst_host, const void* src_dev, size_t nbytes, agrs_event* done);
AGRS_API int agrs_event_wait(agrs_ctx* ctx, agrs_event* ev);                                     /* blocks the host */
AGRS_API int agrs_copy(agrs_ctx* ctx, void* dst_dev, const void* src_dev, size_t nbytes);      /* to_owned() deep copy, tensor.rs:148,160 */
AGRS_API int agrs_fill_f32(agrs_ctx* ctx, float* x, float value, size_t n);                     /* storage.rs:75; zeros/ones tensor.rs:67-84; zero_grad tensor.rs:135-140 */

/* ---- GEMM (replaces Tensor::dot's CUDA arm, src/tensor/tensor.rs:319) ----
 * C[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[N] broadcast over rows), row-major, leading dims in elements.
 * transA=0: A is [M,K] (lda>=K); transA=1: A is stored [K,M] (lda>=M)   -- dW = X^T G (operators.rs:160)
 * transB=0: B is [K,N] (ldb>=N); transB=1: B is stored [N,K] (ldb>=K)   -- dX = G W^T (operators.rs:159)
 * so the reference's materialised t_clone() copies (storage.rs:84-88) never exist on the device.
 * bias != NULL fuses `+ b` of Linear::forward / Conv2D::forward (operators.rs:144, :379) into the epilogue.
 * Large aligned shapes run the TMA + tcgen05 split kernel (f32-faithful: with announced operand magnitudes x * 2^s = h + l in
 * fp16 and three fp16 products; otherwise x = hi + lo with hi the tf32 part, hi*hi on the tf32 tensor pipe and the two
 * correction terms lo*x + x*lo on bf16 operands; finite inputs: <= 1e-5 of max|C|; NaN and
 * zero propagate as in f32; an inf INPUT -- or a finite one above bf16's largest value, 3.39e38, the top 0.4 % of f32's last
 * binade -- yields a non-finite result that is NaN where f32 gives inf; sums folded across K chunks longer than
 * AGRS_TUNE_TC_KCHUNK flush subnormal results to zero);
 * skinny outputs (N <= 32: the Conv2D lowering, Linear(3456,10)) run HBM-bound streaming kernels; everything else runs the
 * tiled SIMT f32 kernel.  All are CUDA; the choice never changes semantics. */
AGRS_API int agrs_gemm_f32(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K,
                  const float* A, int64_t lda, const float* B, int64_t ldb,
                  float* C, int64_t ldc, const float* bias);
/* Linear::forward followed by an activation (the Layer loop of src/nn/layers.rs:24-43 feeding operators.rs:140-144 into
 * :112-116 / :195-200) as ONE call: Z = op(A) op(B) + bias (the pre-activation -- it stays a tensor of its own because the
 * activation's graph node keeps it as its variable, operators.rs:120-129, :214-224) and Y = act(Z), both [M,N].  Where the GEMM
 * kernel produces the final value of an element itself it can also write the activation: the SIMT kernel does, from the register
 * that holds the sum (small models are launch-bound: one launch fewer per layer); the pair kernel does under
 * AGRS_TUNE_TC_ACT_WARPS = 1 -- once the K chunks of a tile have been folded into Z, two otherwise idle warps fetch the tile back
 * from L2 and store act(Z) while the tensor cores are on the next tile.  Everywhere else (pair kernel by default, K split over CTA
 * pairs, one-CTA and skinny kernels, outputs that are not 16-byte aligned) the activation runs as a second
 * pass inside this call.  Either way Z and Y carry the bits of agrs_gemm_f32 followed by agrs_relu_fwd_f32 / agrs_sigmoid_fwd_f32.
 * Honours agrs_gemm_operand_absmax (operands) and agrs_next_output_absmax (reports the magnitude of Y). */
enum agrs_act { AGRS_ACT_NONE = 0, AGRS_ACT_RELU = 1, AGRS_ACT_SIGMOID = 2 };
AGRS_API int agrs_gemm_bias_act_f32(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K,
                  const float* A, int64_t lda, const float* B, int64_t ldb,
                  float* Z, int64_t ldz, const float* bias, int act, float* Y, int64_t ldy);
AGRS_API int agrs_last_gemm_act_fused(agrs_ctx* ctx);     /* 1 when the most recent agrs_gemm_bias_act_f32 wrote Y from the GEMM's epilogue */
enum agrs_gemm_path { AGRS_GEMM_AUTO = 0, AGRS_GEMM_SIMT = 1, AGRS_GEMM_TCGEN05 = 2, AGRS_GEMM_SKINNY = 3 /* HBM-bound kernels for N <= 32 outputs */ };
/* Operand magnitudes for the fp16 split mode (AGRS_TUNE_TC_CROSS 4): the bit pattern of the largest finite |x| of A and of B,
 * each ONE device word the caller keeps with its tensor (what CudaData would carry next to `ptr`, storage.rs:5-8).  Applies to the
 * next agrs_gemm_f32 / agrs_gemm_f32_scatter call on this context only; that GEMM then scales each operand by one exact power
 * of two and skips its own two magnitude passes.  Any value >= the true maximum (within ~2^10 of it) serves; NULLs (the default)
 * let the GEMM find per-row / per-column magnitudes itself.  Ignored by every other mode and path. */
AGRS_API int agrs_gemm_operand_absmax(agrs_ctx* ctx, const uint32_t* a_bits, const uint32_t* b_bits);
AGRS_API int agrs_gemm_wants_absmax(agrs_ctx* ctx, int64_t M, int64_t N, int64_t K); /* 1 when a GEMM of this shape would use them under the current tuning */
/* out_bits[0] = bit pattern of max finite |x[i]| (0 for n == 0): the stand-alone pass for tensors no producer reported on */
AGRS_API int agrs_absmax_f32(agrs_ctx* ctx, const float* x, size_t n, uint32_t* out_bits);
/* The NEXT elementwise launch on this context (relu / sigmoid forward and backward, mse_bwd, scalar and same-length binary ops,
 * neg, powi, sgd_step without broadcast) also reports the magnitude of what it writes into out_bits[0] -- the producer touches
 * every element anyway, so the GEMM that consumes its output needs no pass of its own.  One request serves one call. */
AGRS_API int agrs_next_output_absmax(agrs_ctx* ctx, uint32_t* out_bits);
AGRS_API int agrs_set_gemm_path(agrs_ctx* ctx, int path); /* tests/bench: force one path; TCGEN05 on unaligned operands -> AGRS_ERR_UNSUPPORTED */
AGRS_API int agrs_last_gemm_path(agrs_ctx* ctx);          /* path the most recent agrs_gemm_f32 took */
/* kernel tuning knobs (tests, benches, profiling); never change results beyond f32 rounding */
enum agrs_tuning_key {
  AGRS_TUNE_TC_BN = 0,          /* tcgen05 tile width: 256 (default) or 128 */
  AGRS_TUNE_TC_EXPLICIT_HI = 1, /* 1: write the tf32-truncated hi tile explicitly instead of relying on the MMA ignoring the low 13 bits */
  AGRS_TUNE_TC_MIN_MACS = 2,    /* AUTO path: minimum M*N*K for the tcgen05 kernel */
  AGRS_TUNE_TC_KCHUNK = 3,      /* K elements accumulated in TMEM between round-to-nearest folds into C (0 = all of K; default 1024) */
  AGRS_TUNE_TC_2CTA = 4,        /* 1 (default): cta_group::2 pair kernel (256x256 tiles) when M > 128; 0: always the one-CTA kernel */
  AGRS_TUNE_TC_SPLITK = 5,      /* pair kernel: 0 (default) = pick the K split that fills the last wave of CTA pairs; n = force n splits */
  AGRS_TUNE_TC_CROSS = 6,       /* pair kernel, the two correction terms a_lo*b + a*b_lo: 0 = tf32 operands (3xTF32), 1 / 2 = bf16 operands
                                   at twice the MMA rate (SWIZZLE_64B / unswizzled tiles); measured max error 9e-6 / 6e-6 of max|C|.
                                   3 = no tf32 pass: x = bf16 + bf16, three bf16 products (6 MMA slots per k-block instead of 8,
                                   representation error ~4e-6 instead of ~1e-6).  4 = x scaled per row of op(A) / column of op(B) by an
                                   exact power of two, then fp16 + fp16: three fp16 products, representation error ~7e-8, preceded by two
                                   small absmax kernels per launch -- unless the caller announced one magnitude per operand
                                   (agrs_gemm_operand_absmax), which then serve as the scales.  5 (default) = mode 4 for calls with
                                   announced magnitudes, mode 1 for calls without: no magnitude pass ever runs.  Env AGRS_TC_CROSS
                                   overrides the default at context creation */
  AGRS_TUNE_TC_EPI_TMA = 7,     /* pair kernel epilogue: 1 (default) = 32x32 pieces staged in shared memory, written by TMA tensor stores and
                                   folded across K chunks by TMA reduce-add; 0 = direct vector stores from registers.  Same values either way */
  AGRS_TUNE_TC_REUSE_SCALES = 8,/* cross mode 4, measurement aid: 1 = skip the absmax kernels and reuse the operand magnitudes the previous
                                   launch left in the workspace -- only meaningful when the operands have not changed.  Default 0 */
  AGRS_TUNE_TC_WAVE_SYNC = 9,   /* pair kernel: the TMA producers of a wave of tiles start together so that the CTA pairs keep sharing
                                   operand blocks through the L2 (a bounded wait, never a dependency; results do not change).  0 = off,
                                   1 = whenever there is more than one wave, 2 (default) = large problems only (from 8192^3).  Env AGRS_TC_WAVE_SYNC overrides the default at context creation */
  AGRS_TUNE_TC_ACT_WARPS = 10   /* pair kernel under agrs_gemm_bias_act_f32: 1 = two otherwise idle warps per CTA read each finished tile of Z back from
                                   L2 and write act(Z) while the tensor cores are on the next tile (no second launch, Z not re-read from HBM);
                                   0 (default) = the activation runs as a second pass after the kernel.  Same bits either way.  Off by default because
                                   it does not pay on a power-capped B200: time-neutral on 8 x Linear(8192) + Sigmoid, 1-2 % slower on
                                   4 x Linear(4096) + ReLU.  Env AGRS_TC_ACT_WARPS overrides the default at context creation */
};
AGRS_API int agrs_set_tuning(agrs_ctx* ctx, int key, int64_t value);
AGRS_API int agrs_get_tuning(agrs_ctx* ctx, int key, int64_t* value);

/* ---- whole-step CUDA graphs (launch-bound configurations: fit_sine runs ~11 tiny kernels per step) ----
 * Between agrs_capture_begin and agrs_capture_end every launch, memset, copy, stream-ordered alloc/free and NCCL call made on
 * this context is RECORDED instead of run.  The graph is then replayed with agrs_graph_launch: one driver call per `times`
 * training steps instead of one per kernel.  Rules: run at least one ordinary step first (workspaces are grown outside
 * capture); nothing that blocks (download, sync, timer_end, equal) may be called while capturing; buffers allocated before the
 * capture and released during it are released when the capture ends; buffers allocated during the capture keep their
 * addresses across replays, so tensors created in the captured step (the loss) can be read after any replay. */
typedef struct agrs_graph agrs_graph;
AGRS_API int agrs_capture_begin(agrs_ctx* ctx);
AGRS_API int agrs_capture_end(agrs_ctx* ctx, agrs_graph** out);
AGRS_API int agrs_graph_launch(agrs_ctx* ctx, agrs_graph* graph, int times);
AGRS_API int agrs_graph_destroy(agrs_ctx* ctx, agrs_graph* graph);
AGRS_API int agrs_graph_info(agrs_graph* graph, uint64_t* kernel_nodes, uint64_t* total_nodes);

/* device-side stopwatch on the context's stream (CUDA events): begin records, end records + synchronises + returns ms */
AGRS_API int agrs_timer_begin(agrs_ctx* ctx);
AGRS_API int agrs_timer_end(agrs_ctx* ctx, float* ms);
/* live kernel timing for bench.py's roofline: while enabled, every agrs_gemm_f32 launch is bracketed by a CUDA-event
 * pair on the context's stream.  agrs_profile_gemm synchronises, sums the elapsed times and algorithmic flops (2*M*N*K)
 * of the launches that took `path` since the last call, and resets them. */
AGRS_API int agrs_profile_enable(agrs_ctx* ctx, int on);
AGRS_API int agrs_profile_gemm(agrs_ctx* ctx, int path, uint64_t* launches, double* total_ms, double* total_flops);

/* ---- activations / loss ---- */
AGRS_API int agrs_relu_fwd_f32(agrs_ctx* ctx, const float* x, float* y, size_t n);                   /* operators.rs:115 -> functional_ndarray.rs:14-22 */
AGRS_API int agrs_relu_bwd_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, size_t n);  /* operators.rs:103 -> functional_ndarray.rs:23-34 */
AGRS_API int agrs_sigmoid_fwd_f32(agrs_ctx* ctx, const float* x, float* y, size_t n);                /* operators.rs:189-201 */
AGRS_API int agrs_sigmoid_bwd_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, size_t n); /* operators.rs:214-224: ((1-s)*s)*g, s recomputed from x */
/* out[0] = mean((p-t)^2)  (operators.rs:237) */
AGRS_API int agrs_mse_fwd_f32(agrs_ctx* ctx, const float* pred, const float* target, size_t n, float* out1);
/* out = ((p - t) * 2 / batch) * upstream[0]; upstream is a DEVICE scalar: no host round trip (operators.rs:247-253) */
AGRS_API int agrs_mse_bwd_f32(agrs_ctx* ctx, const float* pred, const float* target, const float* upstream1,
                     float batch, size_t n, float* out);

/* Fusions across the Operator boundary: the gradient that ReLU / Sigmoid / MSE backward hands to the Linear layer
 * below is the tensor whose sum over rows that layer needs for db (operators.rs:161).  These compute the elementwise result
 * [rows, cols] AND colsum[c] = sum_r result[r, c] in ONE pass -- the same values, bit for bit, as the unfused pair
 * (agrs_*_bwd_f32 then agrs_sum_axis_f32(outer 1, len rows, inner cols)): same functors, same summation tree.
 * Need cols % 4 == 0 and 16-byte aligned tensors (AGRS_ERR_UNSUPPORTED otherwise); honour agrs_next_output_absmax. */
AGRS_API int agrs_relu_bwd_colsum_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, int64_t rows, int64_t cols, float* colsum);
AGRS_API int agrs_sigmoid_bwd_colsum_f32(agrs_ctx* ctx, const float* x, const float* g, float* dx, int64_t rows, int64_t cols, float* colsum);
AGRS_API int agrs_mse_bwd_colsum_f32(agrs_ctx* ctx, const float* pred, const float* target, const float* upstream1, float batch,
                                     int64_t rows, int64_t cols, float* out, float* colsum);

/* ---- tensor arithmetic (src/tensor/storage_arithmetic.rs, src/tensor/utils.rs:16-75) ----
 * out[i] = a[i] (op) b[j],  j = i            when nb == n
 *                           j = i % nb       when bmode == AGRS_BCAST_TRAILING  ([O] or [1,O] onto [B,O]; [1,1] onto anything)
 *                           j = i / (n / nb) when bmode == AGRS_BCAST_LEADING   ([B,1] onto [B,O])
 * out may alias a (the reference's in-place forms A + &B, A -= &B, A * &B). */
enum agrs_binop { AGRS_ADD = 0, AGRS_SUB = 1, AGRS_MUL = 2 };
enum agrs_bcast { AGRS_BCAST_TRAILING = 0, AGRS_BCAST_LEADING = 1 };
AGRS_API int agrs_binary_f32(agrs_ctx* ctx, int op, float* out, const float* a, const float* b, size_t n, size_t nb, int bmode);
/* out[i] = a[i]*s | a[i]/s | s-a[i] | a[i]+s  (storage_arithmetic.rs:257-351) ; out may alias a */
enum agrs_scalarop { AGRS_SMUL = 0, AGRS_SDIV = 1, AGRS_SRSUB = 2, AGRS_SADD = 3 };
AGRS_API int agrs_scalar_f32(agrs_ctx* ctx, int op, float* out, const float* a, float s, size_t n);
AGRS_API int agrs_neg_f32(agrs_ctx* ctx, float* out, const float* a, size_t n);            /* storage_arithmetic.rs:202-208 (todo!() even on CPU) */
AGRS_API int agrs_powi_f32(agrs_ctx* ctx, float* out, const float* a, int exp, size_t n);  /* tensor.rs:220-236 */
AGRS_API int agrs_equal_f32(agrs_ctx* ctx, const float* a, const float* b, size_t n, int* host_equal); /* storage_arithmetic.rs:211-253; blocks */
AGRS_API int agrs_transpose_f32(agrs_ctx* ctx, const float* in, int64_t rows, int64_t cols, float* out); /* t_clone, storage.rs:84-88 (API completeness only) */

/* ---- reductions (src/tensor/storage.rs:91-112) ---- */
AGRS_API int agrs_sum_f32(agrs_ctx* ctx, const float* x, size_t n, float* out1);  /* storage.rs:91 */
AGRS_API int agrs_mean_f32(agrs_ctx* ctx, const float* x, size_t n, float* out1); /* storage.rs:101 */
/* x viewed as [outer, len, inner]; out[o, i] = sum_l x[o, l, i].
 * db = g.sum_axis(0) (operators.rs:161, :405): outer=1, len=B, inner=O.
 * col.sum_axis(-1)   (operators.rs:443):       outer=B*P*KK, len=C_in, inner=1. */
AGRS_API int agrs_sum_axis_f32(agrs_ctx* ctx, const float* x, int64_t outer, int64_t len, int64_t inner, float* out);

/* ---- optimiser (src/nn/optim.rs:25-34) ---- w[i] -= lr * g[i % ng]  (ng == n, or g broadcast: bias [1,O] -= [O]) */
AGRS_API int agrs_sgd_step_f32(agrs_ctx* ctx, float* w, const float* g, float lr, size_t n, size_t ng);
/* the same update for `count` parameters (each with a gradient of its own length) in one launch per 16 of them: SGD::step's loop
 * over the parameter list (optim.rs:25-34) without one launch per tensor.  out_absmax (may be NULL, entries may be NULL): where to
 * leave the magnitude word of each updated parameter (agrs_next_output_absmax's by-product, per tensor). */
AGRS_API int agrs_sgd_step_multi_f32(agrs_ctx* ctx, int count, float* const* w, const float* const* g, const size_t* n,
                                     uint32_t* const* out_absmax, float lr);

/* ---- Conv2D lowering (src/operators/functional_ndarray.rs:40-138, operators.rs:330,341,364-376) ---- */
/* col[b, y*Wo+x, c*k*k+i*k+j] = im[b,c,y*s-p+i,x*s-p+j] or 0;  col is [B, Ho*Wo, C*k*k] */
AGRS_API int agrs_im2col_f32(agrs_ctx* ctx, const float* im, int64_t B, int64_t C, int64_t H, int64_t W,
                    int64_t k, int64_t stride, int64_t pad, float* col);
/* overlap-add inverse; col is [B, Ho*Wo, C*k*k]; im_out is [B, C, (Ho-1)s+k, (Wo-1)s+k] (padded size, as the reference) */
AGRS_API int agrs_col2im_f32(agrs_ctx* ctx, const float* col, int64_t B, int64_t C, int64_t Ho, int64_t Wo,
                    int64_t k, int64_t stride, float* im_out);
/* Implicit-GEMM Conv2D for single-channel images: same results as im2col + GEMM (+ col2im, + the two reductions)
 * of operators.rs:350-449 without ever materialising col -- every kernel reads col[b, p, i*k+j] = im[b, 0, y*s-pad+i, x*s-pad+j]
 * from the sample staged in shared memory.  im [B,1,H,W]; filt [k,k,Co]; out / grad [B,Ho,Wo,Co] (NHWC);
 * dx [B,1,(Ho-1)s+k,(Wo-1)s+k] (padded size, as the reference); dw [k,k,Co]; db [Co].  agrs_conv2d_implicit_ok tells whether a
 * shape is taken (C == 1, Co <= 8, k*k*Co + Co <= 256, one sample's image + gradient within 160 KB); otherwise the calls return
 * AGRS_ERR_UNSUPPORTED and the caller lowers with agrs_im2col_f32 + agrs_gemm_f32. */
AGRS_API int agrs_conv2d_implicit_ok(int64_t B, int64_t C, int64_t H, int64_t W, int64_t k, int64_t c_out, int64_t stride, int64_t pad);
AGRS_API int agrs_conv2d_fwd_f32(agrs_ctx* ctx, const float* im, const float* filt, const float* bias, int64_t B, int64_t H, int64_t W,
                                 int64_t k, int64_t c_out, int64_t stride, int64_t pad, float* out);
/* Conv2D::forward followed by ReLU / Sigmoid (layers.rs:24-43 feeding operators.rs:350-390 into :112-116 / :195-200) as one launch:
 * out as agrs_conv2d_fwd_f32 writes it (the pre-activation, kept for the activation node's backward) and y = act(out), same layout,
 * written by the same streaming stage of the kernel; `act` is an agrs_act.  Same bits as the two separate calls; honours
 * agrs_next_output_absmax (reports on y). */
AGRS_API int agrs_conv2d_fwd_act_f32(agrs_ctx* ctx, const float* im, const float* filt, const float* bias, int64_t B, int64_t H, int64_t W,
                                     int64_t k, int64_t c_out, int64_t stride, int64_t pad, float* out, int act, float* y);
AGRS_API int agrs_conv2d_bwd_f32(agrs_ctx* ctx, const float* im, const float* filt, const float* grad, int64_t B, int64_t H, int64_t W,
                                 int64_t k, int64_t c_out, int64_t stride, int64_t pad, float* dx, float* dw, float* db);
/* out[c*kk + r, o] = filt[r, o] for every c: the [k,k,Co] -> [Ci,k,k,Co] broadcast (storage.rs:114-120) */
AGRS_API int agrs_broadcast_filter_f32(agrs_ctx* ctx, const float* filt, int64_t kk, int64_t c_out, int64_t c_in, float* out);

/* ---- data-parallel gradient exchange (new: the reference has no multi-device concept) ----
 * One process per GPU; NCCL (over NVLink 5 / NVSwitch) is loaded at run time (libnccl.so.2), never linked.
 * Collectives run on a communication stream owned by the comm: each call is ordered after everything
 * enqueued so far on the context's stream, and agrs_comm_join makes the context's stream wait for all
 * collectives enqueued so far -- so an all-reduce of layer l's gradient overlaps the backward of layer l-1. */
typedef struct agrs_comm agrs_comm;
#define AGRS_COMM_ID_BYTES 128
AGRS_API int agrs_comm_unique_id(void* out_id128);  /* rank 0 makes it, the launcher broadcasts it (ncclGetUniqueId) */
AGRS_API int agrs_comm_create(agrs_ctx* ctx, int world, int rank, const void* id128, agrs_comm** out);
AGRS_API int agrs_comm_destroy(agrs_comm* comm);
AGRS_API int agrs_comm_info(agrs_comm* comm, int* world, int* rank);
/* in place; average=0: sum over ranks, average=1: mean over ranks (ncclAvg: the sum divided by the world size) */
AGRS_API int agrs_comm_allreduce_f32(agrs_comm* comm, float* buf, size_t n, int average);
AGRS_API int agrs_comm_join(agrs_comm* comm);

/* ---- fused data-parallel optimiser step over NVLink peer memory (csrc/fused_dp.cu) ----
 * Each rank keeps parameters and parameter gradients in an ARENA: one agrs_peer_alloc block laid out identically on every
 * rank ([flag words | ... weight region at w_off ... | ... gradient region at g_off ...], offsets in floats).  The arenas are
 * exported (agrs_peer_export -> 64-byte handle, exchanged by the launcher) and mapped by every peer (agrs_peer_open).
 * The flat parameter space is dealt out in granules of 2^granule_log2 floats (granule g belongs to rank g % world); the gradient
 * region of an arena holds one slot per source rank for the slice this rank owns.  During backward every rank pushes its
 * gradients into its slot at the owners (agrs_gemm_f32_scatter from the dW GEMM's epilogue, agrs_fused_scatter_f32 otherwise).
 * The step is then finished, replacing all-reduce + agrs_sgd_step_f32 (optim.rs:25-34), in one of two ways:
 *  - per bucket, overlapped with the rest of backward: agrs_fused_begin_step(lr) before backward; agrs_fused_bucket_ready(b, off, n)
 *    as soon as every gradient of the contiguous range [off, off+n) (one layer's W and b) has been pushed -- a flag round, then
 *    on a communication stream the local sum of the slots + w -= lr * mean(g) on the owned granules, then the all-gather of the
 *    new weights by the copy engines (one strided peer copy per peer); agrs_fused_finish_step() after backward makes the compute
 *    stream wait for everything that is still in flight;
 *  - agrs_fused_sgd_step(lr): ONE kernel after backward (cross-rank barrier, slot sum, SGD, all-gather with P2P stores, barrier).
 * Every rank must make the same calls in the same order. */
#define AGRS_PEER_HANDLE_BYTES 64
typedef struct agrs_fused agrs_fused;
AGRS_API int agrs_peer_alloc(agrs_ctx* ctx, size_t nbytes, void** out);          /* zeroed; plain cudaMalloc (IPC-exportable) */
AGRS_API int agrs_peer_free(agrs_ctx* ctx, void* ptr);
AGRS_API int agrs_peer_export(agrs_ctx* ctx, void* ptr, void* handle64);
AGRS_API int agrs_peer_open(agrs_ctx* ctx, const void* handle64, void** out);    /* maps a PEER's arena into this process */
AGRS_API int agrs_peer_close(agrs_ctx* ctx, void* ptr);
AGRS_API size_t agrs_fused_flag_floats(void);                                    /* floats reserved at the start of an arena */
AGRS_API int agrs_fused_max_buckets(void);
/* arenas[p]: rank p's arena as mapped here (arenas[rank] = the local block); total: floats per region, a multiple of
 * 2^granule_log2 * world; granule_log2 in 5 .. 24 (14 = 64 KB granules: whole rows for the copy engines, a few rows of a weight
 * gradient per granule so that a GEMM tile's pushes still spread over all ranks); w_alt: see agrs_fused_weights_offset (0 = none) */
AGRS_API int agrs_fused_create(agrs_ctx* ctx, int world, int rank, void* const* arenas, size_t w_off, size_t w_alt, size_t g_off, size_t total,
                               int granule_log2, agrs_fused** out);
/* w_alt != 0: a second weight region of the same size (initialised with the same values).  The bucket protocol then reads the
 * current region and writes the new weights into the other one; agrs_fused_finish_step swaps them and the caller re-points its
 * parameter tensors (agrs_fused_weights_offset).  Nothing the exchange writes is read by the running step, so a layer's exchange
 * may start before that layer's dX GEMM has read W. */
AGRS_API int agrs_fused_weights_offset(agrs_fused* f, size_t* w_off_now);
/* while agrs_profile_enable is on, agrs_fused_finish_step brackets its waits with CUDA events: the time the compute stream spent
 * waiting for the exchange, i.e. the part backward did not hide.  Synchronises, returns the sum since the last call and resets. */
AGRS_API int agrs_fused_exposed_ms(agrs_fused* f, double* total_ms, uint64_t* steps);
/* how this exchange synchronises and moves data: stream memory operations in use, copy streams, all-gather by kernel stores */
AGRS_API int agrs_fused_info(agrs_fused* f, int* memops, int* copy_streams, int* allgather_kernel);
AGRS_API int agrs_fused_begin_step(agrs_fused* f, float lr);
/* bucket < agrs_fused_max_buckets(); off and n multiples of 2^granule_log2 * world */
AGRS_API int agrs_fused_bucket_ready(agrs_fused* f, int bucket, size_t off, size_t n);
AGRS_API int agrs_fused_finish_step(agrs_fused* f);
AGRS_API int agrs_fused_destroy(agrs_fused* f);
/* the scatter half of the reduce-scatter: push `n` gradient values that belong at float offset `off` of the (flat) parameter
 * space into this rank's slot at their owners.  src is device memory, 16-byte aligned; off a multiple of 4. */
AGRS_API int agrs_fused_scatter_f32(agrs_fused* f, const float* src, size_t off, size_t n);
/* agrs_gemm_f32 whose result C ([M,N] contiguous, ldc == N) is ALSO pushed to the owners' slots as the tiles are produced: the
 * GEMM -> reduce-scatter of the weight gradient in one kernel (P2P stores from the epilogue of the tcgen05 pair kernel).  Shapes
 * that kernel does not take run agrs_gemm_f32 followed by agrs_fused_scatter_f32: same result. */
AGRS_API int agrs_gemm_f32_scatter(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                                   const float* B, int64_t ldb, float* C, const float* bias, agrs_fused* f, size_t off);
AGRS_API int agrs_fused_sgd_step(agrs_fused* f, float lr);

AGRS_API const char* agrs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGRS_H_ */
