/*
 * agrs_engine.h -- C ABI of libagrs_engine.so: the host engine above the kernels of agrs.h.
 *
 * The reference's host side is Rust (Tensor / Operator / Node / Layer / NN / SGD); no Rust toolchain
 * exists in the build image, so the host engine is C++ (autograd.rs_b200/engine/) mirroring those
 * types one to one, and this header is the handle-based surface through which a test harness, a
 * Python front end or a Rust crate drives it.  Each entry point cites the reference item it mirrors
 * (paths relative to the reference root).  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 or an agrs_status (agrs.h); the message of the failure -- the
 *     reference's panic text where it has one -- is returned by agrs_eng_last_error() (thread-local);
 *   - an agrs_tensor* is one `Tensor<f32>` HANDLE (tensor.rs:27-34): clones share the data cell,
 *     the grad cell and the graph pointer.  Every handle returned here is owned by the caller and
 *     released with agrs_eng_tensor_free (dropping a handle frees device memory only when it was
 *     the last owner -- CudaData's Drop, storage.rs:10-14);
 *   - single host thread per device (Rc<RefCell>: utils/mod.rs:5); nothing here blocks except
 *     *_to_host, *_item, *_sum, *_equal and device_sync.
 */
#ifndef AGRS_ENGINE_H_
#define AGRS_ENGINE_H_

#include "agrs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agrs_device agrs_device;
typedef struct agrs_tensor agrs_tensor;
typedef struct agrs_dp agrs_dp;

AGRS_API const char* agrs_eng_last_error(void);

/* ---- device ---- */
AGRS_API int agrs_eng_device_create(int ordinal, agrs_device** out);
AGRS_API int agrs_eng_device_create_on_stream(int ordinal, void* cuda_stream, agrs_device** out);
AGRS_API int agrs_eng_device_destroy(agrs_device* dev);
AGRS_API agrs_ctx* agrs_eng_device_ctx(agrs_device* dev);        /* the kernel-level context underneath */
AGRS_API int agrs_eng_device_sync(agrs_device* dev);
AGRS_API int agrs_eng_device_seed(agrs_device* dev, uint64_t seed); /* the reference's init is unseeded (init.rs:27,48) */
/* MatMul::backward flavour: 1 = literal reference (no transposes, defect D3, operators.rs:183-184; default), 0 = mathematically correct */
AGRS_API int agrs_eng_device_set_matmul_backward_literal(agrs_device* dev, int literal);
/* Fusions across the Operator boundary, on by default, bit-identical to the unfused path:
 * backward_colsum: ReLU / Sigmoid / MSE backward of a [B, O] tensor also leaves the sum over rows of its result with the buffer,
 * which Linear::backward of the layer below takes as db (operators.rs:161) instead of re-reading the gradient. */
AGRS_API int agrs_eng_device_set_fusions(agrs_device* dev, int backward_colsum);
/* forward_activation (on by default, bit-identical either way): a Linear operator feeding ReLU / Sigmoid -- inside one Layer's
 * subgraph, or through agrs_eng_op_forward_pair -- runs as one agrs_gemm_bias_act_f32 call: the activation is written by the GEMM's
 * epilogue next to the pre-activation, which stays the activation node's variable (operators.rs:120-129, :214-224). */
AGRS_API int agrs_eng_device_set_forward_fusion(agrs_device* dev, int on);
/* Conv2D with one input channel (on by default): implicit-GEMM kernels that never materialise col; the node then keeps its three
 * inputs only.  0 = always the literal lowering of operators.rs:350-449 (im2col + GEMM, col cached as the 4th variable). */
AGRS_API int agrs_eng_device_set_conv_implicit(agrs_device* dev, int on);

/* ---- Tensor constructors (tensor.rs:37-92, init.rs:19-49) ---- */
AGRS_API int agrs_eng_tensor_from_host(agrs_device* dev, const float* host, const int64_t* shape, int ndim, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_full(agrs_device* dev, const int64_t* shape, int ndim, float value, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_uniform(agrs_device* dev, const int64_t* shape, int ndim, float low, float high, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_kaiming_uniform(agrs_device* dev, const int64_t* shape, int ndim, agrs_tensor** out);
/* overwrite t's data from a host buffer of t.len() floats.  blocking=0: asynchronous on the device's stream -- the
 * caller keeps `host` alive and unmodified until the next sync (pinned memory, agrs_host_alloc, for real overlap). */
AGRS_API int agrs_eng_tensor_copy_from_host(agrs_tensor* t, const float* host, int64_t n, int blocking);
/* the same copy on the device's COPY stream (agrs_upload_prefetch_event): overlaps the training step already enqueued.  The
 * tensor keeps the copy's completion event and the first engine call that reads or overwrites it makes the compute stream wait
 * for that copy alone (agrs_eng_device_prefetch_join still waits for all of them).  `host` must be pinned. */
AGRS_API int agrs_eng_tensor_prefetch_from_host(agrs_tensor* t, const float* host, int64_t n);
AGRS_API int agrs_eng_device_prefetch_join(agrs_device* dev);
AGRS_API int agrs_eng_tensor_clone(const agrs_tensor* t, agrs_tensor** out);    /* #[derive(Clone)]: shares data, grad, graph */
AGRS_API int agrs_eng_tensor_to_owned(const agrs_tensor* t, agrs_tensor** out); /* tensor.rs:93-105: graph=None, still shares data */
AGRS_API int agrs_eng_tensor_free(agrs_tensor* t);

/* ---- metadata / read-back (tensor.rs:107-133, :364-421; storage.rs:139-177) ---- */
AGRS_API int agrs_eng_tensor_ndim(const agrs_tensor* t, int* ndim);
AGRS_API int agrs_eng_tensor_shape(const agrs_tensor* t, int64_t* shape_out, int cap);
AGRS_API int agrs_eng_tensor_len(const agrs_tensor* t, int64_t* len);
AGRS_API int agrs_eng_tensor_is_contiguous(const agrs_tensor* t, int* yes);
AGRS_API int agrs_eng_tensor_to_host(const agrs_tensor* t, float* host_out, int64_t n); /* logical row-major order; blocks */
/* the same read-back without stalling the pipeline (agrs_download_async): enqueued behind the work that produces t, returns at
 * once; agrs_eng_event_wait blocks until that copy has landed.  `host_out` must be pinned and stay alive; contiguous tensors only. */
AGRS_API int agrs_eng_event_create(agrs_device* dev, agrs_event** out);
AGRS_API int agrs_eng_event_destroy(agrs_device* dev, agrs_event* ev);
AGRS_API int agrs_eng_event_wait(agrs_device* dev, agrs_event* ev);
AGRS_API int agrs_eng_tensor_to_host_async(const agrs_tensor* t, float* host_out, int64_t n, agrs_event* done);
AGRS_API int agrs_eng_tensor_data_ptr(const agrs_tensor* t, void** device_ptr);         /* address identity tests (tensor.rs:474-492) */
AGRS_API int agrs_eng_tensor_requires_grad(const agrs_tensor* t, int* yes);
AGRS_API int agrs_eng_tensor_set_requires_grad(agrs_tensor* t, int yes);               /* field on the HANDLE, not shared (tensor.rs:28) */
AGRS_API int agrs_eng_tensor_has_grad(const agrs_tensor* t, int* yes);
AGRS_API int agrs_eng_tensor_grad(const agrs_tensor* t, agrs_tensor** out);            /* a handle on the grad storage; *out = NULL when grad is None */
/* graph introspection (autograd.rs:12-47): op name of the node that produced t ("" if none), #variables, #parents that are Some */
AGRS_API int agrs_eng_tensor_graph_info(const agrs_tensor* t, char* op_name, int cap, int* n_variables, int* n_parents);

/* ---- shape ops: mutate the SHARED storage metadata (tensor.rs:165-186, :370-415) ---- */
AGRS_API int agrs_eng_tensor_reshape(agrs_tensor* t, const int64_t* shape, int ndim);
AGRS_API int agrs_eng_tensor_t(agrs_tensor* t);                              /* in-place transpose of the shared storage */
AGRS_API int agrs_eng_tensor_t_clone(const agrs_tensor* t, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_unsqueeze(const agrs_tensor* t, int axis, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_squeeze(const agrs_tensor* t, agrs_tensor** out);

/* ---- arithmetic with the reference's allocation rules (tensor_arithmetic.rs:41-338) ----
 * op: agrs_binop / agrs_scalarop of agrs.h.  inplace=0: `&a op &b` -> new tensor; inplace=1: `a op &b` -> writes a's storage,
 * *out is a handle on a.  scalar inplace=1 is `a *= s` / `a /= s`. */
AGRS_API int agrs_eng_tensor_binary(int op, int inplace, agrs_tensor* a, const agrs_tensor* b, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_scalar(int op, int inplace, agrs_tensor* a, float s, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_neg(const agrs_tensor* a, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_equal(const agrs_tensor* a, const agrs_tensor* b, int* equal);  /* blocks */
AGRS_API int agrs_eng_tensor_fill(agrs_tensor* t, float v);
AGRS_API int agrs_eng_tensor_powi(agrs_tensor* t, int exp, int inplace, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_sum(const agrs_tensor* t, float* host_out);                      /* blocks */
AGRS_API int agrs_eng_tensor_sum_axis(const agrs_tensor* t, int axis, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_mean(const agrs_tensor* t, agrs_tensor** out);
AGRS_API int agrs_eng_tensor_dot(const agrs_tensor* a, const agrs_tensor* b, agrs_tensor** out); /* tensor.rs:301-320 */

/* ---- autograd (tensor.rs:135-163, :328-357; autograd.rs:49-86) ---- */
AGRS_API int agrs_eng_tensor_zero_grad(agrs_tensor* t);
AGRS_API int agrs_eng_tensor_accumulate_grad(agrs_tensor* t, const agrs_tensor* grad_tensor); /* accumulate_grad_from_grad_tensor */
AGRS_API int agrs_eng_tensor_backward(agrs_tensor* t);

/* ---- operators (operators.rs:18-479) ----
 * `params` is read only for AGRS_OP_CONV2D: {in_channels, out_channels, k_size, stride, pad} (operators.rs:66-73). */
enum agrs_op_kind {
  AGRS_OP_RELU = 0, AGRS_OP_SIGMOID = 1, AGRS_OP_LINEAR = 2, AGRS_OP_MATMUL = 3, AGRS_OP_MSE = 4,
  AGRS_OP_MEAN = 5, AGRS_OP_IDENTITY = 6, AGRS_OP_CONV2D = 7,
  AGRS_OP_FLATTEN = 8 /* not in the reference: [B,...] -> [B, prod] view so Conv2D can feed Linear (SURVEY gap G1) */
};
/* Operator::forward: computes, then attach_to_eager_graph (operators.rs:27-49) */
AGRS_API int agrs_eng_op_forward(int op_kind, const int64_t* params, agrs_tensor* const* xs, int n_xs, agrs_tensor** out);
/* Operator::backward: one gradient per input (MSE: one in total).  xs for Conv2D includes the cached col as xs[3]. */
/* second.forward(vec![first.forward(xs)]) -- the loop body of Layer::forward_default (layers.rs:24-43) for two consecutive
 * operators -- with the same two graph nodes and the same bits; Linear + ReLU / Sigmoid go out as one GEMM call (see
 * agrs_eng_device_set_forward_fusion), every other pair as two calls. */
AGRS_API int agrs_eng_op_forward_pair(int kind_first, const int64_t* params_first, int kind_second, const int64_t* params_second,
                                      agrs_tensor* const* xs, int n_xs, agrs_tensor** out);
AGRS_API int agrs_eng_op_backward(int op_kind, const int64_t* params, agrs_tensor* const* xs, int n_xs, const agrs_tensor* grad,
                                  agrs_tensor** grads_out, int cap, int* n_grads);
AGRS_API int agrs_eng_im2col(const agrs_tensor* im, int64_t k, int64_t stride, int64_t pad, agrs_tensor** out);          /* operators.rs:324-332 */
AGRS_API int agrs_eng_im2col_backward(const agrs_tensor* col, int64_t h_out, int64_t w_out, int64_t stride, int64_t k,
                                      agrs_tensor** out);                                                                /* operators.rs:333-342 */

/* ---- nn (optim.rs:25-34, model.rs:14-21) ---- */
AGRS_API int agrs_eng_sgd_step(agrs_tensor* const* params, int n, float lr);
AGRS_API int agrs_eng_zero_grad(agrs_tensor* const* params, int n); /* only parameters whose grad is Some */
/* Model serialisation (a TODO of the reference, README.md:71): "AGRSWTS1" | u32 count | per tensor { u32 ndim | u64 dims | f32 data },
 * little-endian, logical row-major order.  load writes into the existing parameter storages (clones follow; arena-resident
 * parameters stay in the arena) and fails with AGRS_ERR_SHAPE on a count / shape mismatch. */
AGRS_API int agrs_eng_save_parameters(agrs_tensor* const* params, int n, const char* path);
AGRS_API int agrs_eng_load_parameters(agrs_tensor* const* params, int n, const char* path);

/* ---- data parallel training (new: the reference has no multi-device concept; BASELINE north star) ----
 * One process per GPU.  Every rank holds identical parameters and a shard of the minibatch rows; after backward the
 * parameter gradients are summed over ranks (NCCL over NVLink) and divided by the world size, which is exactly the
 * literal MSE gradient of the global batch (operators.rs:247 divides by the LOCAL batch).
 * `unique_id` is the 128-byte ncclUniqueId made on rank 0 (agrs_comm_unique_id, agrs.h) and shared by the launcher. */
AGRS_API int agrs_eng_dp_create(agrs_device* dev, int world, int rank, const void* unique_id, agrs_dp** out);
AGRS_API int agrs_eng_dp_destroy(agrs_dp* dp);
/* register the parameters whose gradients are exchanged.  overlap=1: each all-reduce is enqueued on the communication
 * stream the moment backward accumulates that gradient, overlapping the rest of the backward pass. */
AGRS_API int agrs_eng_dp_register(agrs_dp* dp, agrs_tensor* const* params, int n, int overlap);
/* all-reduce whatever has not been reduced yet, make the compute stream wait for the communication stream, and turn
 * the sums into means.  Call between backward() and sgd_step(). */
AGRS_API int agrs_eng_dp_sync_grads(agrs_dp* dp);
/* Fused optimiser step over NVLink peer memory (agrs_fused_*, agrs.h): reduce-scatter of the gradients (pushed by the dW GEMM's
 * epilogue), w -= lr * mean(g) on the owned slice and all-gather of the new weights; it replaces agrs_eng_dp_sync_grads +
 * agrs_eng_sgd_step for the registered parameters.  Set-up, once, after agrs_eng_dp_register:
 *   agrs_eng_dp_fused_prepare(dp, handle)   moves the parameters into an exportable arena (their storages are re-pointed: every
 *                                           clone follows) and returns its 64-byte IPC handle;
 *   the launcher gathers the handles of all ranks (world x 64 bytes, rank order);
 *   agrs_eng_dp_fused_connect(dp, handles)  maps the peers' arenas.
 * Then, every iteration, either  agrs_eng_dp_fused_begin(dp, lr); backward(); agrs_eng_dp_fused_finish(dp); zero_grad()  -- each
 * layer's exchange and update start during backward as soon as its gradients exist and overlap the backward of the earlier
 * layers (every parameter must receive exactly one gradient per step) -- or  backward(); agrs_eng_dp_fused_sgd_step(dp, lr);
 * zero_grad()  (one kernel after backward).  The tensors' grad() keep the LOCAL gradient of this rank (the mean only exists
 * inside the kernels).  agrs_eng_dp_fused_disable returns to the NCCL exchange when a peer could not connect. */
AGRS_API int agrs_eng_dp_fused_prepare(agrs_dp* dp, void* handle64_out);
AGRS_API int agrs_eng_dp_fused_connect(agrs_dp* dp, const void* handles);
AGRS_API int agrs_eng_dp_fused_disable(agrs_dp* dp);
AGRS_API int agrs_eng_dp_fused_begin(agrs_dp* dp, float lr);
AGRS_API int agrs_eng_dp_fused_finish(agrs_dp* dp);
AGRS_API int agrs_eng_dp_fused_sgd_step(agrs_dp* dp, float lr);
/* profiling (agrs_profile_enable on): how long the compute stream waited inside agrs_eng_dp_fused_finish since the last call */
AGRS_API int agrs_eng_dp_fused_exposed_ms(agrs_dp* dp, double* total_ms, uint64_t* steps);
/* mean over ranks of a 1-element tensor (the loss value), in place */
AGRS_API int agrs_eng_dp_mean_scalar(agrs_dp* dp, agrs_tensor* t);

#ifdef __cplusplus
}
#endif
#endif /* AGRS_ENGINE_H_ */
