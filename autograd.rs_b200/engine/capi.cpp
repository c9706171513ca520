// extern "C" surface of the host engine (include/agrs_engine.h): handles in, handles out, Panic -> status + message.
#include <cstring>

#include "agrs_engine.h"
#include "autograd.hpp"
#include "dp.hpp"
#include "nn.hpp"

using namespace agrs;

struct agrs_device {
  std::shared_ptr<Device> holder;  // buffers share ownership: tensor handles may be released after the device handle
  Device& dev;
  agrs_device(int ordinal, void* stream, bool on_stream) : holder(std::make_shared<Device>(ordinal, stream, on_stream)), dev(*holder) {}
};
struct agrs_tensor {
  Tensor t;
};
struct agrs_dp {
  DataParallel dp;
  agrs_dp(Device* d, int world, int rank, const void* id) : dp(d, world, rank, id) {}
};

namespace {

thread_local std::string g_err;

template <class F>
int guarded(F&& f) {
  try {
    f();
    return AGRS_OK;
  } catch (const Panic& p) {
    g_err = p.what();
    return p.code ? p.code : AGRS_ERR_INVALID;
  } catch (const std::exception& e) {
    g_err = e.what();
    return AGRS_ERR_INVALID;
  }
}

agrs_tensor* wrap(Tensor t) { return new agrs_tensor{std::move(t)}; }
Shape mkshape(const int64_t* s, int nd) {
  if (nd < 0 || (nd > 0 && !s)) throw Panic("null shape");
  return Shape(s, s + nd);
}
void need(const void* p, const char* what) {
  if (!p) throw Panic(std::string("null ") + what);
}
std::vector<Tensor> gather(agrs_tensor* const* xs, int n) {
  if (n < 0 || (n > 0 && !xs)) throw Panic("null tensor list");
  std::vector<Tensor> v;
  for (int i = 0; i < n; ++i) {
    need(xs[i], "tensor in list");
    v.push_back(xs[i]->t);
  }
  return v;
}
Operators make_op(int kind, const int64_t* params) {
  switch (kind) {
    case AGRS_OP_RELU: return ReLU{};
    case AGRS_OP_SIGMOID: return Sigmoid{};
    case AGRS_OP_LINEAR: return Linear{};
    case AGRS_OP_MATMUL: return MatMul{};
    case AGRS_OP_MSE: return MeanSquaredError{};
    case AGRS_OP_MEAN: return Mean{};
    case AGRS_OP_IDENTITY: return Identity{};
    case AGRS_OP_FLATTEN: return Flatten{};
    case AGRS_OP_CONV2D: {
      need(params, "Conv2D parameters");
      Conv2D c;
      c.in_channels = params[0];
      c.out_channels = params[1];
      c.k_size = params[2];
      c.stride = params[3];
      c.pad = params[4];
      if (c.k_size < 1 || c.stride < 1 || c.pad < 0 || c.in_channels < 1 || c.out_channels < 1) throw Panic("Conv2D: bad geometry");
      return c;
    }
  }
  throw Panic("unknown operator kind " + std::to_string(kind));
}

}  // namespace

extern "C" {

const char* agrs_eng_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------- device
int agrs_eng_device_create(int ordinal, agrs_device** out) {
  return guarded([&] {
    need(out, "out");
    *out = new agrs_device(ordinal, nullptr, false);
  });
}
int agrs_eng_device_create_on_stream(int ordinal, void* stream, agrs_device** out) {
  return guarded([&] {
    need(out, "out");
    *out = new agrs_device(ordinal, stream, true);
  });
}
int agrs_eng_device_destroy(agrs_device* d) {
  return guarded([&] {
    need(d, "device");
    delete d;
  });
}
agrs_ctx* agrs_eng_device_ctx(agrs_device* d) { return d ? d->dev.ctx() : nullptr; }
int agrs_eng_device_sync(agrs_device* d) {
  return guarded([&] {
    need(d, "device");
    d->dev.sync();
  });
}
int agrs_eng_device_seed(agrs_device* d, uint64_t seed) {
  return guarded([&] {
    need(d, "device");
    d->dev.seed(seed);
  });
}
int agrs_eng_device_set_fusions(agrs_device* d, int backward_colsum) {
  return guarded([&] {
    need(d, "device");
    d->dev.fuse_backward_colsum = backward_colsum != 0;
  });
}
int agrs_eng_device_set_forward_fusion(agrs_device* d, int on) {
  return guarded([&] {
    need(d, "device");
    d->dev.fuse_forward_activation = on != 0;
  });
}
int agrs_eng_device_set_conv_implicit(agrs_device* d, int on) {
  return guarded([&] {
    need(d, "device");
    d->dev.conv_implicit = on != 0;
  });
}
int agrs_eng_device_set_matmul_backward_literal(agrs_device* d, int literal) {
  return guarded([&] {
    need(d, "device");
    d->dev.matmul_backward_literal = literal != 0;
  });
}

// ---------------------------------------------------------------- constructors
int agrs_eng_tensor_from_host(agrs_device* d, const float* host, const int64_t* shape, int ndim, agrs_tensor** out) {
  return guarded([&] {
    need(d, "device");
    need(out, "out");
    Shape s = mkshape(shape, ndim);
    if (numel(s) > 0) need(host, "host data");
    *out = wrap(Tensor::from(&d->dev, host, s));
  });
}
int agrs_eng_tensor_full(agrs_device* d, const int64_t* shape, int ndim, float value, agrs_tensor** out) {
  return guarded([&] {
    need(d, "device");
    need(out, "out");
    *out = wrap(Tensor::full(&d->dev, mkshape(shape, ndim), value));
  });
}
int agrs_eng_tensor_uniform(agrs_device* d, const int64_t* shape, int ndim, float low, float high, agrs_tensor** out) {
  return guarded([&] {
    need(d, "device");
    need(out, "out");
    *out = wrap(Tensor::uniform(&d->dev, mkshape(shape, ndim), low, high));
  });
}
int agrs_eng_tensor_kaiming_uniform(agrs_device* d, const int64_t* shape, int ndim, agrs_tensor** out) {
  return guarded([&] {
    need(d, "device");
    need(out, "out");
    *out = wrap(Tensor::kaiming_uniform(&d->dev, mkshape(shape, ndim)));
  });
}
int agrs_eng_tensor_copy_from_host(agrs_tensor* t, const float* host, int64_t n, int blocking) {
  return guarded([&] {
    need(t, "tensor");
    if (n != t->t.len()) throw Panic("copy_from_host: buffer holds " + std::to_string(n) + " elements, tensor has " + std::to_string(t->t.len()));
    if (n == 0) return;
    need(host, "host data");
    Storage& s = *t->t.data;
    if (s.transposed) throw Panic("copy_from_host into a transposed view", AGRS_ERR_UNSUPPORTED);
    s.dev()->check(agrs_upload(s.dev()->ctx(), s.wptr(), host, size_t(n) * sizeof(float)));
    if (blocking) s.dev()->sync();
  });
}
int agrs_eng_tensor_prefetch_from_host(agrs_tensor* t, const float* host, int64_t n) {
  return guarded([&] {
    need(t, "tensor");
    if (n != t->t.len()) throw Panic("prefetch_from_host: buffer holds " + std::to_string(n) + " elements, tensor has " + std::to_string(t->t.len()));
    if (n == 0) return;
    need(host, "host data");
    Storage& s = *t->t.data;
    if (s.transposed) throw Panic("prefetch_from_host into a transposed view", AGRS_ERR_UNSUPPORTED);
    float* dst = s.wptr();  // (orders this copy behind an earlier one into the same buffer that nobody has read yet)
    Buffer& b = *s.buf;
    if (!b.landing) s.dev()->check(agrs_event_create(s.dev()->ctx(), &b.landing));
    s.dev()->check(agrs_upload_prefetch_event(s.dev()->ctx(), dst, host, size_t(n) * sizeof(float), b.landing));
    b.landing_pending = true;  // the first kernel that touches the buffer waits for this copy and no other
  });
}
int agrs_eng_device_prefetch_join(agrs_device* d) {
  return guarded([&] {
    need(d, "device");
    d->dev.check(agrs_prefetch_join(d->dev.ctx()));
  });
}
int agrs_eng_tensor_clone(const agrs_tensor* t, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(t->t.clone());
  });
}
int agrs_eng_tensor_to_owned(const agrs_tensor* t, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(t->t.to_owned());
  });
}
int agrs_eng_tensor_free(agrs_tensor* t) {
  return guarded([&] { delete t; });
}

// ---------------------------------------------------------------- metadata
int agrs_eng_tensor_ndim(const agrs_tensor* t, int* ndim) {
  return guarded([&] {
    need(t, "tensor");
    need(ndim, "out");
    *ndim = int(t->t.ndim());
  });
}
int agrs_eng_tensor_shape(const agrs_tensor* t, int64_t* shape_out, int cap) {
  return guarded([&] {
    need(t, "tensor");
    const Shape& s = t->t.shape();
    if (int(s.size()) > cap) throw Panic("shape buffer too small");
    if (!s.empty()) need(shape_out, "out");
    for (size_t i = 0; i < s.size(); ++i) shape_out[i] = s[i];
  });
}
int agrs_eng_tensor_len(const agrs_tensor* t, int64_t* len) {
  return guarded([&] {
    need(t, "tensor");
    need(len, "out");
    *len = t->t.len();
  });
}
int agrs_eng_event_create(agrs_device* d, agrs_event** out) {
  return guarded([&] {
    need(d, "device");
    need(out, "out");
    d->dev.check(agrs_event_create(d->dev.ctx(), out));
  });
}
int agrs_eng_event_destroy(agrs_device* d, agrs_event* ev) {
  return guarded([&] {
    need(d, "device");
    d->dev.check(agrs_event_destroy(d->dev.ctx(), ev));
  });
}
int agrs_eng_event_wait(agrs_device* d, agrs_event* ev) {
  return guarded([&] {
    need(d, "device");
    need(ev, "event");
    d->dev.check(agrs_event_wait(d->dev.ctx(), ev));
  });
}
int agrs_eng_tensor_to_host_async(const agrs_tensor* t, float* host_out, int64_t n, agrs_event* done) {
  return guarded([&] {
    need(t, "tensor");
    need(done, "event");
    if (n != t->t.len()) throw Panic("to_host_async: buffer holds " + std::to_string(n) + " elements, tensor has " + std::to_string(t->t.len()));
    if (n != 0) need(host_out, "out");
    const Storage& s = *t->t.data;
    if (s.transposed) throw Panic("to_host_async of a transposed view", AGRS_ERR_UNSUPPORTED);
    s.dev()->check(agrs_download_async(s.dev()->ctx(), host_out, s.rptr(), size_t(n) * sizeof(float), done));
  });
}
int agrs_eng_tensor_is_contiguous(const agrs_tensor* t, int* yes) {
  return guarded([&] {
    need(t, "tensor");
    need(yes, "out");
    *yes = t->t.is_contiguous();
  });
}
int agrs_eng_tensor_to_host(const agrs_tensor* t, float* host_out, int64_t n) {
  return guarded([&] {
    need(t, "tensor");
    if (n != t->t.len()) throw Panic("to_host: buffer holds " + std::to_string(n) + " elements, tensor has " + std::to_string(t->t.len()));
    if (n == 0) return;
    need(host_out, "out");
    const Storage& s = *t->t.data;
    if (!s.transposed) {  // straight into the caller's (possibly pinned) buffer
      s.dev()->check(agrs_download(s.dev()->ctx(), host_out, s.rptr(), size_t(n) * sizeof(float)));
    } else {
      auto v = s.to_host();
      std::memcpy(host_out, v.data(), size_t(n) * sizeof(float));
    }
  });
}
int agrs_eng_tensor_data_ptr(const agrs_tensor* t, void** p) {
  return guarded([&] {
    need(t, "tensor");
    need(p, "out");
    t->t.data->buf->mx_valid = false;  // the caller may write through the raw pointer
    *p = t->t.data->buf->ptr;
  });
}
int agrs_eng_tensor_requires_grad(const agrs_tensor* t, int* yes) {
  return guarded([&] {
    need(t, "tensor");
    need(yes, "out");
    *yes = t->t.requires_grad;
  });
}
int agrs_eng_tensor_set_requires_grad(agrs_tensor* t, int yes) {
  return guarded([&] {
    need(t, "tensor");
    t->t.requires_grad = yes != 0;
  });
}
int agrs_eng_tensor_has_grad(const agrs_tensor* t, int* yes) {
  return guarded([&] {
    need(t, "tensor");
    need(yes, "out");
    *yes = t->t.has_grad();
  });
}
int agrs_eng_tensor_grad(const agrs_tensor* t, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = t->t.has_grad() ? wrap(Tensor(t->t.grad->get())) : nullptr;
  });
}
int agrs_eng_tensor_graph_info(const agrs_tensor* t, char* op_name, int cap, int* n_variables, int* n_parents) {
  return guarded([&] {
    need(t, "tensor");
    std::string name;
    int nv = 0, np = 0;
    if (t->t.graph) {
      name = operator_name(t->t.graph->op);
      nv = int(t->t.graph->variables.size());
      for (auto& p : t->t.graph->parents) np += p ? 1 : 0;
    }
    if (op_name && cap > 0) {
      std::strncpy(op_name, name.c_str(), size_t(cap) - 1);
      op_name[cap - 1] = 0;
    }
    if (n_variables) *n_variables = nv;
    if (n_parents) *n_parents = np;
  });
}

// ---------------------------------------------------------------- shape ops
int agrs_eng_tensor_reshape(agrs_tensor* t, const int64_t* shape, int ndim) {
  return guarded([&] {
    need(t, "tensor");
    t->t.reshape(mkshape(shape, ndim));
  });
}
int agrs_eng_tensor_t(agrs_tensor* t) {
  return guarded([&] {
    need(t, "tensor");
    t->t.t();
  });
}
int agrs_eng_tensor_t_clone(const agrs_tensor* t, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(t->t.t_clone());
  });
}
int agrs_eng_tensor_unsqueeze(const agrs_tensor* t, int axis, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(t->t.unsqueeze(axis));
  });
}
int agrs_eng_tensor_squeeze(const agrs_tensor* t, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(t->t.squeeze());
  });
}

// ---------------------------------------------------------------- arithmetic
int agrs_eng_tensor_binary(int op, int inplace, agrs_tensor* a, const agrs_tensor* b, agrs_tensor** out) {
  return guarded([&] {
    need(a, "tensor a");
    need(b, "tensor b");
    need(out, "out");
    Tensor r;
    switch (op) {
      case AGRS_ADD: r = inplace ? add_inplace(a->t, b->t) : add_new(a->t, b->t); break;
      case AGRS_SUB: r = inplace ? sub_inplace(a->t, b->t) : sub_new(a->t, b->t); break;
      case AGRS_MUL: r = inplace ? mul_inplace(a->t, b->t) : mul_new(a->t, b->t); break;
      default: throw Panic("unknown binary op " + std::to_string(op));
    }
    *out = wrap(r);
  });
}
int agrs_eng_tensor_scalar(int op, int inplace, agrs_tensor* a, float s, agrs_tensor** out) {
  return guarded([&] {
    need(a, "tensor");
    need(out, "out");
    if (inplace) {
      if (op == AGRS_SMUL) mul_assign_scalar(a->t, s);
      else if (op == AGRS_SDIV) div_assign_scalar(a->t, s);
      else throw Panic("only `*=` and `/=` exist in place (tensor_arithmetic.rs:285-323)");
      *out = wrap(a->t);
      return;
    }
    switch (op) {
      case AGRS_SMUL: *out = wrap(mul_scalar(a->t, s)); break;
      case AGRS_SDIV: *out = wrap(div_scalar(a->t, s)); break;
      case AGRS_SRSUB: *out = wrap(rsub_scalar(s, a->t)); break;
      default: throw Panic("unknown scalar op " + std::to_string(op));
    }
  });
}
int agrs_eng_tensor_neg(const agrs_tensor* a, agrs_tensor** out) {
  return guarded([&] {
    need(a, "tensor");
    need(out, "out");
    *out = wrap(neg(a->t));
  });
}
int agrs_eng_tensor_equal(const agrs_tensor* a, const agrs_tensor* b, int* eq) {
  return guarded([&] {
    need(a, "tensor a");
    need(b, "tensor b");
    need(eq, "out");
    *eq = equal(a->t, b->t);
  });
}
int agrs_eng_tensor_fill(agrs_tensor* t, float v) {
  return guarded([&] {
    need(t, "tensor");
    t->t.fill(v);
  });
}
int agrs_eng_tensor_powi(agrs_tensor* t, int exp, int inplace, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(inplace ? t->t.powi_inplace(exp) : t->t.powi(exp));
  });
}
int agrs_eng_tensor_sum(const agrs_tensor* t, float* host_out) {
  return guarded([&] {
    need(t, "tensor");
    need(host_out, "out");
    *host_out = t->t.sum();
  });
}
int agrs_eng_tensor_sum_axis(const agrs_tensor* t, int axis, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(t->t.sum_axis(axis));
  });
}
int agrs_eng_tensor_mean(const agrs_tensor* t, agrs_tensor** out) {
  return guarded([&] {
    need(t, "tensor");
    need(out, "out");
    *out = wrap(t->t.mean());
  });
}
int agrs_eng_tensor_dot(const agrs_tensor* a, const agrs_tensor* b, agrs_tensor** out) {
  return guarded([&] {
    need(a, "tensor a");
    need(b, "tensor b");
    need(out, "out");
    *out = wrap(a->t.dot(b->t));
  });
}

// ---------------------------------------------------------------- autograd
int agrs_eng_tensor_zero_grad(agrs_tensor* t) {
  return guarded([&] {
    need(t, "tensor");
    t->t.zero_grad();
  });
}
int agrs_eng_tensor_accumulate_grad(agrs_tensor* t, const agrs_tensor* g) {
  return guarded([&] {
    need(t, "tensor");
    need(g, "grad tensor");
    t->t.accumulate_grad_from_grad_tensor(g->t);
  });
}
int agrs_eng_tensor_backward(agrs_tensor* t) {
  return guarded([&] {
    need(t, "tensor");
    t->t.backward();
  });
}

// ---------------------------------------------------------------- operators
int agrs_eng_op_forward(int op_kind, const int64_t* params, agrs_tensor* const* xs, int n_xs, agrs_tensor** out) {
  return guarded([&] {
    need(out, "out");
    Operators op = make_op(op_kind, params);
    *out = wrap(forward(op, gather(xs, n_xs)));
  });
}
int agrs_eng_op_forward_pair(int kind_first, const int64_t* params_first, int kind_second, const int64_t* params_second,
                             agrs_tensor* const* xs, int n_xs, agrs_tensor** out) {
  return guarded([&] {
    need(out, "out");
    *out = wrap(forward_pair(make_op(kind_first, params_first), make_op(kind_second, params_second), gather(xs, n_xs)));
  });
}
int agrs_eng_op_backward(int op_kind, const int64_t* params, agrs_tensor* const* xs, int n_xs, const agrs_tensor* grad,
                         agrs_tensor** grads_out, int cap, int* n_grads) {
  return guarded([&] {
    need(grad, "grad");
    need(n_grads, "n_grads");
    Operators op = make_op(op_kind, params);
    auto gs = backward(op, gather(xs, n_xs), grad->t);
    if (int(gs.size()) > cap) throw Panic("gradient list buffer too small");
    if (!gs.empty()) need(grads_out, "grads_out");
    for (size_t i = 0; i < gs.size(); ++i) grads_out[i] = wrap(gs[i]);
    *n_grads = int(gs.size());
  });
}
int agrs_eng_im2col(const agrs_tensor* im, int64_t k, int64_t stride, int64_t pad, agrs_tensor** out) {
  return guarded([&] {
    need(im, "tensor");
    need(out, "out");
    *out = wrap(Conv2D::im2col(im->t, k, stride, pad));
  });
}
int agrs_eng_im2col_backward(const agrs_tensor* col, int64_t h_out, int64_t w_out, int64_t stride, int64_t k, agrs_tensor** out) {
  return guarded([&] {
    need(col, "tensor");
    need(out, "out");
    *out = wrap(Conv2D::im2col_backward(col->t, h_out, w_out, stride, k));
  });
}

// ---------------------------------------------------------------- nn
int agrs_eng_sgd_step(agrs_tensor* const* params, int n, float lr) {
  return guarded([&] { nn::sgd_step(gather(params, n), lr); });
}
int agrs_eng_save_parameters(agrs_tensor* const* params, int n, const char* path) {
  return guarded([&] {
    need(path, "path");
    nn::save_parameters(gather(params, n), path);
  });
}
int agrs_eng_load_parameters(agrs_tensor* const* params, int n, const char* path) {
  return guarded([&] {
    need(path, "path");
    nn::load_parameters(gather(params, n), path);
  });
}
int agrs_eng_zero_grad(agrs_tensor* const* params, int n) {
  return guarded([&] { nn::zero_grad(gather(params, n)); });
}

// ---------------------------------------------------------------- data parallel
int agrs_eng_dp_create(agrs_device* d, int world, int rank, const void* unique_id, agrs_dp** out) {
  return guarded([&] {
    need(d, "device");
    need(out, "out");
    need(unique_id, "unique id");
    *out = new agrs_dp(&d->dev, world, rank, unique_id);
  });
}
int agrs_eng_dp_destroy(agrs_dp* dp) {
  return guarded([&] { delete dp; });
}
int agrs_eng_dp_register(agrs_dp* dp, agrs_tensor* const* params, int n, int overlap) {
  return guarded([&] {
    need(dp, "dp");
    dp->dp.register_parameters(gather(params, n), overlap != 0);
  });
}
int agrs_eng_dp_sync_grads(agrs_dp* dp) {
  return guarded([&] {
    need(dp, "dp");
    dp->dp.sync_grads();
  });
}
int agrs_eng_dp_fused_prepare(agrs_dp* dp, void* handle64) {
  return guarded([&] {
    need(dp, "dp");
    need(handle64, "handle buffer");
    dp->dp.fused_prepare(handle64);
  });
}
int agrs_eng_dp_fused_connect(agrs_dp* dp, const void* handles) {
  return guarded([&] {
    need(dp, "dp");
    need(handles, "handles");
    dp->dp.fused_connect(handles);
  });
}
int agrs_eng_dp_fused_disable(agrs_dp* dp) {
  return guarded([&] {
    need(dp, "dp");
    dp->dp.fused_disable();
  });
}
int agrs_eng_dp_fused_begin(agrs_dp* dp, float lr) {
  return guarded([&] {
    need(dp, "dp");
    dp->dp.fused_begin(lr);
  });
}
int agrs_eng_dp_fused_finish(agrs_dp* dp) {
  return guarded([&] {
    need(dp, "dp");
    dp->dp.fused_finish();
  });
}
int agrs_eng_dp_fused_exposed_ms(agrs_dp* dp, double* total_ms, uint64_t* steps) {
  return guarded([&] {
    need(dp, "dp");
    need(total_ms, "out");
    *total_ms = dp->dp.fused_exposed_ms(steps);
  });
}
int agrs_eng_dp_fused_sgd_step(agrs_dp* dp, float lr) {
  return guarded([&] {
    need(dp, "dp");
    dp->dp.fused_sgd_step(lr);
  });
}
int agrs_eng_dp_mean_scalar(agrs_dp* dp, agrs_tensor* t) {
  return guarded([&] {
    need(dp, "dp");
    need(t, "tensor");
    dp->dp.mean_scalar(t->t);
  });
}

}  // extern "C"
