// Operators and the backward walk of the host engine (see autograd.hpp).
#include "autograd.hpp"

namespace agrs {

namespace {

StoragePtr contig(const StoragePtr& s) { return s->transposed ? s->contiguous_copy() : s; }

// C = op(a) * op(b) (+ bias row).  `ta`/`tb` ask for the LOGICAL transpose of the operand; a storage that is
// itself a transposed view flips the flag back.  Nothing is ever copied: majors go to the GEMM as flags.
// `act` != 0 (Linear feeding ReLU / Sigmoid, f2 fusion): the call also produces *y_out = act(C), from the GEMM's epilogue where
// that kernel holds final values, as a second pass inside the same C call otherwise (agrs_gemm_bias_act_f32).
Tensor gemm(const Tensor& a, bool ta, const Tensor& b, bool tb, const float* bias = nullptr, GradCell* scatter_to = nullptr, int act = 0,
            Tensor* y_out = nullptr) {
  if (a.ndim() != 2 || b.ndim() != 2) throw Panic("Ndarray only supports 1d/2d matmul!");
  if (a.dev() != b.dev()) throw Panic("Tensors must be on same device");
  const Shape& as = a.shape();
  const Shape& bs = b.shape();
  const int64_t M = ta ? as[1] : as[0], K = ta ? as[0] : as[1];
  const int64_t K2 = tb ? bs[1] : bs[0], N = tb ? bs[0] : bs[1];
  if (K != K2)
    throw Panic("ndarray: inputs " + std::to_string(M) + " × " + std::to_string(K) + " and " + std::to_string(K2) + " × " +
                    std::to_string(N) + " are not compatible for matrix multiplication",
                AGRS_ERR_SHAPE);
  const bool fa = ta != a.data->transposed, fb = tb != b.data->transposed;  // memory-level flags
  // memory row pitch: A stored [M,K] (fa=0) -> K, stored [K,M] (fa=1) -> M; B stored [K,N] (fb=0) -> N, [N,K] (fb=1) -> K
  auto out = std::make_shared<Storage>(a.dev(), Shape{M, N});
  if (a.dev()->gemm_wants_magnitudes(M, N, K))  // fp16 split mode: one magnitude per operand, kept with the buffer
    a.dev()->check(agrs_gemm_operand_absmax(a.dev()->ctx(), a.data->buf->magnitude(), b.data->buf->magnitude()));
  if (act) {
    auto y = std::make_shared<Storage>(a.dev(), Shape{M, N});
    y->expect_magnitude_report();  // the activation is the next layer's GEMM operand: whoever writes it reports its magnitude
    a.dev()->check(agrs_gemm_bias_act_f32(a.dev()->ctx(), fa, fb, M, N, K, a.data->rptr(), fa ? M : K, b.data->rptr(), fb ? K : N,
                                          out->buf->ptr, N, bias, act, y->buf->ptr, N));
    *y_out = Tensor(y);
  } else if (scatter_to) {
    // data-parallel weight gradient: GEMM -> reduce-scatter in one kernel, the epilogue pushes each tile to its owner over NVLink
    a.dev()->check(agrs_gemm_f32_scatter(a.dev()->ctx(), fa, fb, M, N, K, a.data->rptr(), fa ? M : K, b.data->rptr(), fb ? K : N,
                                         out->buf->ptr, bias, scatter_to->scatter, scatter_to->scatter_off));
    scatter_to->scattered = true;
  } else {
    a.dev()->check(agrs_gemm_f32(a.dev()->ctx(), fa, fb, M, N, K, a.data->rptr(), fa ? M : K, b.data->rptr(), fb ? K : N, out->buf->ptr,
                                 N, bias));
  }
  return Tensor(out);
}

// A freshly computed parameter gradient can be pushed to its owners by the GEMM itself when the coming accumulation is a plain
// hand-over (first touch or lazily zeroed cell) and the parameter takes part in a fused data-parallel exchange.
GradCell* scatter_target(const Tensor& param, int64_t n) {
  GradCell* c = param.grad.get();
  if (c && c->scatter && c->next_accumulate_is_handover() && param.len() == n && !param.data->transposed) return c;
  return nullptr;
}

Tensor fresh_like(const Tensor& x) { return Tensor(std::make_shared<Storage>(x.dev(), x.shape())); }

// f2 fusion: a 2-D gradient [B, O] produced by an activation / loss backward is, in an MLP, what the Linear layer below sums
// over rows for db -- the producing kernel can leave that sum with the buffer.  Returns the [O] storage to fill, or null.
StoragePtr wants_colsum(const Storage& like) {
  if (!like.dev()->fuse_backward_colsum || like.ndim() != 2 || like.transposed) return nullptr;
  const int64_t rows = like.shape[0], cols = like.shape[1];
  if (rows < 64 || cols < 32 || cols % 4 != 0) return nullptr;
  return std::make_shared<Storage>(like.dev(), Shape{cols});
}
void leave_colsum(Storage& out, const StoragePtr& cs) {
  out.buf->colsum = cs;
  out.buf->colsum_cols = out.shape[1];
}
// db = grad.sum_axis(0) (operators.rs:161): taken from the producer when it left one, else one pass over grad
Tensor rows_sum(const Tensor& grad) {
  Buffer& b = *grad.data->buf;
  if (b.colsum && grad.ndim() == 2 && !grad.data->transposed && grad.shapei(1) == b.colsum_cols) {
    Tensor db(b.colsum);
    b.colsum.reset();
    return db;
  }
  return grad.sum_axis(0);
}

// ---------------------------------------------------------------- per-operator forward
Tensor fwd(const ReLU&, std::vector<Tensor>& xs) {
  auto x = contig(xs.at(0).data);
  auto out = std::make_shared<Storage>(x->dev(), x->shape);
  out->expect_magnitude_report();
  x->dev()->check(agrs_relu_fwd_f32(x->dev()->ctx(), x->rptr(), out->buf->ptr, size_t(x->len())));  // operators.rs:112-116
  return Tensor(out);
}

Tensor fwd(const Sigmoid&, std::vector<Tensor>& xs) {
  auto x = contig(xs.at(0).data);
  auto out = std::make_shared<Storage>(x->dev(), x->shape);
  out->expect_magnitude_report();
  x->dev()->check(agrs_sigmoid_fwd_f32(x->dev()->ctx(), x->rptr(), out->buf->ptr, size_t(x->len())));  // operators.rs:195-200
  return Tensor(out);
}

// the bias of a Linear layer as the GEMM epilogue can add it: one contiguous row of w's width
bool bias_is_row(const Tensor& w, const Tensor& b) {
  return w.ndim() == 2 && b.is_contiguous() && b.len() == w.shapei(1) && broadcast_kind(Shape{1, w.shapei(1)}, b.shape()) == Bcast::Same;
}

Tensor fwd(const Linear&, std::vector<Tensor>& xs) {
  // x.dot(w) + b, b added in place into the fresh dot result (operators.rs:140-144): the add is the GEMM epilogue
  const Tensor &x = xs.at(0), &w = xs.at(1), &b = xs.at(2);
  if (bias_is_row(w, b))
    return gemm(x, false, w, false, b.data->rptr());
  return add_inplace(gemm(x, false, w, false), b);
}

Tensor fwd(const MatMul&, std::vector<Tensor>& xs) { return gemm(xs.at(0), false, xs.at(1), false); }  // operators.rs:172-174

Tensor fwd(const MeanSquaredError&, std::vector<Tensor>& xs) {
  // (pred - target).powi_inplace(2).mean(None) reshaped to [1,1] (operators.rs:235-238): one fused reduction
  const Tensor &p = xs.at(0), &t = xs.at(1);
  if (p.shape() == t.shape() && p.is_contiguous() && t.is_contiguous() && p.len() > 0) {
    auto out = std::make_shared<Storage>(p.dev(), Shape{1, 1});
    p.dev()->check(agrs_mse_fwd_f32(p.dev()->ctx(), p.data->rptr(), t.data->rptr(), size_t(p.len()), out->buf->ptr));
    return Tensor(out);
  }
  Tensor m = sub_new(p, t).powi_inplace(2).mean();
  m.reshape({1, 1});
  return m;
}

Tensor fwd(const Mean&, std::vector<Tensor>& xs) { return xs.at(0).mean(); }  // operators.rs:263

Tensor fwd(const Identity&, std::vector<Tensor>& xs) { return xs.at(0).clone(); }  // operators.rs:279: shares data AND grad

Tensor fwd(const Flatten&, std::vector<Tensor>& xs) {
  const Tensor& x = xs.at(0);
  if (x.ndim() < 1) throw Panic("Flatten needs at least one (batch) dimension");
  if (!x.is_contiguous()) throw Panic("Flatten of a transposed view");
  int64_t b = x.shapei(0);
  return Tensor(std::make_shared<Storage>(x.data->buf, Shape{b, b ? x.len() / b : 0}));  // a view: same device buffer
}

Tensor kernel_matrix(const Conv2D& op, const Tensor& filters) {
  // filt [k,k,Co] broadcast over input channels -> [Ci*k*k, Co] (operators.rs:362-376)
  const int64_t kk = op.k_size * op.k_size;
  if (filters.len() != kk * op.out_channels || !filters.is_contiguous())
    throw Panic("called `Option::unwrap()` on a `None` value (cannot broadcast filters " + shape_str(filters.shape()) + " to [" +
                    std::to_string(op.in_channels) + ", " + std::to_string(op.k_size) + ", " + std::to_string(op.k_size) + ", " +
                    std::to_string(op.out_channels) + "])",
                AGRS_ERR_SHAPE);
  if (op.in_channels == 1) return Tensor(std::make_shared<Storage>(filters.data->buf, Shape{kk, op.out_channels}));
  auto km = std::make_shared<Storage>(filters.dev(), Shape{op.in_channels * kk, op.out_channels});
  filters.dev()->check(
      agrs_broadcast_filter_f32(filters.dev()->ctx(), filters.data->rptr(), kk, op.out_channels, op.in_channels, km->buf->ptr));
  return Tensor(km);
}

// the shapes the implicit-GEMM Conv2D kernels take (agrs_conv2d_*_f32)
bool conv_implicit_ok(const Conv2D& op, const Tensor& im, const Tensor& filters, const Tensor& bias) {
  return im.dev()->conv_implicit && im.ndim() == 4 && im.is_contiguous() && filters.is_contiguous() && bias.is_contiguous() &&
         op.in_channels == 1 && im.shapei(1) == 1 && filters.len() == op.k_size * op.k_size * op.out_channels && bias.len() == op.out_channels &&
         agrs_conv2d_implicit_ok(im.shapei(0), 1, im.shapei(2), im.shapei(3), op.k_size, op.out_channels, op.stride, op.pad) != 0;
}

Tensor fwd(const Conv2D& op, std::vector<Tensor>& xs) {
  const Tensor &im = xs.at(0), &filters = xs.at(1), &bias = xs.at(2);
  if (conv_implicit_ok(op, im, filters, bias)) {
    // one input channel: col[b, p, i*k+j] is read from the image inside the kernel; nothing is cached for backward (the node keeps
    // its three inputs; the literal path below pushes col as a 4th variable, operators.rs:386-388)
    auto [h_out, w_out] = op.get_output_shape(im.shapei(2), im.shapei(3));
    auto out = std::make_shared<Storage>(im.dev(), Shape{im.shapei(0), h_out, w_out, op.out_channels});  // NHWC (operators.rs:383)
    im.dev()->check(agrs_conv2d_fwd_f32(im.dev()->ctx(), im.data->rptr(), filters.data->rptr(), bias.data->rptr(), im.shapei(0), im.shapei(2),
                                        im.shapei(3), op.k_size, op.out_channels, op.stride, op.pad, out->buf->ptr));
    return Tensor(out);
  }
  Tensor col = Conv2D::im2col(im, op.k_size, op.stride, op.pad);  // [B, P, C*k*k]
  Tensor kernel = kernel_matrix(op, filters);
  const int64_t kkc = op.k_size * op.k_size * op.in_channels;
  col.reshape({col.shapei(0) * col.shapei(1), kkc});              // operators.rs:378 (panics if C != in_channels)
  Tensor t = (bias.is_contiguous() && bias.len() == op.out_channels) ? gemm(col, false, kernel, false, bias.data->rptr())
                                                                     : add_inplace(gemm(col, false, kernel, false), bias);
  auto [h_out, w_out] = op.get_output_shape(im.shapei(2), im.shapei(3));
  t.reshape({im.shapei(0), h_out, w_out, op.out_channels});       // NHWC (operators.rs:383)
  xs.push_back(col);                                              // cached for backward (operators.rs:386-387)
  return t;
}

// ---------------------------------------------------------------- per-operator backward
std::vector<Tensor> bwd(const ReLU&, const std::vector<Tensor>& xs, const Tensor& grad) {
  // relu_backward_ndarray: g.clone() masked where x <= 0, zipped by iteration order (functional_ndarray.rs:23-34)
  auto x = contig(xs.at(0).data);
  auto g = contig(grad.data);
  if (x->len() != g->len())
    throw Panic("relu backward: grad has " + std::to_string(g->len()) + " elements, input " + std::to_string(x->len()), AGRS_ERR_SHAPE);
  auto out = std::make_shared<Storage>(g->dev(), g->shape);
  out->expect_magnitude_report();
  if (StoragePtr cs = wants_colsum(*out)) {
    g->dev()->check(agrs_relu_bwd_colsum_f32(g->dev()->ctx(), x->rptr(), g->rptr(), out->buf->ptr, out->shape[0], out->shape[1], cs->buf->ptr));
    leave_colsum(*out, cs);
    return {Tensor(out)};
  }
  g->dev()->check(agrs_relu_bwd_f32(g->dev()->ctx(), x->rptr(), g->rptr(), out->buf->ptr, size_t(g->len())));
  return {Tensor(out)};
}

std::vector<Tensor> bwd(const Sigmoid&, const std::vector<Tensor>& xs, const Tensor& grad) {
  // sig = sigmoid(x); dx = ((1 - sig) * sig) * g  (operators.rs:219-223), one fused pass
  auto x = contig(xs.at(0).data);
  auto g = contig(grad.data);
  if (x->len() != g->len()) {
    Tensor sig = fwd(Sigmoid{}, const_cast<std::vector<Tensor>&>(xs));
    Tensor dx = mul_inplace(rsub_scalar(1.0f, sig), sig);
    return {mul_inplace(dx, grad)};
  }
  auto out = std::make_shared<Storage>(x->dev(), x->shape);
  out->expect_magnitude_report();
  if (StoragePtr cs = wants_colsum(*out)) {
    x->dev()->check(agrs_sigmoid_bwd_colsum_f32(x->dev()->ctx(), x->rptr(), g->rptr(), out->buf->ptr, out->shape[0], out->shape[1], cs->buf->ptr));
    leave_colsum(*out, cs);
    return {Tensor(out)};
  }
  x->dev()->check(agrs_sigmoid_bwd_f32(x->dev()->ctx(), x->rptr(), g->rptr(), out->buf->ptr, size_t(x->len())));
  return {Tensor(out)};
}

std::vector<Tensor> bwd(const Linear&, const std::vector<Tensor>& xs, const Tensor& grad) {
  const Tensor &x = xs.at(0), &w = xs.at(1);
  GradCell* wcell = x.ndim() == 2 && grad.ndim() == 2 ? scatter_target(w, x.shapei(1) * grad.shapei(1)) : nullptr;
  if (wcell) {
    // Data-parallel training with the fused exchange: the three results are independent, so the parameter gradients come FIRST
    // and leave for their owners at once; this layer's exchange (and, weights being double-buffered, its update) then runs
    // under the dX GEMM below -- for the first layer that is what hides it, nothing follows dX in the backward walk.
    Tensor dw = gemm(x, true, grad, false, nullptr, wcell);   // x.t_clone().dot(g)    operators.rs:160
    Tensor db = rows_sum(grad);                               // 1-D [O]               operators.rs:161
    if (xs.size() > 2) {
      GradCell* bcell = scatter_target(xs[2], db.len());
      if (bcell && db.is_contiguous()) {
        db.dev()->check(agrs_fused_scatter_f32(bcell->scatter, db.data->rptr(), bcell->scatter_off, size_t(db.len())));
        bcell->scattered = true;
      }
    }
    Tensor dx = gemm(grad, false, w, true);                   // g.dot(&w.t_clone())   operators.rs:159
    return {dx, dw, db};
  }
  Tensor dx = gemm(grad, false, w, true);   // g.dot(&w.t_clone())   operators.rs:159
  Tensor dw = gemm(x, true, grad, false);   // x.t_clone().dot(g)    operators.rs:160
  Tensor db = rows_sum(grad);               // 1-D [O]               operators.rs:161 (left by the producer of grad when it could)
  return {dx, dw, db};
}

std::vector<Tensor> bwd(const MatMul&, const std::vector<Tensor>& xs, const Tensor& grad) {
  const Tensor &x = xs.at(0), &w = xs.at(1);
  if (x.dev()->matmul_backward_literal)  // reference defect D3 (operators.rs:183-184): no transposes on the weight side
    return {gemm(grad, false, w, false), gemm(grad, true, x, false)};
  return {gemm(grad, false, w, true), gemm(x, true, grad, false)};
}

std::vector<Tensor> bwd(const MeanSquaredError&, const std::vector<Tensor>& xs, const Tensor& grad) {
  // dpred = (pred - target) * 2.0 / B * grad, B = pred.shape[0] (operators.rs:247-253; literal D4)
  const Tensor &p = xs.at(0), &t = xs.at(1);
  if (p.ndim() < 1) throw Panic("index out of bounds: the len is 0 but the index is 0");
  const float B = float(p.shapei(0));
  if (p.shape() == t.shape() && p.is_contiguous() && t.is_contiguous() && grad.len() == 1) {
    Tensor out = fresh_like(p);
    out.data->expect_magnitude_report();
    if (StoragePtr cs = wants_colsum(*out.data)) {
      p.dev()->check(agrs_mse_bwd_colsum_f32(p.dev()->ctx(), p.data->rptr(), t.data->rptr(), grad.data->rptr(), B, p.shapei(0), p.shapei(1),
                                             out.data->buf->ptr, cs->buf->ptr));
      leave_colsum(*out.data, cs);
      return {out};
    }
    p.dev()->check(agrs_mse_bwd_f32(p.dev()->ctx(), p.data->rptr(), t.data->rptr(), grad.data->rptr(), B, size_t(p.len()), out.data->buf->ptr));
    return {out};
  }
  Tensor d = div_scalar(mul_scalar(sub_new(p, t), 2.0f), B);
  return {mul_inplace(d, grad)};
}

std::vector<Tensor> bwd(const Mean&, const std::vector<Tensor>& xs, const Tensor& grad) {
  const Tensor& x = xs.at(0);  // ones_like(x) * 1.0 / size * grad (operators.rs:268-270)
  Tensor t = Tensor::full(x.dev(), x.shape(), 1.0f / float(x.size()));
  return {mul_inplace(t, grad)};
}

std::vector<Tensor> bwd(const Identity&, const std::vector<Tensor>&, const Tensor& grad) { return {grad}; }

std::vector<Tensor> bwd(const Flatten&, const std::vector<Tensor>& xs, const Tensor& grad) {
  if (grad.len() != xs.at(0).len() || !grad.is_contiguous()) throw Panic("Flatten backward: gradient does not match the input", AGRS_ERR_SHAPE);
  return {Tensor(std::make_shared<Storage>(grad.data->buf, xs.at(0).shape()))};
}

std::vector<Tensor> bwd(const Conv2D& op, const std::vector<Tensor>& xs, const Tensor& grad) {
  const Tensor &im = xs.at(0), &filters = xs.at(1);
  if (grad.ndim() != 4) throw Panic("index out of bounds: Conv2D backward expects a BxH_outxW_outxC_out gradient", AGRS_ERR_SHAPE);
  Tensor g = grad.clone();
  g.reshape({g.shapei(0) * g.shapei(1) * g.shapei(2), g.shapei(3)});  // mutates the shared storage shape, as operators.rs:402-403
  if (xs.size() == 3 && conv_implicit_ok(op, im, filters, xs.at(2)) && g.is_contiguous() &&
      g.shapei(1) == op.out_channels) {
    // forward cached no col: db, dw and dx straight from the image and the gradient, one kernel (per-sample overlap-add: D1)
    auto [ho, wo] = op.get_output_shape(im.shapei(2), im.shapei(3));
    if (g.shapei(0) != im.shapei(0) * ho * wo) throw Panic("Conv2D backward: gradient does not match the output shape", AGRS_ERR_SHAPE);
    auto dx = std::make_shared<Storage>(im.dev(), Shape{im.shapei(0), 1, (ho - 1) * op.stride + op.k_size, (wo - 1) * op.stride + op.k_size});
    auto dw = std::make_shared<Storage>(im.dev(), Shape{op.k_size, op.k_size, op.out_channels});
    auto db = std::make_shared<Storage>(im.dev(), Shape{op.out_channels});
    im.dev()->check(agrs_conv2d_bwd_f32(im.dev()->ctx(), im.data->rptr(), filters.data->rptr(), g.data->rptr(), im.shapei(0), im.shapei(2), im.shapei(3),
                                        op.k_size, op.out_channels, op.stride, op.pad, dx->buf->ptr, dw->buf->ptr, db->buf->ptr));
    return {Tensor(dx), Tensor(dw), Tensor(db)};
  }
  if (xs.size() < 4) throw Panic("index out of bounds: the len is 3 but the index is 3 (Conv2D backward without the cached col)");
  Tensor db = g.sum_axis(0);                                         // :405
  Tensor kernels = kernel_matrix(op, filters);                       // :408-419
  Tensor dxcol = gemm(g, false, kernels, true);                      // g.dot(kernels.t())  :422
  auto [h_out, w_out] = op.get_output_shape(im.shapei(2), im.shapei(3));
  // The reference views dxcol as [1, B*P, KKC] and its col2im walks i in 0..B*P with y = i / W_out: that panics for
  // B > 1 (defect D1).  Per-sample generalisation: view as [B, P, KKC]; identical to the literal code at B = 1.
  dxcol.reshape({im.shapei(0), h_out * w_out, dxcol.shapei(1)});
  Tensor dx = Conv2D::im2col_backward(dxcol, h_out, w_out, op.stride, op.k_size);  // padded-size dx (TODO at :427 kept)
  Tensor col = xs.at(3).clone();
  col.reshape({col.shapei(0), col.shapei(1) / op.in_channels, op.in_channels});    // :438-442, also mutates the cached tensor
  Tensor csum = op.in_channels == 1 ? Tensor(std::make_shared<Storage>(col.data->buf, Shape{col.shapei(0), col.shapei(1)}))
                                    : col.sum_axis(-1);                             // :443
  Tensor dw = gemm(csum, true, g, false);                                          // col.t().dot(&g)  :445
  dw.reshape({op.k_size, op.k_size, op.out_channels});                             // :446
  return {dx, dw, db};
}

}  // namespace

// ---------------------------------------------------------------- Conv2D helpers
std::pair<int64_t, int64_t> Conv2D::get_output_shape(int64_t h, int64_t w) const {
  if (h + 2 * pad < k_size || w + 2 * pad < k_size) throw Panic("attempt to subtract with overflow", AGRS_ERR_SHAPE);
  return {(h - k_size + 2 * pad) / stride + 1, (w - k_size + 2 * pad) / stride + 1};
}

Tensor Conv2D::im2col(const Tensor& im, int64_t k, int64_t stride, int64_t pad) {
  if (im.ndim() != 4) throw Panic("Expected image of shape BCHW, got shape " + shape_str(im.shape()), AGRS_ERR_SHAPE);  // functional_ndarray.rs:100-103
  auto x = contig(im.data);
  const int64_t B = x->shape[0], C = x->shape[1], H = x->shape[2], W = x->shape[3];
  if (H + 2 * pad < k || W + 2 * pad < k) throw Panic("attempt to subtract with overflow", AGRS_ERR_SHAPE);
  const int64_t ho = (H - k + 2 * pad) / stride + 1, wo = (W - k + 2 * pad) / stride + 1;
  auto col = std::make_shared<Storage>(x->dev(), Shape{B, ho * wo, C * k * k});
  x->dev()->check(agrs_im2col_f32(x->dev()->ctx(), x->rptr(), B, C, H, W, k, stride, pad, col->buf->ptr));
  return Tensor(col);
}

Tensor Conv2D::im2col_backward(const Tensor& col, int64_t h_out, int64_t w_out, int64_t stride, int64_t k) {
  if (col.ndim() != 3)
    throw Panic("Expected tensor of shape BxH_out*W_outxK*K*C, got shape " + shape_str(col.shape()), AGRS_ERR_SHAPE);  // :134-137
  auto c = contig(col.data);
  const int64_t B = c->shape[0], P = c->shape[1], K = c->shape[2];
  if (P != h_out * w_out) throw Panic("range end index out of range for slice (col2im: P != H_out*W_out)", AGRS_ERR_SHAPE);
  const int64_t C = K / (k * k);
  const int64_t h = (h_out - 1) * stride + k, w = (w_out - 1) * stride + k;
  auto im = std::make_shared<Storage>(c->dev(), Shape{B, C, h, w});
  c->dev()->check(agrs_col2im_f32(c->dev()->ctx(), c->rptr(), B, C, h_out, w_out, k, stride, im->buf->ptr));
  return Tensor(im);
}

// ---------------------------------------------------------------- dispatch
std::string operator_name(const Operators& op) {
  static const char* names[] = {"ReLU", "Sigmoid", "Linear", "MatMul", "MeanSquaredError", "Mean", "Identity", "Conv2D", "Flatten"};
  return names[op.index()];
}

bool Node::is_root_node() const {
  for (auto& p : parents)
    if (p) return false;
  return true;
}

void attach_to_eager_graph(std::vector<Tensor>& inputs, Tensor& op_output, const Operators& op) {
  std::vector<std::shared_ptr<Node>> parents(inputs.size());
  for (size_t i = 0; i < inputs.size(); ++i) {
    parents[i] = inputs[i].graph;  // Some only for inputs generated by another operator
    inputs[i].graph = nullptr;     // on the local handle only (operators.rs:41)
  }
  auto node = std::make_shared<Node>();
  node->op = op;
  node->variables = inputs;
  node->parents = std::move(parents);
  op_output.graph = node;
}

Tensor forward(const Operators& op, std::vector<Tensor> xs) {
  Tensor t = std::visit([&](auto const& o) { return fwd(o, xs); }, op);
  if (std::holds_alternative<Identity>(op)) t.graph = nullptr;
  attach_to_eager_graph(xs, t, op);
  return t;
}

// Linear followed by ReLU / Sigmoid (the Layer loop of layers.rs:24-43 on [Linear, act], or two consecutive layers of a model):
// the same two graph nodes -- the activation's node keeps the pre-activation as its variable (operators.rs:120-129, :214-224) --
// from ONE GEMM call whose epilogue writes the activation next to the pre-activation.
Tensor forward_pair(const Operators& first, const Operators& second, std::vector<Tensor> xs) {
  const int act = std::holds_alternative<ReLU>(second) ? AGRS_ACT_RELU : std::holds_alternative<Sigmoid>(second) ? AGRS_ACT_SIGMOID : AGRS_ACT_NONE;
  if (act && std::holds_alternative<Linear>(first) && xs.size() == 3 && xs[0].dev()->fuse_forward_activation && xs[0].ndim() == 2 &&
      xs[1].dev() == xs[0].dev() && bias_is_row(xs[1], xs[2])) {
    Tensor y;
    Tensor z = gemm(xs[0], false, xs[1], false, xs[2].data->rptr(), nullptr, act, &y);
    attach_to_eager_graph(xs, z, first);
    std::vector<Tensor> zs{z};
    attach_to_eager_graph(zs, y, second);
    return y;
  }
  if (act && std::holds_alternative<Conv2D>(first) && xs.size() == 3 && xs[0].dev()->fuse_forward_activation &&
      conv_implicit_ok(std::get<Conv2D>(first), xs[0], xs[1], xs[2])) {
    // the implicit-GEMM Conv2D kernel streams its staged output out once: the same stream writes the activation next to it
    const Conv2D& op = std::get<Conv2D>(first);
    const Tensor &im = xs[0], &filters = xs[1], &bias = xs[2];
    auto [h_out, w_out] = op.get_output_shape(im.shapei(2), im.shapei(3));
    const Shape shape{im.shapei(0), h_out, w_out, op.out_channels};  // NHWC (operators.rs:383)
    auto zs_ = std::make_shared<Storage>(im.dev(), shape);
    auto ys_ = std::make_shared<Storage>(im.dev(), shape);
    ys_->expect_magnitude_report();
    im.dev()->check(agrs_conv2d_fwd_act_f32(im.dev()->ctx(), im.data->rptr(), filters.data->rptr(), bias.data->rptr(), im.shapei(0), im.shapei(2),
                                            im.shapei(3), op.k_size, op.out_channels, op.stride, op.pad, zs_->buf->ptr, act, ys_->buf->ptr));
    Tensor z(zs_), y(ys_);
    attach_to_eager_graph(xs, z, first);
    std::vector<Tensor> zs{z};
    attach_to_eager_graph(zs, y, second);
    return y;
  }
  Tensor z = forward(first, std::move(xs));
  return forward(second, {z});
}

std::vector<Tensor> backward(const Operators& op, const std::vector<Tensor>& xs, const Tensor& grad) {
  return std::visit([&](auto const& o) { return bwd(o, xs, grad); }, op);
}

void backward_algo(const std::shared_ptr<Node>& node, const Tensor& prev_grad) {
  // 1. gradient(s) of this operator w.r.t. its inputs (autograd.rs:57-69)
  std::vector<Tensor> grads = backward(node->op, node->variables, prev_grad);
  // 2. accumulate on the input variables; zip stops at the shorter list (:73-79)
  const size_t n = grads.size() < node->variables.size() ? grads.size() : node->variables.size();
  for (size_t i = 0; i < n; ++i) {
    const Tensor& var = node->variables[i];
    if (!var.requires_grad) continue;
    var.accumulate_grad_from_grad_tensor(grads[i]);
    const bool has_parent = i < node->parents.size() && node->parents[i];
    if (!has_parent) grads[i] = Tensor();  // a leaf: the grad cell is now the only owner of the handed-over buffer
    if (var.grad->hook) var.grad->hook->on_accumulated(*var.grad);
  }
  // 3. recurse on parent nodes (:81-85)
  for (size_t i = 0; i < node->parents.size(); ++i)
    if (node->parents[i]) {
      if (i >= grads.size()) throw Panic("index out of bounds: the len is " + std::to_string(grads.size()) + " but the index is " + std::to_string(i));
      backward_algo(node->parents[i], grads[i].clone());
    }
}

}  // namespace agrs
