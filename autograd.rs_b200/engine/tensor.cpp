// Tensor layer of the host engine (see tensor.hpp).  Every arithmetic call below is one or two
// asynchronous launches through the C ABI; only to_host()/item()/sum()/equal() block.
#include "tensor.hpp"

#include <cmath>
#include <sstream>

#include "autograd.hpp"

namespace agrs {

int64_t numel(const Shape& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

std::string shape_str(const Shape& s) {
  std::ostringstream o;
  o << "[";
  for (size_t i = 0; i < s.size(); ++i) o << (i ? ", " : "") << s[i];
  o << "]";
  return o.str();
}

// ------------------------------------------------------------------ Device
Device::Device(int ordinal, void* stream, bool on_stream) {
  int rc = on_stream ? agrs_ctx_create_on_stream(ordinal, stream, &ctx_) : agrs_ctx_create(ordinal, &ctx_);
  if (rc != AGRS_OK)
    throw Panic("no sm_100 CUDA device available for ordinal " + std::to_string(ordinal) + " (the engine has no CPU fallback)", rc);
}
Device::~Device() {
  if (ctx_) agrs_ctx_destroy(ctx_);
}
void Device::check(int rc) const {
  if (rc != AGRS_OK) throw Panic(agrs_last_error(ctx_), rc);
}
float Device::next_uniform(float lo, float hi) {
  uint64_t x = rng_state_;
  x ^= x >> 12;
  x ^= x << 25;
  x ^= x >> 27;
  rng_state_ = x;
  uint64_t r = x * 0x2545F4914F6CDD1Dull;
  float u = float(r >> 40) * (1.0f / 16777216.0f);  // 24 random bits -> [0,1)
  return lo + (hi - lo) * u;
}

// ------------------------------------------------------------------ Buffer / Storage
Buffer::Buffer(Device* d, int64_t n_elems) : dev(d), keep(d->weak_from_this().lock()), n(n_elems) {
  void* p = nullptr;
  dev->check(agrs_alloc(dev->ctx(), size_t(n_elems) * sizeof(float), &p));
  ptr = static_cast<float*>(p);
}
Buffer::Buffer(Device* d, float* external, int64_t n_elems) : dev(d), keep(d->weak_from_this().lock()), ptr(external), n(n_elems), owned(false) {}
Buffer::~Buffer() {
  if (landing) agrs_event_destroy(dev->ctx(), landing);
  if (mx) agrs_free(dev->ctx(), mx);
  if (ptr && owned) agrs_free(dev->ctx(), ptr);
}
uint32_t* Buffer::mx_slot() {
  if (!mx) {
    void* p = nullptr;
    dev->check(agrs_alloc(dev->ctx(), sizeof(uint32_t), &p));
    mx = static_cast<uint32_t*>(p);
  }
  return mx;
}
void Buffer::await_landing() {
  if (!landing_pending) return;
  landing_pending = false;
  dev->check(agrs_wait_event(dev->ctx(), landing));
}
const uint32_t* Buffer::magnitude() {
  if (!mx_valid) {
    await_landing();
    dev->check(agrs_absmax_f32(dev->ctx(), ptr, size_t(n), mx_slot()));
    mx_valid = true;
  }
  return mx;
}
void Storage::expect_magnitude_report() {
  if (buf->n < (int64_t(1) << 16) || !dev()->reports_magnitudes()) return;
  dev()->check(agrs_next_output_absmax(dev()->ctx(), buf->mx_slot()));
  buf->mx_valid = true;
}

Storage::Storage(Device* d, const Shape& s) : buf(std::make_shared<Buffer>(d, numel(s))), shape(s) {
  for (auto v : s)
    if (v < 0) throw Panic("negative dimension in shape " + shape_str(s));
}

float* Storage::wptr() {
  buf->await_landing();
  if (buf.use_count() > 1) {  // aliased by a gradient hand-over: un-share before writing (the reference deep-copied up front)
    auto nb = std::make_shared<Buffer>(dev(), buf->n);
    dev()->check(agrs_copy(dev()->ctx(), nb->ptr, buf->ptr, size_t(buf->n) * sizeof(float)));
    buf = nb;
  }
  buf->mx_valid = false;  // about to be overwritten
  buf->colsum.reset();
  return buf->ptr;
}

StoragePtr Storage::deep_copy() const {
  auto nb = std::make_shared<Buffer>(dev(), buf->n);
  dev()->check(agrs_copy(dev()->ctx(), nb->ptr, rptr(), size_t(buf->n) * sizeof(float)));
  return std::make_shared<Storage>(nb, shape, transposed);
}

StoragePtr Storage::contiguous_copy() const {
  if (!transposed) return deep_copy();
  auto out = std::make_shared<Storage>(dev(), shape);
  // memory is [shape[1], shape[0]] row-major; the logical array is its transpose
  dev()->check(agrs_transpose_f32(dev()->ctx(), rptr(), shape[1], shape[0], out->buf->ptr));
  return out;
}

std::vector<float> Storage::to_host() const {
  std::vector<float> h(static_cast<size_t>(len()), 0.0f);
  if (h.empty()) return h;
  if (transposed) {
    auto c = contiguous_copy();
    dev()->check(agrs_download(dev()->ctx(), h.data(), c->rptr(), h.size() * sizeof(float)));
  } else {
    dev()->check(agrs_download(dev()->ctx(), h.data(), rptr(), h.size() * sizeof(float)));
  }
  return h;
}

static StoragePtr contig(const StoragePtr& s) { return s->transposed ? s->contiguous_copy() : s; }

// ------------------------------------------------------------------ broadcasting
Bcast broadcast_kind(const Shape& a, const Shape& b) {
  if (b.size() > a.size()) throw Panic("could not broadcast array from shape: " + shape_str(b) + " to: " + shape_str(a), AGRS_ERR_SHAPE);
  // align trailing dims (ndarray rule): every b dim equals a's or is 1
  size_t off = a.size() - b.size();
  bool same = true;
  for (size_t i = 0; i < b.size(); ++i) {
    if (b[i] != a[off + i]) {
      same = false;
      if (b[i] != 1) throw Panic("could not broadcast array from shape: " + shape_str(b) + " to: " + shape_str(a), AGRS_ERR_SHAPE);
    }
  }
  if (same && numel(a) == numel(b)) return Bcast::Same;
  if (numel(a) == numel(b)) return Bcast::Same;  // only 1-sized dims differ
  // trailing: the non-1 part of b is a suffix of a  ([O] / [1,O] / [1,1] onto [B,O])
  size_t first = 0;
  while (first < b.size() && b[first] == 1) ++first;
  bool trailing = true;
  for (size_t i = first; i < b.size(); ++i)
    if (b[i] != a[off + i]) trailing = false;
  if (trailing) return Bcast::Trailing;
  // leading: b matches a on a prefix of the aligned dims and is 1 afterwards ([B,1] onto [B,O]); needs equal rank
  if (off == 0) {
    size_t last = b.size();
    while (last > 0 && b[last - 1] == 1) --last;
    bool leading = true;
    for (size_t i = 0; i < last; ++i)
      if (b[i] != a[i]) leading = false;
    if (leading) return Bcast::Leading;
  }
  throw Panic("broadcast of " + shape_str(b) + " onto " + shape_str(a) + " is not mapped to a device kernel", AGRS_ERR_UNSUPPORTED);
}

void storage_binary_into(int op, Storage& out, const Storage& a, const Storage& b) {
  if (a.transposed || b.transposed || out.transposed) throw Panic("elementwise arithmetic on a transposed view: make it contiguous first");
  Bcast k = broadcast_kind(a.shape, b.shape);
  Device* d = a.dev();
  const float* ap = a.rptr();
  const float* bp = b.rptr();
  float* op_ = (&out == &a) ? out.wptr() : out.buf->ptr;
  if (&out == &a) ap = op_;
  d->check(agrs_binary_f32(d->ctx(), op, op_, ap, bp, size_t(a.len()), size_t(b.len()),
                           k == Bcast::Leading ? AGRS_BCAST_LEADING : AGRS_BCAST_TRAILING));
}

// ------------------------------------------------------------------ constructors
Tensor Tensor::from(Device* d, const float* host, const Shape& shape) {
  auto s = std::make_shared<Storage>(d, shape);
  if (s->len()) d->check(agrs_upload(d->ctx(), s->buf->ptr, host, size_t(s->len()) * sizeof(float)));
  d->sync();  // `host` may be a temporary (pageable): make the call self-contained
  return Tensor(s);
}
Tensor Tensor::full(Device* d, const Shape& shape, float v) {
  auto s = std::make_shared<Storage>(d, shape);
  d->check(agrs_fill_f32(d->ctx(), s->buf->ptr, v, size_t(s->len())));
  return Tensor(s);
}
Tensor Tensor::zeros(Device* d, const Shape& shape) { return full(d, shape, 0.0f); }
Tensor Tensor::ones(Device* d, const Shape& shape) { return full(d, shape, 1.0f); }

Tensor Tensor::uniform(Device* d, const Shape& shape, float low, float high) {
  std::vector<float> h(static_cast<size_t>(numel(shape)), 0.0f);
  for (auto& v : h) v = d->next_uniform(low, high);
  return from(d, h, shape);
}

static std::pair<int64_t, int64_t> fan_in_out(const Shape& s) {  // init.rs:30-45 (fan_in uses dim 1)
  if (s.size() < 2) throw Panic("assertion failed: tensor_shape.len() >= 2");
  int64_t rf = 1;
  for (size_t i = 2; i < s.size(); ++i) rf *= s[i];
  return {s[1] * rf, s[0] * rf};
}
std::pair<int64_t, int64_t> calculate_fan_in_and_fan_out(const Shape& s) { return fan_in_out(s); }

Tensor Tensor::kaiming_uniform(Device* d, const Shape& shape) {  // init.rs:19-28
  float negative_slope = std::sqrt(5.0f);
  float gain = std::sqrt(2.0f / (1.0f + negative_slope * negative_slope));
  float stdv = gain / std::sqrt(float(fan_in_out(shape).first));
  float bound = std::sqrt(3.0f) * stdv;
  return uniform(d, shape, -bound, bound);
}

Tensor Tensor::to_owned() const {
  Tensor t = *this;
  t.graph = nullptr;
  return t;
}

std::vector<float> Tensor::grad_to_vec() const {
  if (!has_grad()) throw Panic("called `Option::unwrap()` on a `None` value (tensor has no grad)");
  return grad->get()->to_host();
}

float Tensor::item() const {
  if (len() != 1) throw Panic("item() on a tensor with " + std::to_string(len()) + " elements");
  float v;
  dev()->check(agrs_download(dev()->ctx(), &v, data->rptr(), sizeof(float)));
  return v;
}

// ------------------------------------------------------------------ shape ops
void Tensor::reshape(const Shape& shape) {
  if (!data->is_contiguous()) throw Panic("called `Result::unwrap()` on an `Err` value: ShapeError/IncompatibleLayout: incompatible memory layout");
  if (numel(shape) != data->len())
    throw Panic("called `Result::unwrap()` on an `Err` value: ShapeError/IncompatibleShape: cannot reshape " + shape_str(data->shape) +
                    " into " + shape_str(shape),
                AGRS_ERR_SHAPE);
  data->shape = shape;  // the cell is shared: every clone sees the new shape (tensor.rs:493-502)
}
Tensor Tensor::unsqueeze(int ax) const {
  Tensor t = *this;
  int a = ax < 0 ? int(ndim()) : ax;
  if (size_t(a) > ndim()) throw Panic("assertion failed: ax <= self.ndim()");
  Shape s = shape();
  if (size_t(a) < s.size()) s.insert(s.begin() + a, 1); else s.push_back(1);
  t.reshape(s);
  return t;
}
Tensor Tensor::squeeze() const {
  Tensor t = *this;
  Shape s;
  for (auto d : shape())
    if (d != 1) s.push_back(d);
  t.reshape(s);
  return t;
}
void Tensor::swap_axes(int a, int b) const {
  int nd = int(ndim());
  if (a < 0) a += nd;
  if (b < 0) b += nd;
  if (a < 0 || b < 0 || a >= nd || b >= nd) throw Panic("swap_axes: axis out of range");
  if (a == b) return;
  if (nd != 2) throw Panic("swap_axes on a " + std::to_string(nd) + "-d device tensor is not supported (2-d only)", AGRS_ERR_UNSUPPORTED);
  std::swap(data->shape[0], data->shape[1]);
  data->transposed = !data->transposed;
}
const Tensor& Tensor::t() const {
  if (ndim() != 2) throw Panic("assertion `left == right` failed\n  left: " + std::to_string(ndim()) + "\n right: 2");
  swap_axes(0, -1);
  return *this;
}
Tensor Tensor::t_clone() const {
  // x.t().to_owned(): same bytes, transposed logical view (storage.rs:84-88); GEMMs consume it as an operand major
  if (ndim() != 2) {
    if (ndim() < 2) return Tensor(data->deep_copy());
    throw Panic("t_clone on a " + std::to_string(ndim()) + "-d device tensor is not supported", AGRS_ERR_UNSUPPORTED);
  }
  auto s = data->deep_copy();
  std::swap(s->shape[0], s->shape[1]);
  s->transposed = !s->transposed;
  return Tensor(s);
}

// ------------------------------------------------------------------ reductions / maps
float Tensor::sum() const {
  Device* d = dev();
  Buffer out(d, 1);
  d->check(agrs_sum_f32(d->ctx(), data->rptr(), size_t(len()), out.ptr));
  float v;
  d->check(agrs_download(d->ctx(), &v, out.ptr, sizeof(float)));
  return v;
}
Tensor Tensor::sum_axis(int axis) const {
  int nd = int(ndim());
  if (axis < 0) axis += nd;
  if (axis < 0 || axis >= nd) throw Panic("sum_axis: axis out of range");
  auto src = contig(data);
  int64_t outer = 1, inner = 1;
  for (int i = 0; i < axis; ++i) outer *= src->shape[i];
  for (int i = axis + 1; i < nd; ++i) inner *= src->shape[i];
  Shape os;
  for (int i = 0; i < nd; ++i)
    if (i != axis) os.push_back(src->shape[i]);
  auto out = std::make_shared<Storage>(dev(), os);
  dev()->check(agrs_sum_axis_f32(dev()->ctx(), src->rptr(), outer, src->shape[axis], inner, out->buf->ptr));
  return Tensor(out);
}
Tensor Tensor::mean() const {
  if (len() == 0) throw Panic("called `Option::unwrap()` on a `None` value (mean of an empty tensor)");
  auto out = std::make_shared<Storage>(dev(), Shape{1});
  dev()->check(agrs_mean_f32(dev()->ctx(), data->rptr(), size_t(len()), out->buf->ptr));
  return Tensor(out);
}
Tensor Tensor::mean_axis(int axis) const {
  int nd = int(ndim());
  if (axis < 0) axis += nd;
  Tensor s = sum_axis(axis);
  float n = float(shape().at(size_t(axis)));
  Device* d = dev();
  d->check(agrs_scalar_f32(d->ctx(), AGRS_SDIV, s.data->buf->ptr, s.data->rptr(), n, size_t(s.len())));
  return s;
}
Tensor Tensor::powi(int exp) const {
  auto out = std::make_shared<Storage>(dev(), shape());
  out->transposed = data->transposed;
  dev()->check(agrs_powi_f32(dev()->ctx(), out->buf->ptr, data->rptr(), exp, size_t(len())));
  return Tensor(out);
}
Tensor Tensor::powi_inplace(int exp) const {
  float* p = data->wptr();
  dev()->check(agrs_powi_f32(dev()->ctx(), p, p, exp, size_t(len())));
  return *this;
}
void Tensor::fill(float v) const { dev()->check(agrs_fill_f32(dev()->ctx(), data->wptr(), v, size_t(len()))); }

Tensor Tensor::dot(const Tensor& other) const {
  if (ndim() > 2 || other.ndim() > 2) throw Panic("Ndarray only supports 1d/2d matmul!");
  if (ndim() != 2 || other.ndim() != 2)
    throw Panic("called `Result::unwrap()` on an `Err` value: ShapeError/IncompatibleShape: incompatible shapes (dot needs 2-d operands)",
                AGRS_ERR_SHAPE);
  if (dev() != other.dev()) throw Panic("Tensors must be on same device");
  const int64_t M = shapei(0), K = shapei(1), K2 = other.shapei(0), N = other.shapei(1);
  if (K != K2)
    throw Panic("ndarray: inputs " + std::to_string(M) + " × " + std::to_string(K) + " and " + std::to_string(K2) + " × " +
                    std::to_string(N) + " are not compatible for matrix multiplication",
                AGRS_ERR_SHAPE);
  auto out = std::make_shared<Storage>(dev(), Shape{M, N});
  const bool ta = data->transposed, tb = other.data->transposed;
  dev()->check(agrs_gemm_f32(dev()->ctx(), ta, tb, M, N, K, data->rptr(), ta ? M : K, other.data->rptr(), tb ? K : N, out->buf->ptr, N,
                             nullptr));
  return Tensor(out);
}

// ------------------------------------------------------------------ autograd plumbing
const StoragePtr& GradCell::get() {
  if (g && zero_pending) {
    zero_pending = false;
    g->dev()->check(agrs_fill_f32(g->dev()->ctx(), g->wptr(), 0.0f, size_t(g->len())));
  }
  return g;
}

void Tensor::zero_grad() const {
  // grad = Some(zeros(shape of DATA)) (tensor.rs:135-140).  The zeros are LAZY: the cell records "all zero" and the memset
  // only runs if somebody reads the gradient before the next accumulation (GradCell::get); the usual next event is
  // backward's `zeros += dW`, which then degenerates to taking dW's buffer (0 + x == x) -- no memset, no add pass.
  auto& g = grad->g;
  if (g && g->len() == data->len() && g->buf.use_count() == 1) {
    g->shape = data->shape;
    g->transposed = false;
  } else {
    g = std::make_shared<Storage>(dev(), data->shape);
  }
  grad->zero_pending = true;
}

void Tensor::accumulate_grad_from_grad_tensor(const Tensor& gt) const {
  auto& g = grad->g;
  auto src = contig(gt.data);
  if (!g) {
    // None => Some(grad.to_owned()).  The device buffer is handed over instead of copied; whoever
    // writes in place first (Storage::wptr) un-shares it, so the observable semantics are the copy's.
    g = std::make_shared<Storage>(src->buf, src->shape, false);
  } else if (grad->zero_pending && src->len() == g->len() && broadcast_kind(g->shape, src->shape) == Bcast::Same) {
    // zeros(shape) += b with nothing to broadcast: the result is b's values in the zeros' shape ([1,O] += [O] keeps [1,O])
    g = std::make_shared<Storage>(src->buf, g->shape, false);
    grad->zero_pending = false;
  } else {
    // Some(a) => a += b with ndarray broadcasting of b onto a (todo!() at HEAD: defect D2)
    grad->get();
    storage_binary_into(AGRS_ADD, *g, *g, *src);
    grad->scattered = false;  // what a scattering GEMM pushed to the owners was only the new term: the sum must be pushed again
  }
}

void Tensor::accumulate_grad(const Tensor& other) const {
  if (!other.has_grad()) return;
  Tensor tmp(other.grad->get());
  accumulate_grad_from_grad_tensor(tmp);
}

void Tensor::backward() const {
  if (!requires_grad) throw Panic("not implemented");  // unimplemented!() tensor.rs:329-331
  if (!graph) throw Panic("Variable has no attached graph");
  backward_algo(graph, ones_like(*this));  // seed lives on the device: no host round trip
}

Tensor ones_like(const Tensor& t) { return Tensor::ones(t.dev(), t.shape()); }
Tensor zeros_like(const Tensor& t) { return Tensor::zeros(t.dev(), t.shape()); }

// ------------------------------------------------------------------ arithmetic
static Tensor binary_new(int op, const Tensor& a, const Tensor& b) {
  if (a.dev() != b.dev()) throw Panic("Tensors must be on same device");
  auto A = contig(a.data), B = contig(b.data);
  // ndarray co-broadcasts &a op &b; take the larger operand's shape as the result
  bool swap = false;
  if (numel(B->shape) > numel(A->shape) || (numel(B->shape) == numel(A->shape) && B->shape.size() > A->shape.size())) swap = true;
  if (!swap) {
    auto out = std::make_shared<Storage>(a.dev(), A->shape);
    storage_binary_into(op, *out, *A, *B);
    return Tensor(out);
  }
  auto out = std::make_shared<Storage>(a.dev(), B->shape);
  storage_binary_into(op, *out, *B, *A);  // computes b op a
  if (op == AGRS_SUB) a.dev()->check(agrs_neg_f32(a.dev()->ctx(), out->buf->ptr, out->buf->ptr, size_t(out->len())));
  return Tensor(out);
}
Tensor add_new(const Tensor& a, const Tensor& b) { return binary_new(AGRS_ADD, a, b); }
Tensor sub_new(const Tensor& a, const Tensor& b) { return binary_new(AGRS_SUB, a, b); }
Tensor mul_new(const Tensor& a, const Tensor& b) { return binary_new(AGRS_MUL, a, b); }

static Tensor binary_inplace(int op, Tensor a, const Tensor& b) {
  if (a.dev() != b.dev()) throw Panic("Tensors must be on same device");
  if (a.data->transposed) throw Panic("in-place arithmetic on a transposed view is not supported", AGRS_ERR_UNSUPPORTED);
  auto B = contig(b.data);
  storage_binary_into(op, *a.data, *a.data, *B);
  return a;
}
Tensor add_inplace(Tensor a, const Tensor& b) { return binary_inplace(AGRS_ADD, std::move(a), b); }
Tensor sub_inplace(Tensor a, const Tensor& b) { return binary_inplace(AGRS_SUB, std::move(a), b); }
Tensor mul_inplace(Tensor a, const Tensor& b) { return binary_inplace(AGRS_MUL, std::move(a), b); }

static Tensor scalar_new(int op, const Tensor& a, float s) {
  auto out = std::make_shared<Storage>(a.dev(), a.shape());
  out->transposed = a.data->transposed;
  a.dev()->check(agrs_scalar_f32(a.dev()->ctx(), op, out->buf->ptr, a.data->rptr(), s, size_t(a.len())));
  return Tensor(out);
}
Tensor mul_scalar(const Tensor& a, float s) { return scalar_new(AGRS_SMUL, a, s); }
Tensor div_scalar(const Tensor& a, float s) { return scalar_new(AGRS_SDIV, a, s); }
Tensor rsub_scalar(float s, const Tensor& a) { return scalar_new(AGRS_SRSUB, a, s); }
void mul_assign_scalar(const Tensor& a, float s) {
  float* p = a.data->wptr();
  a.dev()->check(agrs_scalar_f32(a.dev()->ctx(), AGRS_SMUL, p, p, s, size_t(a.len())));
}
void div_assign_scalar(const Tensor& a, float s) {
  float* p = a.data->wptr();
  a.dev()->check(agrs_scalar_f32(a.dev()->ctx(), AGRS_SDIV, p, p, s, size_t(a.len())));
}
Tensor neg(const Tensor& a) {
  auto out = std::make_shared<Storage>(a.dev(), a.shape());
  out->transposed = a.data->transposed;
  a.dev()->check(agrs_neg_f32(a.dev()->ctx(), out->buf->ptr, a.data->rptr(), size_t(a.len())));
  return Tensor(out);
}
bool equal(const Tensor& a, const Tensor& b) {
  if (a.shape() != b.shape()) return false;  // ndarray ==: shape mismatch is `false`
  auto A = contig(a.data), B = contig(b.data);
  int eq = 0;
  a.dev()->check(agrs_equal_f32(a.dev()->ctx(), A->rptr(), B->rptr(), size_t(A->len()), &eq));
  return eq != 0;
}

}  // namespace agrs
