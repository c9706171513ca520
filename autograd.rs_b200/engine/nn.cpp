#include "nn.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace agrs {
std::pair<int64_t, int64_t> calculate_fan_in_and_fan_out(const Shape& s);  // tensor.cpp (init.rs:30-45)
namespace nn {

Tensor Layer::forward_default(std::vector<Tensor> xs) const {
  Tensor res = xs.at(0).clone();
  const auto ops = get_op_subgraph();
  for (size_t i = 0; i < ops.size(); ++i) {
    // a Linear / Conv2D operator feeding an activation inside one layer's subgraph goes out as one call (forward_pair decides)
    if (i + 1 < ops.size() && (std::holds_alternative<agrs::Linear>(ops[i]) || std::holds_alternative<agrs::Conv2D>(ops[i])) &&
        (std::holds_alternative<agrs::ReLU>(ops[i + 1]) || std::holds_alternative<agrs::Sigmoid>(ops[i + 1]))) {
      res = agrs::forward_pair(ops[i], ops[i + 1], xs);
      ++i;
    } else {
      res = agrs::forward(ops[i], xs);
    }
    xs = {res.clone()};
  }
  return res;
}

Linear::Linear(Device* d, int64_t in_features, int64_t out_features, bool /*bias: optional bias is a TODO in the reference, layers.rs:79*/) {
  Shape w_shape{in_features, out_features};
  float bound = 1.0f / std::sqrt(float(calculate_fan_in_and_fan_out(w_shape).first));  // fan_in = dim 1 = out_features (layers.rs:82-84)
  w_ = Tensor::kaiming_uniform(d, w_shape);
  bias_ = Tensor::uniform(d, {1, out_features}, -bound, bound);
}

Tensor Linear::forward(std::vector<Tensor> xs) const {
  if (xs.size() != 1) throw Panic("assertion failed: xs.len() == 1");
  for (auto& p : parameters()) xs.push_back(p);
  return forward_default(std::move(xs));
}

Conv2D::Conv2D(Device* d, int64_t in_channels, int64_t out_channels, int64_t k_size, int64_t stride, int64_t pad) {
  Shape w_shape{k_size, k_size, out_channels};
  float bound = 1.0f / std::sqrt(float(calculate_fan_in_and_fan_out(w_shape).first));  // layers.rs:155-157
  kernels_ = Tensor::kaiming_uniform(d, w_shape);
  bias_ = Tensor::uniform(d, {out_channels}, -bound, bound);
  op_ = agrs::Conv2D{in_channels, out_channels, k_size, stride, pad};
}

Tensor Conv2D::forward(std::vector<Tensor> xs) const {
  if (xs.size() != 1) throw Panic("assertion failed: xs.len() == 1");
  for (auto& p : parameters()) xs.push_back(p);
  return forward_default(std::move(xs));
}

std::vector<Tensor> NN::parameters() const {
  std::vector<Tensor> ps;
  for (auto& l : layers())
    for (auto& p : l->parameters()) ps.push_back(p);
  return ps;
}

void zero_grad(const std::vector<Tensor>& params) {
  for (auto& p : params)
    if (p.has_grad()) p.zero_grad();
}
void NN::zero_grad() const { nn::zero_grad(parameters()); }

void sgd_step(const std::vector<Tensor>& params, float lr) {
  // parameters whose gradient has their own length (every Linear / Conv2D parameter after a backward pass) are updated by ONE
  // launch (agrs_sgd_step_multi_f32); a gradient that broadcasts onto its parameter keeps the single-tensor call
  std::vector<float*> ws;
  std::vector<const float*> gs;
  std::vector<size_t> ns;
  std::vector<uint32_t*> mxs;
  Device* dev = nullptr;
  for (auto& p : params) {
    if (!p.has_grad()) throw Panic("called `Option::unwrap()` on a `None` value");  // w_grad.as_ref().unwrap(), optim.rs:30
    Storage& w = *p.data;
    const Storage& g = *p.grad->get();
    if (w.transposed || g.transposed) throw Panic("SGD on a transposed parameter view is not supported", AGRS_ERR_UNSUPPORTED);
    Bcast k = broadcast_kind(w.shape, g.shape);
    if (k == Bcast::Leading) throw Panic("SGD: gradient " + shape_str(g.shape) + " broadcasts along rows of " + shape_str(w.shape), AGRS_ERR_UNSUPPORTED);
    // tmp = grad * lr; w -= tmp fused in one pass with the same two roundings
    float* wp = w.wptr();
    if (w.len() == g.len() && (dev == nullptr || dev == w.dev())) {
      dev = w.dev();
      uint32_t* mx = nullptr;
      if (w.buf->n >= (int64_t(1) << 16) && dev->reports_magnitudes()) {  // the updated weights are the next step's GEMM operands
        mx = w.buf->mx_slot();
        w.buf->mx_valid = true;
      }
      ws.push_back(wp);
      gs.push_back(g.rptr());
      ns.push_back(size_t(w.len()));
      mxs.push_back(mx);
      continue;
    }
    w.dev()->check(agrs_sgd_step_f32(w.dev()->ctx(), wp, g.rptr(), lr, size_t(w.len()), size_t(g.len())));
  }
  if (!ws.empty()) dev->check(agrs_sgd_step_multi_f32(dev->ctx(), int(ws.size()), ws.data(), gs.data(), ns.data(), mxs.data(), lr));
}
void SGD::step() { sgd_step(parameters_, lr_); }

namespace {
struct File {
  FILE* f;
  File(const std::string& path, const char* mode) : f(fopen(path.c_str(), mode)) {
    if (!f) throw Panic("cannot open " + path);
  }
  ~File() { fclose(f); }
  void put(const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) throw Panic("short write while saving parameters");
  }
  void get(void* p, size_t n) {
    if (n && fread(p, 1, n, f) != n) throw Panic("parameter file is truncated");
  }
};
const char kMagic[8] = {'A', 'G', 'R', 'S', 'W', 'T', 'S', '1'};
}  // namespace

void save_parameters(const std::vector<Tensor>& params, const std::string& path) {
  File out(path, "wb");
  out.put(kMagic, 8);
  const uint32_t count = uint32_t(params.size());
  out.put(&count, 4);
  for (auto& p : params) {
    const uint32_t nd = uint32_t(p.ndim());
    out.put(&nd, 4);
    for (auto d : p.shape()) {
      const uint64_t v = uint64_t(d);
      out.put(&v, 8);
    }
    const std::vector<float> h = p.to_vec();  // logical order, whatever the storage's major
    out.put(h.data(), h.size() * sizeof(float));
  }
}

void load_parameters(const std::vector<Tensor>& params, const std::string& path) {
  File in(path, "rb");
  char magic[8];
  in.get(magic, 8);
  if (memcmp(magic, kMagic, 8) != 0) throw Panic(path + " is not an AGRSWTS1 parameter file");
  uint32_t count = 0;
  in.get(&count, 4);
  if (count != params.size())
    throw Panic("parameter file holds " + std::to_string(count) + " tensors, the model has " + std::to_string(params.size()), AGRS_ERR_SHAPE);
  for (auto& p : params) {
    uint32_t nd = 0;
    in.get(&nd, 4);
    Shape s(nd);
    for (auto& d : s) {
      uint64_t v = 0;
      in.get(&v, 8);
      d = int64_t(v);
    }
    if (s != p.shape()) throw Panic("parameter file has shape " + shape_str(s) + " where the model has " + shape_str(p.shape()), AGRS_ERR_SHAPE);
    if (p.data->transposed) throw Panic("load_parameters into a transposed view is not supported", AGRS_ERR_UNSUPPORTED);
    std::vector<float> h(size_t(p.len()));
    in.get(h.data(), h.size() * sizeof(float));
    if (!h.empty()) {
      p.dev()->check(agrs_upload(p.dev()->ctx(), p.data->wptr(), h.data(), h.size() * sizeof(float)));
      p.dev()->sync();  // `h` is pageable and dies here
    }
  }
}

}  // namespace nn
}  // namespace agrs
