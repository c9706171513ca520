// Host engine, nn layer: C++ mirror of the reference's src/nn
//   trait Layer + Linear/Conv2D/ReLU/Sigmoid/MeanSquaredError/Identity   src/nn/layers.rs:16-186
//   trait NN                                                            src/nn/model.rs:7-23
//   trait Optimizer, SGD                                                src/nn/optim.rs:4-35
#pragma once
#include <memory>
#include <vector>

#include "autograd.hpp"

namespace agrs {
namespace nn {

// layers.rs:16-44
class Layer {
 public:
  virtual ~Layer() = default;
  virtual std::vector<Operators> get_op_subgraph() const = 0;
  virtual std::vector<Tensor> parameters() const { return {}; }
  virtual Tensor forward(std::vector<Tensor> xs) const { return forward_default(std::move(xs)); }
  Tensor forward_default(std::vector<Tensor> xs) const;  // sequentially feed each operator's output into the next (:24-43)
};

struct Identity : Layer {
  std::vector<Operators> get_op_subgraph() const override { return {agrs::Identity{}}; }
};
struct ReLU : Layer {
  std::vector<Operators> get_op_subgraph() const override { return {agrs::ReLU{}}; }
};
struct Sigmoid : Layer {
  std::vector<Operators> get_op_subgraph() const override { return {agrs::Sigmoid{}}; }
};
struct MeanSquaredError : Layer {
  std::vector<Operators> get_op_subgraph() const override { return {agrs::MeanSquaredError{}}; }
};
struct Flatten : Layer {  // gap G1
  std::vector<Operators> get_op_subgraph() const override { return {agrs::Flatten{}}; }
};

// layers.rs:66-105: w [in, out] kaiming_uniform, bias [1, out] uniform(+-1/sqrt(fan_in(w)))
class Linear : public Layer {
 public:
  Linear(Device* d, int64_t in_features, int64_t out_features, bool bias = true);
  Linear(Tensor w, Tensor bias) : w_(std::move(w)), bias_(std::move(bias)) {}  // injected weights (parity harness)
  std::vector<Operators> get_op_subgraph() const override { return {agrs::Linear{}}; }
  std::vector<Tensor> parameters() const override { return {w_.clone(), bias_.clone()}; }  // clones share storage
  Tensor forward(std::vector<Tensor> xs) const override;

 private:
  Tensor w_, bias_;
};

// layers.rs:143-186: kernels [k, k, Co], bias [Co]
class Conv2D : public Layer {
 public:
  Conv2D(Device* d, int64_t in_channels, int64_t out_channels, int64_t k_size, int64_t stride = 1, int64_t pad = 0);
  Conv2D(Tensor kernels, Tensor bias, agrs::Conv2D op) : kernels_(std::move(kernels)), bias_(std::move(bias)), op_(op) {}
  std::vector<Operators> get_op_subgraph() const override { return {op_}; }
  std::vector<Tensor> parameters() const override { return {kernels_.clone(), bias_.clone()}; }
  Tensor forward(std::vector<Tensor> xs) const override;

 private:
  Tensor kernels_, bias_;
  agrs::Conv2D op_;
};

// model.rs:7-23: the user defines layers() and forward()
class NN {
 public:
  virtual ~NN() = default;
  virtual std::vector<std::shared_ptr<Layer>> layers() const = 0;
  virtual Tensor forward(std::vector<Tensor> xs) const = 0;
  std::vector<Tensor> parameters() const;
  void zero_grad() const;  // only parameters whose grad is Some (lazy init is kept, :14-21)
};
void zero_grad(const std::vector<Tensor>& params);

// optim.rs:4-35
class Optimizer {
 public:
  virtual ~Optimizer() = default;
  virtual void step() = 0;
  void zero_grad(const NN& model) const { model.zero_grad(); }
};
class SGD : public Optimizer {
 public:
  SGD(std::vector<Tensor>& parameters, float lr) : parameters_(parameters), lr_(lr) {}
  void step() override;  // w -= grad * lr with ndarray broadcasting of grad onto w (:25-34); panics if a grad is None

 private:
  std::vector<Tensor>& parameters_;
  float lr_;
};
void sgd_step(const std::vector<Tensor>& params, float lr);

// Model serialisation (the reference lists it as a TODO: README.md:71).  One self-describing little-endian file:
//   "AGRSWTS1" | u32 count | per tensor { u32 ndim | u64 dim[ndim] | f32 data[prod(dim)] in logical (row-major) order }
// save reads the parameters back (blocks); load writes INTO the existing parameter storages (every clone follows, parameters
// living in a data-parallel arena stay there) and panics on a count / shape mismatch.
void save_parameters(const std::vector<Tensor>& params, const std::string& path);
void load_parameters(const std::vector<Tensor>& params, const std::string& path);

}  // namespace nn
}  // namespace agrs
