// Host engine, operator + autograd layer: C++ mirror of the reference's
//   trait Operator / op structs / enum Operators   src/operators/operators.rs:18-479
//   Node, backward_algo                            src/autograd/autograd.rs:12-86
// The graph walk stays on the host (north star); every forward/backward is a handful of asynchronous
// launches through agrs.h -- nothing here ever reads a device value back.
#pragma once
#include <memory>
#include <string>
#include <variant>
#include <vector>

#include "tensor.hpp"

namespace agrs {

struct Identity {};
struct ReLU {};
struct Sigmoid {};
struct Mean {};
struct MeanSquaredError {
  // Data-parallel training shards the minibatch: the literal rule divides by the LOCAL batch
  // (operators.rs:247), so replicas average their gradients after the all-reduce.
};
struct Linear {};
struct MatMul {};
struct Flatten {};  // not in the reference (gap G1): zero-cost [B, ...] -> [B, prod(...)] view, backward views back
struct Conv2D {     // operators.rs:66-73
  int64_t in_channels = 1, out_channels = 1, k_size = 1, stride = 1, pad = 0;
  std::pair<int64_t, int64_t> get_output_shape(int64_t h, int64_t w) const;  // :316-320
  // im2col / col2im wrappers with the reference's shape checks (operators.rs:324-342, functional_ndarray.rs:100-103,134-137)
  static Tensor im2col(const Tensor& im, int64_t k_size, int64_t stride, int64_t pad);
  static Tensor im2col_backward(const Tensor& col, int64_t h_out, int64_t w_out, int64_t stride, int64_t k_size);
};

using Operators = std::variant<ReLU, Sigmoid, Linear, MatMul, MeanSquaredError, Mean, Identity, Conv2D, Flatten>;  // operators.rs:455-464
std::string operator_name(const Operators& op);                                                                    // Into<String>, :466-479

// autograd.rs:12-47
struct Node {
  Operators op;
  std::vector<Tensor> variables;
  std::vector<std::shared_ptr<Node>> parents;  // null = input was not produced by an operator
  bool is_root_node() const;
};

// trait Operator (operators.rs:18-50): forward builds the eager graph node, backward returns one gradient per input
// (MeanSquaredError returns a single gradient: the target gets none, autograd.rs:73 zips to the shorter list).
Tensor forward(const Operators& op, std::vector<Tensor> xs);
std::vector<Tensor> backward(const Operators& op, const std::vector<Tensor>& xs, const Tensor& grad);
// forward(second, {forward(first, xs)}) -- same nodes, same bits -- as one call where the pair is Linear or (single-channel,
// implicit-GEMM) Conv2D + ReLU / Sigmoid and the device's fuse_forward_activation is on (the kernel that produces the
// pre-activation writes the activation next to it where that pays)
Tensor forward_pair(const Operators& first, const Operators& second, std::vector<Tensor> xs);
void attach_to_eager_graph(std::vector<Tensor>& inputs, Tensor& op_output, const Operators& op);  // operators.rs:27-49

void backward_algo(const std::shared_ptr<Node>& node, const Tensor& prev_grad);  // autograd.rs:49-86

}  // namespace agrs
