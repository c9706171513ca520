// Host engine, tensor layer: the C++ mirror of the reference's src/tensor (Rust is not available
// in the build image).  Names and semantics follow the reference one to one:
//   Tensor<f32>            src/tensor/tensor.rs:27-34   (shared data cell, shared grad cell, graph pointer)
//   StorageType::CudaData  src/tensor/storage.rs:5-45   (here: Storage = device buffer + host-side shape metadata)
//   arithmetic rules       src/tensor/tensor_arithmetic.rs (which forms allocate, which are in place)
// This file includes NO CUDA header: the device is reached only through the C ABI in agrs.h,
// exactly as the Rust crate would through its -sys binding.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "agrs.h"

namespace agrs {

using Shape = std::vector<int64_t>;

// The reference panics (panic!/unwrap); the C++ mirror throws Panic with the same message.
struct Panic : std::runtime_error {
  int code;
  explicit Panic(const std::string& m, int c = AGRS_ERR_INVALID) : std::runtime_error(m), code(c) {}
};

int64_t numel(const Shape& s);
std::string shape_str(const Shape& s);

// One device context (agrs_ctx) + host RNG for the initialisers.  Single host thread, like Rc<RefCell>.
// When owned by a shared_ptr (the C API does so), every Buffer keeps the Device alive, so tensor handles may outlive the
// device handle: the context is destroyed with the last buffer.
class Device : public std::enable_shared_from_this<Device> {
 public:
  explicit Device(int ordinal, void* stream = nullptr, bool on_stream = false);
  ~Device();
  Device(const Device&) = delete;
  agrs_ctx* ctx() const { return ctx_; }
  void check(int rc) const;  // non-zero status -> Panic(agrs_last_error)
  void sync() const { check(agrs_sync(ctx_)); }
  void seed(uint64_t s) { rng_state_ = s ? s : 0x9E3779B97F4A7C15ull; }
  // fp16 split mode of the GEMM: would a product of this shape use operand magnitudes / is this output worth reporting on
  bool gemm_wants_magnitudes(int64_t M, int64_t N, int64_t K) const { return agrs_gemm_wants_absmax(ctx_, M, N, K) != 0; }
  bool reports_magnitudes() const { return agrs_gemm_wants_absmax(ctx_, 1 << 20, 1 << 20, 1 << 20) != 0; }
  float next_uniform(float lo, float hi);  // xorshift64*: the reference's init is unseeded (init.rs:27,48); only bounds are contractual
  // MatMul::backward flavour (reference defect D3, operators.rs:183-184): literal by default
  bool matmul_backward_literal = true;
  // Fusions across the Operator boundary; same results bit for bit, so on by default; switchable for A/B tests
  bool fuse_backward_colsum = true;
  bool fuse_forward_activation = true;  // Linear + ReLU / Sigmoid: the activation written by the GEMM's epilogue (forward_pair)
  // Conv2D with one input channel: implicit-GEMM kernels, no col matrix.  Off = always the literal lowering
  // (im2col + GEMM, col cached as the node's 4th variable, operators.rs:386-388).
  bool conv_implicit = true;

 private:
  agrs_ctx* ctx_ = nullptr;
  uint64_t rng_state_ = 0x9E3779B97F4A7C15ull;
};

// Device memory, freed stream-ordered when the last Storage referring to it goes away
// (CudaData's Drop, storage.rs:10-14).
struct Buffer {
  Device* dev;
  std::shared_ptr<Device> keep;  // null when the Device is not shared-owned (plain C++ use: the Device must outlive its tensors)
  float* ptr = nullptr;
  int64_t n = 0;
  bool owned = true;             // false: a window into memory owned elsewhere (a parameter living in the data-parallel arena)
  // Magnitude of the contents (bit pattern of the largest finite |x|), one device word, for the split GEMM's fp16 mode
  // (agrs_gemm_operand_absmax): reported by the kernel that wrote the buffer, or found by one pass the first time a GEMM asks.
  // Every write access (Storage::wptr) invalidates it.
  uint32_t* mx = nullptr;
  bool mx_valid = false;
  uint32_t* mx_slot();           // allocates the word on first use
  const uint32_t* magnitude();   // valid magnitude: the producer's report, else agrs_absmax_f32 now
  // Sum over rows of the [rows, cols] contents, left here by the backward kernel that wrote the buffer (agrs_*_bwd_colsum_f32)
  // for the Linear layer below, whose db it is.  Single use: Linear::backward takes it; any write access drops it.
  std::shared_ptr<struct Storage> colsum;
  int64_t colsum_cols = 0;
  // A host-to-device copy into this buffer that is still in flight on the copy stream (agrs_upload_prefetch_event): the first
  // reader or writer makes the compute stream wait for exactly this copy (Storage::rptr / wptr), not for every prefetch
  // issued so far -- a step's forward pass starts when X has landed while T is still on the bus.
  agrs_event* landing = nullptr;
  bool landing_pending = false;
  void await_landing();
  Buffer(Device* d, int64_t n_elems);
  Buffer(Device* d, float* external, int64_t n_elems);  // non-owning
  ~Buffer();
  Buffer(const Buffer&) = delete;
};

// StorageType::CudaData: a buffer plus the metadata the reference left as todo!() (storage.rs:49-55,81).
// `transposed` records an in-place t()/swap_axes on a 2-D array (tensor.rs:165-186): logical shape is
// `shape`, memory is the row-major [shape[1], shape[0]] array, i.e. the array is column-major.
struct Storage {
  std::shared_ptr<Buffer> buf;
  Shape shape;
  bool transposed = false;

  Storage(Device* d, const Shape& s);
  Storage(std::shared_ptr<Buffer> b, const Shape& s, bool t = false) : buf(std::move(b)), shape(s), transposed(t) {}
  Device* dev() const { return buf->dev; }
  int64_t len() const { return numel(shape); }
  size_t ndim() const { return shape.size(); }
  bool is_contiguous() const { return !transposed; }  // storage.rs:63-69 is_standard_layout
  const float* rptr() const {
    if (buf->landing_pending) buf->await_landing();
    return buf->ptr;
  }
  // pointer for in-place writes: un-shares the buffer first if a gradient hand-over aliases it
  float* wptr();
  std::shared_ptr<Storage> deep_copy() const;              // ArrayD::to_owned / CudaData clone (tensor.rs:148,160)
  std::shared_ptr<Storage> contiguous_copy() const;        // standard-layout copy (materialises a transposed view)
  std::vector<float> to_host() const;                      // logical (row-major) order; blocks
  // Call right before the ONE elementwise launch that fills this (fresh or just-wptr()'d) storage: that launch then also reports
  // the magnitude of what it writes (agrs_next_output_absmax).  No-op for small tensors and outside the GEMM's fp16 mode.
  void expect_magnitude_report();
};
using StoragePtr = std::shared_ptr<Storage>;

struct GradCell;
// Observer of gradient accumulation (data-parallel training hangs its all-reduce here): called by backward_algo
// right after a gradient was accumulated into the cell and the walk has dropped its own reference to it.
struct GradHook {
  virtual ~GradHook() = default;
  virtual void on_accumulated(GradCell& cell) = 0;
};

struct GradCell {  // Rc<RefCell<Option<StorageType<f32>>>>, tensor.rs:32
  StoragePtr g;
  bool zero_pending = false;   // g is logically all zeros but its buffer has not been cleared yet (lazy zero_grad)
  GradHook* hook = nullptr;    // not owned
  // Data-parallel exchange in fused mode (engine/dp.cpp, csrc/fused_dp.cu): where this parameter's gradient belongs in the flat
  // parameter space.  An operator that produces the gradient as a fresh tensor may use a GEMM whose epilogue pushes the result
  // to the owning ranks as it is produced (agrs_gemm_f32_scatter) and then sets `scattered`; otherwise the hook pushes it.
  agrs_fused* scatter = nullptr;
  size_t scatter_off = 0;
  bool scattered = false;  // the gradient taken over by the last accumulation is already at its owners
  bool next_accumulate_is_handover() const { return !g || zero_pending; }
  const StoragePtr& get();     // the gradient storage with any pending zero fill applied: what every reader must use
};

struct Node;

class Tensor {
 public:
  bool requires_grad = true;
  std::shared_ptr<Node> graph;     // tensor.rs:30
  StoragePtr data;                 // tensor.rs:31  (shared by clones)
  std::shared_ptr<GradCell> grad;  // tensor.rs:32  (shared by clones)

  Tensor() = default;
  explicit Tensor(StoragePtr s) : data(std::move(s)), grad(std::make_shared<GradCell>()) {}  // Tensor::from(storage), tensor.rs:50-61

  // ---- constructors (tensor.rs:37-92) ----
  static Tensor from(Device* d, const float* host, const Shape& shape);
  static Tensor from(Device* d, const std::vector<float>& host, const Shape& shape) { return from(d, host.data(), shape); }
  static Tensor zeros(Device* d, const Shape& shape);
  static Tensor ones(Device* d, const Shape& shape);
  static Tensor full(Device* d, const Shape& shape, float v);
  static Tensor uniform(Device* d, const Shape& shape, float low, float high);  // init.rs:47-49
  static Tensor kaiming_uniform(Device* d, const Shape& shape);                 // init.rs:19-28
  Tensor clone() const { return *this; }                                        // #[derive(Clone)]: shares data, grad, graph
  Tensor to_owned() const;                                                      // tensor.rs:93-105: graph=None, shares data+grad (Q1)

  // ---- metadata (tensor.rs:364-421) ----
  Device* dev() const { return data->dev(); }
  const Shape& shape() const { return data->shape; }
  int64_t shapei(size_t i) const { return data->shape.at(i); }
  size_t ndim() const { return data->ndim(); }
  int64_t len() const { return data->len(); }
  int64_t size() const { return len(); }
  bool is_contiguous() const { return data->is_contiguous(); }
  std::vector<float> to_vec() const { return data->to_host(); }  // data().into_raw_vec(): D2H, blocks
  bool has_grad() const { return grad && grad->g; }
  std::vector<float> grad_to_vec() const;
  float item() const;  // single-element read back

  // ---- shape ops: mutate the SHARED storage metadata, visible through every clone (tensor.rs:370-415) ----
  void reshape(const Shape& shape);
  Tensor unsqueeze(int ax) const;  // tensor.rs:395-409 (ax < 0 always appends, Q3)
  Tensor squeeze() const;          // tensor.rs:410-415
  const Tensor& t() const;  // tensor.rs:165-169: swaps axes IN PLACE on shared storage, returns self
  Tensor t_clone() const;   // tensor.rs:170-173
  void swap_axes(int a, int b) const;

  // ---- reductions / maps ----
  float sum() const;                        // tensor.rs:188-190 (reads back)
  Tensor sum_axis(int axis) const;          // tensor.rs:422-427
  Tensor mean() const;                      // tensor.rs:428-434, axis=None -> [1]
  Tensor mean_axis(int axis) const;
  Tensor powi(int exp) const;               // tensor.rs:221-231
  Tensor powi_inplace(int exp) const;       // tensor.rs:232-235
  void fill(float v) const;                 // tensor.rs:204-206
  Tensor dot(const Tensor& other) const;    // tensor.rs:301-320

  // ---- autograd plumbing ----
  void zero_grad() const;                                          // tensor.rs:135-140
  void accumulate_grad_from_grad_tensor(const Tensor& g) const;    // tensor.rs:154-163 (+= implemented: defect D2)
  void accumulate_grad(const Tensor& other) const;                 // tensor.rs:142-152
  void backward() const;                                           // tensor.rs:328-357
};

Tensor ones_like(const Tensor& t);   // tensor.rs:438-446
Tensor zeros_like(const Tensor& t);  // tensor.rs:448-451

// ---- arithmetic with the reference's allocation rules (tensor_arithmetic.rs) ----
// "new" forms (&A op &B, scalar forms, neg): fresh storage, grad=None, no graph.
Tensor add_new(const Tensor& a, const Tensor& b);   // &a + &b      :41-53
Tensor sub_new(const Tensor& a, const Tensor& b);   // &a - &b      :157-167
Tensor mul_new(const Tensor& a, const Tensor& b);   // &a * &b      :216-225
// in-place forms (A op &B, A op B, A op= &B): write into a's storage (visible through clones), return a
Tensor add_inplace(Tensor a, const Tensor& b);      // a + &b, a + b, a += &b   :56-112
Tensor sub_inplace(Tensor a, const Tensor& b);      // a - &b, a - b, a -= &b   :138-204
Tensor mul_inplace(Tensor a, const Tensor& b);      // a * &b, a * b (keeps a's storage object)  :227-262
Tensor mul_scalar(const Tensor& a, float s);        // &a * s, a * s  -> new   :265-283
Tensor div_scalar(const Tensor& a, float s);        // &a / s, a / s  -> new   :295-313
Tensor rsub_scalar(float s, const Tensor& a);       // s - &a         -> new   :326-338
void mul_assign_scalar(const Tensor& a, float s);   // a *= s                  :285-293
void div_assign_scalar(const Tensor& a, float s);   // a /= s                  :315-323
Tensor neg(const Tensor& a);                        // -a -> new (todo!() in the reference, Q4)  :114-136
bool equal(const Tensor& a, const Tensor& b);       // a == b: data only       :206-213

// storage-level helpers used by operators and the optimiser (storage_arithmetic.rs)
enum class Bcast { Same, Trailing, Leading };
// how `b` broadcasts onto `a` under ndarray rules restricted to what the device kernels map; throws Panic otherwise
Bcast broadcast_kind(const Shape& a, const Shape& b);
void storage_binary_into(int op, Storage& out, const Storage& a, const Storage& b);  // out may be a

}  // namespace agrs
