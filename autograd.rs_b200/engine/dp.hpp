// Host engine, data-parallel layer (new: the reference has no multi-device concept).
// One process per GPU; parameters replicated, minibatch rows sharded, parameter gradients averaged over ranks with
// NCCL on a communication stream (agrs_comm_*, agrs.h).  The literal MSE gradient divides by the LOCAL batch
// (operators.rs:247), so the mean over ranks of the local gradients is exactly the literal gradient of the global batch.
#pragma once
#include <unordered_map>
#include <vector>

#include "tensor.hpp"

namespace agrs {

class DataParallel : public GradHook {
 public:
  DataParallel(Device* dev, int world, int rank, const void* unique_id);
  ~DataParallel() override;
  DataParallel(const DataParallel&) = delete;

  int world() const { return world_; }
  int rank() const { return rank_; }

  // Parameters whose gradients are exchanged.  overlap: enqueue each all-reduce as soon as backward has accumulated
  // that gradient (backward runs last layer -> first, autograd.rs:81-85), so it overlaps the rest of the walk.
  void register_parameters(const std::vector<Tensor>& params, bool overlap);
  // all-reduce what is still local, then make the compute stream wait for the communication stream
  void sync_grads();
  void mean_scalar(const Tensor& t);  // loss value: mean of the shard means (equal shards)

  void on_accumulated(GradCell& cell) override;

  // ---- fused exchange + optimiser step over NVLink peer memory (csrc/fused_dp.cu): replaces sync_grads() + SGD::step ----
  // fused_prepare  lays the registered parameters out in an arena (agrs_peer_alloc), moves their values there and re-points
  //                the parameter storages at it (every clone of the parameter sees the move), exports the arena (64 bytes);
  // fused_connect  maps the arenas of all ranks (handles gathered by the launcher, world x 64 bytes, rank order);
  // during backward every parameter gradient is pushed into this rank's slot at the ranks that own it: by the epilogue of the
  //                dW GEMM itself (Linear::backward asks for agrs_gemm_f32_scatter) or by on_accumulated (scatter kernel);
  // then EITHER, per step,
  //   fused_begin(lr) before backward: from now on, the moment backward has accumulated (and pushed) every gradient of a BUCKET
  //                (one layer's W and b, contiguous in the arena) the exchange of that bucket starts on a communication stream --
  //                local sum of the slots, w -= lr * mean(g) on the owned granules, all-gather by the copy engines -- while the
  //                backward of the earlier layers keeps the compute stream busy (agrs_fused_bucket_ready);
  //   fused_finish() after backward: the compute stream waits for what is still in flight (the first layer's exchange);
  // OR fused_sgd_step(lr) after backward: one kernel (barrier, slot sum, SGD, all-gather with P2P stores, barrier).
  // p.grad() keeps this rank's LOCAL gradient either way: the mean is never materialised.
  void fused_prepare(void* handle64);
  void fused_connect(const void* handles);
  void fused_disable();  // a peer could not connect: back to the NCCL exchange (parameters stay in the arena, nothing is pushed)
  void fused_begin(float lr);
  void fused_finish();
  void fused_sgd_step(float lr);
  bool fused() const { return fused_ != nullptr; }
  double fused_exposed_ms(uint64_t* steps);  // profiling: time the compute stream waited for the exchange in fused_finish

  // rows [begin, end) of a global batch owned by `rank` (equal shards; the remainder goes to the first ranks)
  static void shard_rows(int64_t rows, int world, int rank, int64_t* begin, int64_t* end);

 private:
  void reduce(GradCell& cell);
  Device* dev_;
  agrs_comm* comm_ = nullptr;
  int world_, rank_;
  bool overlap_ = false;
  std::vector<std::shared_ptr<GradCell>> cells_;
  // fused mode
  struct Slot { int64_t off = 0, n = 0; int bucket = 0; };
  struct Bucket { int64_t off = 0, n = 0; int cells = 0, ready = 0; };
  std::vector<Bucket> buckets_;
  bool armed_ = false;       // between fused_begin and fused_finish
  int granule_log2_ = 14;
  std::vector<Tensor> params_;                    // registration order (one per cell)
  std::unordered_map<GradCell*, Slot> slots_;     // float offset of the parameter inside the weight / gradient regions
  float* arena_ = nullptr;
  std::vector<void*> peer_maps_;                  // arenas of all ranks as mapped here ([rank] = arena_)
  size_t w_off_ = 0, w_alt_ = 0, g_off_ = 0, region_ = 0;  // two weight regions (double-buffered by the bucket protocol), gradient slots
  void repoint_parameters();   // after a swap of the weight regions: every parameter storage follows
  agrs_fused* fused_ = nullptr;
  std::unordered_map<GradCell*, bool> dirty_;  // holds local contributions that have not been averaged yet
};

}  // namespace agrs
