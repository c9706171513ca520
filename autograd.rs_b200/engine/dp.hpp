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

  // rows [begin, end) of a global batch owned by `rank` (equal shards; the remainder goes to the first ranks)
  static void shard_rows(int64_t rows, int world, int rank, int64_t* begin, int64_t* end);

 private:
  void reduce(GradCell& cell);
  Device* dev_;
  agrs_comm* comm_ = nullptr;
  int world_, rank_;
  bool overlap_ = false;
  std::vector<std::shared_ptr<GradCell>> cells_;
  std::unordered_map<GradCell*, bool> dirty_;  // holds local contributions that have not been averaged yet
};

}  // namespace agrs
