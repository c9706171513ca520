#include "dp.hpp"

namespace agrs {

DataParallel::DataParallel(Device* dev, int world, int rank, const void* unique_id) : dev_(dev), world_(world), rank_(rank) {
  dev_->check(agrs_comm_create(dev_->ctx(), world, rank, unique_id, &comm_));
}

DataParallel::~DataParallel() {
  for (auto& c : cells_)
    if (c->hook == this) c->hook = nullptr;
  if (comm_) agrs_comm_destroy(comm_);
}

void DataParallel::shard_rows(int64_t rows, int world, int rank, int64_t* begin, int64_t* end) {
  const int64_t base = rows / world, rem = rows % world;
  *begin = rank * base + (rank < rem ? rank : rem);
  *end = *begin + base + (rank < rem ? 1 : 0);
}

void DataParallel::register_parameters(const std::vector<Tensor>& params, bool overlap) {
  overlap_ = overlap;
  for (auto& p : params) {
    if (!p.grad) throw Panic("DataParallel: parameter without a grad cell");
    if (dirty_.count(p.grad.get())) continue;
    cells_.push_back(p.grad);
    dirty_[p.grad.get()] = false;
    p.grad->hook = this;
  }
}

void DataParallel::reduce(GradCell& cell) {
  if (!cell.g) throw Panic("called `Option::unwrap()` on a `None` value (DataParallel: parameter has no grad)");
  Storage& g = *cell.get();
  if (g.transposed) throw Panic("DataParallel: transposed gradient view", AGRS_ERR_UNSUPPORTED);
  // in place on the gradient buffer (un-shared first if a hand-over still aliases it); ncclAvg = sum / world
  dev_->check(agrs_comm_allreduce_f32(comm_, g.wptr(), size_t(g.len()), 1));
}

void DataParallel::on_accumulated(GradCell& cell) {
  auto it = dirty_.find(&cell);
  if (it == dirty_.end()) return;
  // Averaging is linear and an averaged gradient is identical on every rank, so mean_r(avg + local_r) = avg + mean_r(local_r):
  // reducing after every accumulation (overlap) or once before the optimiser step gives the same result.
  it->second = true;
  if (!overlap_ || world_ == 1) return;
  reduce(cell);
  it->second = false;
}

void DataParallel::sync_grads() {
  if (world_ > 1) {
    for (auto& c : cells_) {
      bool& dirty = dirty_[c.get()];
      if (dirty) reduce(*c);
      dirty = false;
    }
    dev_->check(agrs_comm_join(comm_));
  } else {
    for (auto& kv : dirty_) kv.second = false;
  }
}

void DataParallel::mean_scalar(const Tensor& t) {
  if (world_ == 1) return;
  dev_->check(agrs_comm_allreduce_f32(comm_, t.data->wptr(), size_t(t.len()), 1));
  dev_->check(agrs_comm_join(comm_));
}

}  // namespace agrs
