#include "dp.hpp"

namespace agrs {

DataParallel::DataParallel(Device* dev, int world, int rank, const void* unique_id) : dev_(dev), world_(world), rank_(rank) {
  dev_->check(agrs_comm_create(dev_->ctx(), world, rank, unique_id, &comm_));
}

DataParallel::~DataParallel() {
  for (auto& c : cells_) {
    if (c->hook == this) c->hook = nullptr;
    c->scatter = nullptr;
    c->scattered = false;
  }
  if (fused_) agrs_fused_destroy(fused_);
  for (size_t p = 0; p < peer_maps_.size(); ++p)
    if (int(p) != rank_ && peer_maps_[p]) agrs_peer_close(dev_->ctx(), peer_maps_[p]);
  if (arena_) {
    // parameters move back to ordinary storage before the arena goes away
    for (auto& t : params_) {
      if (t.data->buf->owned) continue;
      auto nb = std::make_shared<Buffer>(dev_, t.data->buf->n);
      agrs_copy(dev_->ctx(), nb->ptr, t.data->buf->ptr, size_t(nb->n) * sizeof(float));
      t.data->buf = nb;
    }
    agrs_peer_free(dev_->ctx(), arena_);
  }
  if (comm_) agrs_comm_destroy(comm_);
}

void DataParallel::fused_prepare(void* handle64) {
  if (arena_) throw Panic("DataParallel::fused_prepare called twice");
  if (params_.empty()) throw Panic("DataParallel::fused_prepare: register the parameters first");
  // Buckets: a large parameter (a weight matrix) opens one, the small ones that follow (its bias) join it -- the gradients of
  // one layer become ready together during backward.  Every bucket is a granule*world aligned range of the flat parameter space.
  granule_log2_ = 14;  // 64 KB granules
  const int64_t G = int64_t(1) << granule_log2_, GW = G * world_;
  const int max_buckets = agrs_fused_max_buckets();
  int64_t off = 0;
  for (auto& t : params_) {
    if (t.data->transposed) throw Panic("DataParallel(fused): transposed parameter view", AGRS_ERR_UNSUPPORTED);
    const bool opens = buckets_.empty() || (t.len() >= (int64_t(1) << 16) && int(buckets_.size()) < max_buckets);
    if (opens) {
      off = (off + GW - 1) / GW * GW;
      buckets_.push_back(Bucket{off, 0, 0, 0});
    }
    Bucket& bk = buckets_.back();
    slots_[t.grad.get()] = Slot{off, t.len(), int(buckets_.size()) - 1};
    bk.cells += 1;
    off += (t.len() + 63) / 64 * 64;  // 256-byte aligned slots: TMA-mappable, float4-friendly
    bk.n = off - bk.off;
  }
  for (auto& bk : buckets_) bk.n = (bk.n + GW - 1) / GW * GW;
  region_ = size_t((off + GW - 1) / GW * GW);
  w_off_ = agrs_fused_flag_floats();
  w_alt_ = w_off_ + region_;   // second weight region: the exchange writes the new weights where the running step does not read
  g_off_ = w_alt_ + region_;
  void* a = nullptr;
  dev_->check(agrs_peer_alloc(dev_->ctx(), (g_off_ + region_) * sizeof(float), &a));
  arena_ = static_cast<float*>(a);
  for (auto& t : params_) {
    const Slot& sl = slots_[t.grad.get()];
    float* dst = arena_ + w_off_ + sl.off;
    dev_->check(agrs_copy(dev_->ctx(), dst, t.data->rptr(), size_t(sl.n) * sizeof(float)));
    dev_->check(agrs_copy(dev_->ctx(), arena_ + w_alt_ + sl.off, t.data->rptr(), size_t(sl.n) * sizeof(float)));
    t.data->buf = std::make_shared<Buffer>(dev_, dst, sl.n);  // the shared storage cell now points into the arena
    t.grad->scatter_off = size_t(sl.off);
  }
  dev_->sync();
  dev_->check(agrs_peer_export(dev_->ctx(), arena_, handle64));
}

void DataParallel::fused_connect(const void* handles) {
  if (!arena_) throw Panic("DataParallel::fused_connect before fused_prepare");
  if (fused_) throw Panic("DataParallel::fused_connect called twice");
  peer_maps_.assign(size_t(world_), nullptr);
  for (int p = 0; p < world_; ++p) {
    if (p == rank_) {
      peer_maps_[size_t(p)] = arena_;
      continue;
    }
    dev_->check(agrs_peer_open(dev_->ctx(), static_cast<const char*>(handles) + size_t(p) * AGRS_PEER_HANDLE_BYTES, &peer_maps_[size_t(p)]));
  }
  dev_->check(agrs_fused_create(dev_->ctx(), world_, rank_, peer_maps_.data(), w_off_, w_alt_, g_off_, region_, granule_log2_, &fused_));
  for (auto& c : cells_) c->scatter = fused_;  // from now on the dW GEMMs push their tiles to the owners themselves
}

void DataParallel::fused_disable() {
  for (auto& c : cells_) {
    c->scatter = nullptr;
    c->scattered = false;
  }
  if (fused_) agrs_fused_destroy(fused_);
  fused_ = nullptr;
  armed_ = false;
}

void DataParallel::fused_begin(float lr) {
  if (!fused_) throw Panic("DataParallel::fused_begin before fused_connect");
  if (armed_) throw Panic("DataParallel::fused_begin: the previous step was not finished");
  for (auto& kv : dirty_)
    if (kv.second) throw Panic("DataParallel::fused_begin: a parameter already holds a gradient of this step (call it before backward)");
  for (auto& bk : buckets_) bk.ready = 0;
  dev_->check(agrs_fused_begin_step(fused_, lr));
  armed_ = true;
}

void DataParallel::fused_finish() {
  if (!armed_) throw Panic("DataParallel::fused_finish without fused_begin");
  for (auto& c : cells_) {
    bool& dirty = dirty_[c.get()];
    if (!dirty) throw Panic("called `Option::unwrap()` on a `None` value (DataParallel fused step: a parameter received no gradient)");
    dirty = false;
  }
  dev_->check(agrs_fused_finish_step(fused_));
  armed_ = false;
  repoint_parameters();  // the new weights live in the other region: fresh buffers (no magnitude yet) for every parameter storage
}

double DataParallel::fused_exposed_ms(uint64_t* steps) {
  double ms = 0.0;
  if (fused_) dev_->check(agrs_fused_exposed_ms(fused_, &ms, steps));
  return ms;
}

void DataParallel::repoint_parameters() {
  size_t w_now = 0;
  dev_->check(agrs_fused_weights_offset(fused_, &w_now));
  for (auto& t : params_) {
    const Slot& sl = slots_[t.grad.get()];
    float* p = arena_ + w_now + sl.off;
    if (t.data->buf->ptr != p) t.data->buf = std::make_shared<Buffer>(dev_, p, sl.n);
    else t.data->buf->mx_valid = false;
  }
}

void DataParallel::fused_sgd_step(float lr) {
  if (!fused_) throw Panic("DataParallel::fused_sgd_step before fused_connect");
  if (armed_) throw Panic("DataParallel::fused_sgd_step inside fused_begin / fused_finish");
  for (auto& c : cells_) {
    bool& dirty = dirty_[c.get()];
    if (!dirty) throw Panic("called `Option::unwrap()` on a `None` value (DataParallel fused step: a parameter received no gradient)");
    dirty = false;
  }
  dev_->check(agrs_fused_sgd_step(fused_, lr));
  for (auto& t : params_) t.data->buf->mx_valid = false;  // the kernel rewrote every parameter (its own slice and the peers' broadcasts)
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
    params_.push_back(p);
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
  if (!fused_) it->second = true;
  if (fused_) {  // the gradient joins the arena; the exchange itself is the fused step's job
    const Slot& sl = slots_[&cell];
    const Storage& g = *cell.get();
    if (g.transposed || g.len() != sl.n) throw Panic("DataParallel(fused): gradient does not match its parameter", AGRS_ERR_SHAPE);
    if (armed_ && it->second) throw Panic("DataParallel(fused, overlapped): a parameter received a second gradient in one step; use fused_sgd_step for that", AGRS_ERR_UNSUPPORTED);
    it->second = true;
    if (cell.scattered)  // the GEMM that produced it already pushed it, tile by tile
      cell.scattered = false;
    else                 // bias sums, conv filters, split-K results, second accumulations: push it now
      dev_->check(agrs_fused_scatter_f32(fused_, g.rptr(), size_t(sl.off), size_t(sl.n)));
    if (armed_) {        // the last gradient of a bucket starts that bucket's exchange, under the rest of backward
      Bucket& bk = buckets_[size_t(sl.bucket)];
      if (++bk.ready == bk.cells) dev_->check(agrs_fused_bucket_ready(fused_, sl.bucket, size_t(bk.off), size_t(bk.n)));
    }
    return;
  }
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
