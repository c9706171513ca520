"""Data-parallel parity on real GPUs (run under torchrun): N ranks train a small MLP on disjoint shards of one global batch with the
engine's DataParallel (NCCL all-reduce with overlap on and off, then the fused peer-memory optimiser step); rank 0 also runs the ORACLE on the whole global batch.  Checks: every
loss (mean of shard means) and every final weight within 1e-4 of the single-device oracle; weights bit-identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import frontend as F, workloads as W  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = F.Device(local)
    layers, width, rows, steps, lr = 3, 512, 256 * world, 4, 1e-4
    ws, bs = W.mlp_init(layers, width, seed=3)
    x, t = W.mlp_batch(rows, width, width, seed=4)          # the global batch, identical on every rank
    b, e = F.shard_rows(rows, world, rank)
    def all_gather_bytes(h):
        out = [None] * world
        dist.all_gather_object(out, h)
        return out

    for overlap, fused, fo in ((True, False, True), (False, False, True), (True, True, True), (True, True, False)):
        ids = [F.comm_unique_id() if rank == 0 else None]   # an NCCL id is good for ONE communicator
        dist.broadcast_object_list(ids, src=0)
        model = W.build_mlp(dev, ws, bs, "sigmoid")
        dp = F.DataParallel(dev, world, rank, ids[0], model.parameters(), overlap=overlap)
        if fused:
            assert dp.enable_fused(all_gather_bytes), dp.fused_error   # parameters move into the peer-mapped arena; one fused kernel per step
        tr = W.Trainer(model, lr, dp=dp, fused_overlap=fo)
        X, T = F.Tensor.from_numpy(dev, x[b:e]), F.Tensor.from_numpy(dev, t[b:e])
        T.requires_grad = False
        losses = []
        for _ in range(steps):
            loss = tr.step(X, T)
            dp.mean_scalar(loss)
            losses.append(loss.item())
        params = [p.data() for p in model.parameters()]
        for p in params:  # bit-identical replicas
            tt = torch.from_numpy(p.copy()).cuda()
            lo, hi = tt.clone(), tt.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "replicas diverged"
        if rank == 0:
            from oracle import oracle_np as O
            want, olayers = O.mlp_train(x, t, ws, bs, "sigmoid", lr, steps, np.float32)
            el = max(abs(a - c) / abs(c) for a, c in zip(losses, want))
            ew = max(O.rel_err(p, q.data) for p, q in zip(params, [q for l in olayers for q in l.parameters()]))
            print(f"world={world} overlap={overlap} fused={fused} per_layer={fo}: losses {losses} oracle {want}; max rel err loss {el:.2e} weights {ew:.2e}")
            assert el < 1e-4 and ew < 1e-4
        dp.close()
        dist.barrier()
    if rank == 0:
        print("DP PARITY OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
