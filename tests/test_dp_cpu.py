"""CPU (gloo, world_size 2) tests of the data-parallel host logic: row sharding, the unique-id exchange the launcher
does, and the numerics claim the design rests on -- the mean over ranks of the LITERAL local gradients
(MSE backward divides by the local batch, operators.rs:247) equals the literal gradient of the global batch, so
replicas that average their gradients track the single-device oracle step for step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import frontend as F
from autograd_rs_b200 import workloads as W
from oracle import oracle_np as O


def test_shard_rows_partition():
    for rows in (0, 1, 7, 8, 8192, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [F.shard_rows(rows, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_rank_step(layers, x, t, act):
    X, T = O.OTensor(x), O.OTensor(t, requires_grad=False)
    h = X.clone()
    for i, l in enumerate(layers):
        h = l.forward(h)
        if i + 1 < len(layers):
            h = act().forward([h])
    loss = O.OpMSE().forward([h, T.clone()])
    loss.backward()
    return float(loss.data[0, 0])


def _worker(rank, world, port, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the launcher's id exchange (bench.py): rank 0 makes a 128-byte id, everyone receives the same bytes
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    ws, bs = W.mlp_init(3, 64, seed=5)
    x, t = W.mlp_batch(96, 64, 64, seed=6)             # the GLOBAL batch, identical on both ranks
    b, e = F.shard_rows(96, world, rank)
    layers = [O.OLinearLayer(w.copy(), bb.copy()) for w, bb in zip(ws, bs)]
    params = [p for l in layers for p in l.parameters()]
    losses = []
    for _ in range(steps):
        loss = _oracle_rank_step(layers, x[b:e], t[b:e], O.OpReLU)
        lt = torch.tensor([loss], dtype=torch.float64)
        dist.all_reduce(lt)
        losses.append(float(lt.item()) / world)        # mean of the shard means (DataParallel.mean_scalar)
        for p in params:                               # DataParallel.sync_grads: mean over ranks, in place
            g = torch.from_numpy(p.cell["grad"])
            dist.all_reduce(g)
            g /= world
        O.sgd(params, 1e-3)
        O.model_zero_grad(params)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array(losses), **{f"p{i}": p.data for i, p in enumerate(params)})
    dist.destroy_process_group()


def test_two_rank_data_parallel_tracks_single_device(tmp_path):
    world, steps = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), steps, str(tmp_path)), nprocs=world, join=True)
    ws, bs = W.mlp_init(3, 64, seed=5)
    x, t = W.mlp_batch(96, 64, 64, seed=6)
    want, layers = O.mlp_train(x, t, ws, bs, "relu", 1e-3, steps, np.float32)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    assert np.allclose(r0["losses"], want, rtol=1e-5)
    for i, p in enumerate([q for l in layers for q in l.parameters()]):
        assert np.array_equal(r0[f"p{i}"], r1[f"p{i}"])            # replicas stay bit-identical
        assert O.rel_err(r0[f"p{i}"], p.data) < 1e-5               # and track the single-device run


# ---- the fused exchange's ownership map (csrc/fused_dp.cu, gemm_tcgen05_2cta.cu), restated in NumPy -----------------------
GRANULE = 32  # floats (agrs_fused_create's granule_log2 = 5; the engine uses 2^14 = 64 KB granules: same maps, see the last test)


def _owner_and_index(e, world):
    """flat parameter offset e -> (owning rank, float index inside the owner's slice): k_scatter_grad / the GEMM epilogue"""
    g = e // GRANULE
    return g % world, (g // world) * GRANULE + e % GRANULE


def _flat_offset(i, world, rank):
    """float index i of rank's slice -> flat parameter offset: k_fused_dp_sgd"""
    return ((i // GRANULE) * world + rank) * GRANULE + i % GRANULE


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_fused_exchange_ownership_is_a_bijection(world):
    total = GRANULE * world * 7
    shard = total // world
    e = np.arange(total)
    owner, idx = _owner_and_index(e, world)
    assert idx.max() == shard - 1 and np.bincount(owner).tolist() == [shard] * world
    for r in range(world):
        back = _flat_offset(np.arange(shard), world, r)
        assert np.array_equal(np.sort(e[owner == r]), np.sort(back))             # the optimiser visits exactly what it owns
        assert np.array_equal(_owner_and_index(back, world)[1], np.arange(shard))  # and finds slot index i for local index i
    # a 32-float piece pushed by one epilogue thread never straddles two owners; consecutive pieces go to consecutive ranks
    first = np.arange(0, total, GRANULE)
    assert np.array_equal(owner[first], owner[first + GRANULE - 1]) and np.array_equal(owner[first], (first // GRANULE) % world)


@pytest.mark.parametrize("world", [2, 8])
def test_fused_exchange_emulated_equals_mean_gradient_sgd(world):
    """scatter into per-source slots -> slot sum in rank order / world -> SGD on the owned slice -> broadcast, in NumPy with
    the kernels' index maps: every rank ends with w - lr * mean_r(g_r), bit-identical across ranks."""
    rng = np.random.default_rng(world)
    total, lr = GRANULE * world * 5, np.float32(0.01)
    shard = total // world
    w0 = rng.standard_normal(total).astype(np.float32)
    grads = [rng.standard_normal(total).astype(np.float32) for _ in range(world)]
    W = [w0.copy() for _ in range(world)]
    slots = [np.zeros((world, shard), np.float32) for _ in range(world)]   # slots[owner][source]
    e = np.arange(total)
    owner, idx = _owner_and_index(e, world)
    for src in range(world):                                               # k_scatter_grad / scattering epilogue on rank src
        for o in range(world):
            m = owner == o
            slots[o][src, idx[m]] = grads[src][m]
    for r in range(world):                                                 # k_fused_dp_sgd on rank r
        i = np.arange(shard)
        g = np.zeros(shard, np.float32)
        for p in range(world):
            g = g + slots[r][p]
        g = g * np.float32(1.0 / world)
        flat = _flat_offset(i, world, r)
        new = W[r][flat] - g * lr
        for p in range(world):
            W[p][flat] = new
    acc = np.zeros(total, np.float32)
    for p in range(world):
        acc = acc + grads[p]
    want = w0 - (acc * np.float32(1.0 / world)) * lr
    for r in range(world):
        assert np.array_equal(W[r], W[0]) and np.array_equal(W[r], want)


@pytest.mark.parametrize("world", [2, 8])
def test_bucket_copy_geometry_matches_ownership(world):
    """The all-gather of the bucket protocol is ONE strided (2-D) peer copy per peer: rows of G floats, pitch G * world, starting
    at the rank's first granule of the bucket.  With the engine's 64 KB granules that copy must cover exactly the elements the
    ownership map gives the rank inside a granule*world aligned bucket -- nothing else, nothing twice."""
    G = 1 << 14
    off, n = 3 * G * world, 5 * G * world              # a bucket: [off, off + n)
    e = np.arange(off, off + n)
    owner = (e // G) % world
    for r in range(world):
        rows = (n // G) // world
        first = off + r * G
        copied = (first + np.arange(rows)[:, None] * (G * world) + np.arange(G)[None, :]).reshape(-1)
        assert np.array_equal(np.sort(copied), e[owner == r])
        # and k_bucket_sgd's (row, j) -> (flat element, slice index) pair lands on the same elements, slice indices contiguous per row
        g0 = off // G
        flat = ((g0 + np.arange(rows) * world + r) * G)[:, None] + np.arange(G)[None, :]
        idx = ((g0 // world + np.arange(rows)) * G)[:, None] + np.arange(G)[None, :]
        assert np.array_equal(flat.reshape(-1), copied)
        gg = flat // G
        assert np.array_equal(((gg // world) * G + flat % G), idx)      # == the scatter side's index for the same element
