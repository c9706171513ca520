"""All BASELINE.json configurations on one GPU, next to the CPU oracle on the same inputs (bounded):
  C1 fit_sine (examples/fit_sine.rs: 2000 x 3 -> Linear(3,1), MSE, full-batch SGD, 2000 epochs)   launch-bound
  C2 mlp4096  (bench.py's workload)                                                               tensor-bound
  C3 lenet    (im[1024,1,28,28] -> Conv2D(1,6,5) -> ReLU -> Flatten -> Linear(3456,10) -> MSE)     launch/HBM-bound
  C5 mlp8192  (per-GPU shard of the 8-GPU config: 8 x Linear(8192,8192) + Sigmoid, 8192 rows)     tensor-bound
One JSON line per config: samples/s, ms/step, kernels per step, final loss vs oracle where the oracle ran the same steps."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import frontend as F, workloads as W  # noqa: E402
from oracle import oracle_np as O  # noqa: E402  (checker / CPU baseline only)


def time_steps(dev, tr, X, T, steps, warmup):
    for _ in range(warmup):
        loss = tr.step(X, T)  # keep the previous step's graph alive across the next enqueue, as the timed loop does (pool high-water mark)
    dev.sync()
    dev.profile_enable(True)
    dev.profile_gemm(2)
    dev.profile_gemm(1)
    n0 = dev.launches()
    t0 = time.perf_counter()
    dev.timer_begin()
    for _ in range(steps):
        loss = tr.step(X, T)
    ms = dev.timer_end()
    wall = (time.perf_counter() - t0) * 1e3
    dev.profile_enable(False)
    g = dev.profile_gemm(2)
    sg = dev.profile_gemm(1)
    print(json.dumps({"device_ms_per_step": ms / steps, "wall_ms_per_step": wall / steps, "tc_gemm": {"launches": g[0], "ms": g[1], "tflops": g[2] / max(g[1], 1e-9) / 1e9},
                      "simt_gemm": {"launches": sg[0], "ms": sg[1]}}), flush=True)
    return max(ms, wall) / steps, (dev.launches() - n0) / steps, loss.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--cpu", type=int, default=1)
    a = ap.parse_args()
    dev = F.Device(0)
    if os.environ.get("AGRS_NO_FWD_FUSION"):  # A/B: Linear / Conv2D and the activation behind it as two calls
        dev.set_forward_fusion(False)
    out = []

    def emit(r):
        print(json.dumps(r), flush=True)
        out.append(r)

    if a.only in ("", "c1"):
        x, y = O.fit_sine_data(2000, np.float32)
        steps = 2000
        model = W.build_mlp(dev, [np.zeros((3, 1), np.float32)], [np.zeros((1, 1), np.float32)], None)
        tr = W.Trainer(model, 1e-3)
        X, T = F.Tensor.from_numpy(dev, x), F.Tensor.from_numpy(dev, y)
        T.requires_grad = False
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = tr.step(X, T)
        final = loss.item()
        wall = time.perf_counter() - t0
        r = {"config": "C1 fit_sine", "steps": steps, "ms_per_step": wall * 1e3 / steps, "samples_per_s": 2000 * steps / wall, "final_loss": final,
             "kernels_per_step": dev.launches() / steps, "bound": "launch latency (10 tiny kernels per step)"}
        if a.cpu:
            t0 = time.perf_counter()
            want, _ = O.mlp_train(x, y, [np.zeros((3, 1))], [np.zeros((1, 1))], None, 1e-3, steps, np.float32)
            cw = time.perf_counter() - t0
            r.update(oracle_final_loss=want[-1], rel_err=abs(final - want[-1]) / want[-1], cpu_oracle_samples_per_s=2000 * steps / cw)
        emit(r)
        # the same 2000 iterations replayed from a CUDA graph of 10 unrolled steps
        model = W.build_mlp(dev, [np.zeros((3, 1), np.float32)], [np.zeros((1, 1), np.float32)], None)
        tr = W.Trainer(model, 1e-3)
        X, T = F.Tensor.from_numpy(dev, x), F.Tensor.from_numpy(dev, y)
        T.requires_grad = False
        n0 = dev.launches()
        tr.capture(X, T, unroll=10, warmup=0)
        tr.replay(1)   # first launch: graph upload, allocations backed, clocks back up after the CPU leg
        dev.sync()
        t0 = time.perf_counter()
        loss = tr.replay((steps - 10) // 10)
        final = loss.item()
        wall = time.perf_counter() - t0
        rg = {"config": "C1 fit_sine, CUDA graph of 10 steps", "steps": steps, "ms_per_step": wall * 1e3 / (steps - 10), "samples_per_s": 2000 * (steps - 10) / wall,
              "final_loss": final, "graph_kernel_nodes": tr._graph.kernel_nodes, "graph_total_nodes": tr._graph.total_nodes,
              "kernels_per_step": (dev.launches() - n0) / steps}
        if a.cpu:
            rg.update(oracle_final_loss=want[-1], rel_err=abs(final - want[-1]) / want[-1])
        emit(rg)
        tr._graph.close()
    if a.only in ("", "c3"):
        rng = np.random.default_rng(31)
        B = 1024
        im = rng.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)
        t = rng.uniform(0, 1, (B, 10)).astype(np.float32)
        fb = W.kaiming_bound((5, 5, 6))
        filt = rng.uniform(-fb, fb, (5, 5, 6)).astype(np.float32)
        cb = rng.uniform(-0.1, 0.1, (6,)).astype(np.float32)
        w = rng.uniform(-0.05, 0.05, (3456, 10)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, (1, 10)).astype(np.float32)
        net = W.build_lenet(dev, filt, cb, w, b)
        tr = W.Trainer(net, 1e-3)
        X, T = F.Tensor.from_numpy(dev, im), F.Tensor.from_numpy(dev, t)
        T.requires_grad = False
        ms, kps, final = time_steps(dev, tr, X, T, 50, 5)
        r = {"config": "C3 lenet B=1024", "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "kernels_per_step": kps, "final_loss": final,
             "algorithmic_bytes_per_step": 0.49e9, "bound": "launch latency / HBM"}
        if a.cpu:
            conv, lin = O.OConvLayer(filt.copy(), cb.copy(), 1), O.OLinearLayer(w.copy(), b.copy())
            params = conv.parameters() + lin.parameters()
            OX, OT = O.OTensor(im), O.OTensor(t, requires_grad=False)
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                h = O.OpFlatten().forward([O.OpReLU().forward([conv.forward(OX.clone())])])
                loss = O.OpMSE().forward([lin.forward(h), OT.clone()])
                loss.backward()
                O.sgd(params, 1e-3)
                O.model_zero_grad(params)
            r["cpu_oracle_samples_per_s"] = B * n / (time.perf_counter() - t0)
        emit(r)
        tr.capture(X, T, unroll=1, warmup=2)
        t0 = time.perf_counter()
        tr.replay(1)  # the first launch uploads the graph and backs its allocations with physical memory
        dev.sync()
        first_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        loss = tr.replay(50)
        final = loss.item()
        wall = time.perf_counter() - t0
        emit({"config": "C3 lenet B=1024, CUDA graph", "ms_per_step": wall * 1e3 / 50, "samples_per_s": B * 50 / wall, "final_loss": final,
              "graph_kernel_nodes": tr._graph.kernel_nodes, "graph_total_nodes": tr._graph.total_nodes, "first_replay_ms": first_ms})
        tr._graph.close()
    for name in ("mlp4096", "mlp8192"):
        if a.only not in ("", {"mlp4096": "c2", "mlp8192": "c5"}[name]):
            continue
        cfg = W.CONFIGS[name]
        ws, bs = W.mlp_init(cfg["layers"], cfg["width"], seed=0)
        model = W.build_mlp(dev, ws, bs, cfg["act"])
        del ws, bs
        tr = W.Trainer(model, cfg["lr"])
        x, t = W.mlp_batch(cfg["rows"], cfg["width"], cfg["width"], seed=1)
        X, T = F.Tensor.from_numpy(dev, x), F.Tensor.from_numpy(dev, t)
        T.requires_grad = False
        ms, kps, final = time_steps(dev, tr, X, T, 10 if name == "mlp4096" else 5, 3)
        fl = 6.0 * cfg["layers"] * cfg["width"] ** 2 * cfg["rows"]
        emit({"config": {"mlp4096": "C2 ", "mlp8192": "C5 (one GPU's shard) "}[name] + name, "ms_per_step": ms, "samples_per_s": cfg["rows"] / (ms * 1e-3),
              "kernels_per_step": kps, "final_loss": final, "algorithmic_tflops": fl / (ms * 1e-3) / 1e12, "bound": "tensor pipe (3xTF32)"})
        del tr, model, X, T
    dev.close()


if __name__ == "__main__":
    main()
