"""Device timeline of the C3 (LeNet, B = 1024) training step replayed from its CUDA graph, captured with CUPTI through
torch.profiler: every kernel / memcpy / memset node of two replays with start, duration and the gap before it."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import frontend as F, workloads as W  # noqa: E402


def main():
    torch.cuda.init()
    dev = F.Device(0)
    rng = np.random.default_rng(31)
    B = 1024
    im = rng.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)
    t = rng.uniform(0, 1, (B, 10)).astype(np.float32)
    fb = W.kaiming_bound((5, 5, 6))
    net = W.build_lenet(dev, rng.uniform(-fb, fb, (5, 5, 6)).astype(np.float32), rng.uniform(-0.1, 0.1, (6,)).astype(np.float32),
                        rng.uniform(-0.05, 0.05, (3456, 10)).astype(np.float32), rng.uniform(-0.1, 0.1, (1, 10)).astype(np.float32))
    tr = W.Trainer(net, 1e-3)
    X, T = F.Tensor.from_numpy(dev, im), F.Tensor.from_numpy(dev, t)
    T.requires_grad = False
    graph = "--eager" not in sys.argv
    if graph:
        tr.capture(X, T, unroll=1, warmup=2)
        tr.replay(3)
    else:
        for _ in range(3):
            tr.step(X, T)
    dev.sync()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if graph:
            tr.replay(3)
        else:
            for _ in range(3):
                tr.step(X, T)
        dev.sync()
    path = os.path.join(ROOT, "gpurun_out", "c3_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    last = None
    for e in ev:
        gap = e["ts"] - last if last is not None else 0.0
        last = e["ts"] + e["dur"]
        print(f"{e['ts'] - t0:9.1f} us  +{e['dur']:7.1f}  gap {gap:6.1f}  {e['cat']:10s} {e['name'][:90]}")
    os.remove(path)
    dev.close()


if __name__ == "__main__":
    main()
