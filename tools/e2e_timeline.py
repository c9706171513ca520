"""Device timeline of the end-to-end loop (C2, one GPU) captured with CUPTI through torch.profiler: every kernel and memcpy of
three steps with its stream, start and duration -- shows where the loop with per-step input copies loses time against the
device-resident one.  Output: one line per activity (us relative to the first), gaps > 20 us on the compute stream flagged."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import frontend as F, workloads as W  # noqa: E402


def main():
    copies = "--no-copies" not in sys.argv
    torch.cuda.init()
    dev = F.Device(0)
    ws, bs = W.mlp_init(4, 4096, seed=0)
    model = W.build_mlp(dev, ws, bs, "relu")
    tr = W.Trainer(model, 1e-5)
    x, t = W.mlp_batch(8192, 4096, 4096, seed=1, rank=0)
    hx, ht = dev.pinned_empty(x.shape), dev.pinned_empty(t.shape)
    hx[...] = x
    ht[...] = t
    slots = []
    for _ in range(2):
        X, T = F.Tensor.zeros(dev, hx.shape), F.Tensor.zeros(dev, ht.shape)
        T.requires_grad = False
        X.copy_from_numpy(hx, blocking=True)
        T.copy_from_numpy(ht, blocking=True)
        slots.append((X, T))
    tr.reserve_pool(*slots[0])
    hl = [dev.pinned_empty((1,)), dev.pinned_empty((1,))]
    evs = [F.Event(dev), F.Event(dev)]

    def pipelined(n):
        if copies:
            slots[0][0].prefetch_from_numpy(hx)
            slots[0][1].prefetch_from_numpy(ht)
        for i in range(n):
            xs, ts = slots[i % 2]
            if copies and i + 1 < n:
                slots[(i + 1) % 2][0].prefetch_from_numpy(hx)
                slots[(i + 1) % 2][1].prefetch_from_numpy(ht)
            tr.step(xs, ts).read_async(hl[i % 2], evs[i % 2])
            if i > 0:
                evs[(i - 1) % 2].wait()
        evs[(n - 1) % 2].wait()

    pipelined(4)
    dev.sync()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        pipelined(4)
        dev.sync()
    path = os.path.join(ROOT, "gpurun_out", "e2e_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    last_end = {}
    for e in ev:
        s = e["args"].get("stream")
        gap = e["ts"] - last_end.get(s, e["ts"])
        last_end[s] = e["ts"] + e["dur"]
        print(f"{e['ts'] - t0:10.1f} us  +{e['dur']:8.1f}  stream {s}  gap {gap:7.1f}{'  <<<' if gap > 20 else ''}  {e['cat']:10s} {e['name'][:70]}")
    os.remove(path)
    dev.close()


if __name__ == "__main__":
    main()
