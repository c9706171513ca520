"""Where the end-to-end loop of bench.py loses time against the device-resident loop (C2, one GPU): the per-step loss read
(a host sync: the GPU idles until the host has enqueued the next step's first kernels), the input copies, or both.
Prints one JSON line per loop variant (ms per step, CUDA events + wall clock)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import frontend as F, workloads as W  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
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
    for _ in range(3):
        tr.step(*slots[0])
    dev.sync()

    def loop(n, copies, read_every_step, copy_t=True):
        if copies:
            slots[0][0].prefetch_from_numpy(hx)
            if copy_t:
                slots[0][1].prefetch_from_numpy(ht)
        loss = None
        for i in range(n):
            xs, ts = slots[i % 2]
            if copies and i + 1 < n:
                slots[(i + 1) % 2][0].prefetch_from_numpy(hx)
                if copy_t:
                    slots[(i + 1) % 2][1].prefetch_from_numpy(ht)
            loss = tr.step(xs, ts)
            if read_every_step:
                loss.item()
        return loss.item()

    hl = [dev.pinned_empty((1,)), dev.pinned_empty((1,))]
    evs = [F.Event(dev), F.Event(dev)]

    def pipelined(n, copies=True):
        if copies:
            slots[0][0].prefetch_from_numpy(hx)
            slots[0][1].prefetch_from_numpy(ht)
        last = None
        for i in range(n):
            xs, ts = slots[i % 2]
            if copies and i + 1 < n:
                slots[(i + 1) % 2][0].prefetch_from_numpy(hx)
                slots[(i + 1) % 2][1].prefetch_from_numpy(ht)
            tr.step(xs, ts).read_async(hl[i % 2], evs[i % 2])
            if i > 0:
                evs[(i - 1) % 2].wait()
                last = float(hl[(i - 1) % 2][0])
        evs[(n - 1) % 2].wait()
        return float(hl[(n - 1) % 2][0])

    from autograd_rs_b200 import _capi
    # the same two loops with every GEMM launch timed (event pairs): is the difference inside the GEMMs or between kernels?
    for name, fn in (("device-resident, pipelined loss read", lambda n: pipelined(n, copies=False)),
                     ("inputs copied every step, pipelined loss read", lambda n: pipelined(n, copies=True))) * 2:
        fn(2)
        dev.sync()
        dev.profile_enable(True)
        dev.profile_gemm(_capi.GEMM_TCGEN05)
        n0 = dev.launches()
        dev.timer_begin()
        fn(steps)
        ms = dev.timer_end()
        launches = dev.launches() - n0
        dev.profile_enable(False)
        g_n, g_ms, g_fl = dev.profile_gemm(_capi.GEMM_TCGEN05)
        print(json.dumps({"loop": name + " (GEMMs timed)", "ms_per_step": round(ms / steps, 3), "gemm_ms_per_step": round(g_ms / steps, 3),
                          "other_ms_per_step": round((ms - g_ms) / steps, 3), "gemm_launches_per_step": g_n / steps, "launches_per_step": launches / steps}), flush=True)

    for name, fn in (("device-resident, every step's loss read one step late (read_async)", lambda n: pipelined(n, copies=False)),
                     ("inputs copied every step, every step's loss read one step late (bench.py e2e)", lambda n: pipelined(n, copies=True))):
        fn(2)
        dev.sync()
        t0 = time.perf_counter()
        dev.timer_begin()
        fn(steps)
        ms = dev.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"loop": name, "ms_per_step_events": round(ms / steps, 3), "ms_per_step_wall": round(wall / steps, 3), "steps": steps}), flush=True)

    for name, kw in (("device-resident, loss read at the end", dict(copies=False, read_every_step=False)),
                     ("device-resident, loss read every step", dict(copies=False, read_every_step=True)),
                     ("inputs copied every step, loss read at the end", dict(copies=True, read_every_step=False)),
                     ("X only copied every step, loss read every step", dict(copies=True, read_every_step=True, copy_t=False)),
                     ("inputs copied every step, loss read every step, blocking (bench.py e2e until round 2)", dict(copies=True, read_every_step=True))):
        loop(2, **kw)
        dev.sync()
        t0 = time.perf_counter()
        dev.timer_begin()
        loop(steps, **kw)
        ms = dev.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"loop": name, "ms_per_step_events": round(ms / steps, 3), "ms_per_step_wall": round(wall / steps, 3), "steps": steps}), flush=True)
    # host time to enqueue one step while the device is far behind (no sync): the Python + ctypes + engine cost per step
    dev.sync()
    t0 = time.perf_counter()
    for _ in range(4):
        tr.step(*slots[0])
    enq = (time.perf_counter() - t0) * 1e3 / 4
    dev.sync()
    print(json.dumps({"host_enqueue_ms_per_step": round(enq, 3)}), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
