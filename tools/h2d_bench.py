"""Host-to-device input copy micro-benchmark: what bounds bench.py's end-to-end (`e2e`) number.

One step of C2 moves X and T (2 x 8192 x 4096 f32 = 268 MB) from pinned host memory to the device.  This tool times that
transfer alone, per rank, with every rank of a torchrun launch copying at the same time (max over ranks = what a step sees):
  * pinned (cudaMallocHost) vs write-combined pinned (cudaHostAllocWriteCombined)
  * one cudaMemcpyAsync per tensor vs chunks spread over 2 / 4 streams
  * a kernel reading the mapped host buffer directly (zero-copy over PCIe, agrs_scalar_f32 x 1.0)
  * optionally the same copies while a GEMM loop runs on the compute stream (--under-gemm)
Prints one JSON line per variant: GB/s per rank (min over ranks) and aggregate.

  python tools/h2d_bench.py                     # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_bench.py
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256, help="MiB per step (X and T together)")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--under-gemm", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cuda.bindings import runtime as rt
    import autograd_rs_b200  # noqa: F401
    from autograd_rs_b200 import _capi

    def ck(r):
        err = r[0]
        if int(err) != 0:
            raise RuntimeError(f"CUDA error {err}")
        return r[1] if len(r) == 2 else r[1:]

    ck(rt.cudaSetDevice(local))
    ctx = _capi.Ctx(local)
    nbytes = a.mb << 20
    half = nbytes // 2
    dev = ctx.empty(nbytes // 4)
    streams = [ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking)) for _ in range(4)]
    e0, e1 = ck(rt.cudaEventCreate()), ck(rt.cudaEventCreate())
    done = [ck(rt.cudaEventCreateWithFlags(rt.cudaEventDisableTiming)) for _ in range(4)]

    def barrier():
        ck(rt.cudaDeviceSynchronize())
        if dist is not None:
            dist.barrier()

    gemm = None
    if a.under_gemm:
        n = 8192
        A, B, Cc = ctx.empty(n * 4096), ctx.empty(4096 * 4096), ctx.empty(n * 4096)
        ctx.call("agrs_fill_f32", A.ptr, C.c_float(0.5), n * 4096)
        ctx.call("agrs_fill_f32", B.ptr, C.c_float(0.25), 4096 * 4096)

        def gemm(times):
            for _ in range(times):
                ctx.call("agrs_gemm_f32", 0, 0, n, 4096, 4096, A.ptr, 4096, B.ptr, 4096, Cc.ptr, 4096, None)

    def report(name, ms_list, extra=None):
        ms = float(np.median(ms_list))
        worst = ms
        if dist is not None:
            import torch
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            worst = float(t.item())
        if rank == 0:
            row = {"variant": name, "MiB_per_rank": a.mb, "ranks": world, "ms_worst_rank": round(worst, 3), "GBps_per_rank": round(nbytes / worst / 1e6, 2),
                   "GBps_aggregate": round(world * nbytes / worst / 1e6, 1), "under_gemm": bool(a.under_gemm)}
            if extra:
                row.update(extra)
            print(json.dumps(row), flush=True)

    def timed(fn):
        out = []
        for it in range(a.iters + 2):
            barrier()
            if gemm:
                gemm(14)  # ~9 ms of tensor-core work in front of and next to the copy
            ck(rt.cudaEventRecord(e0, streams[0]))
            fn()
            ck(rt.cudaEventRecord(e1, streams[0]))
            ck(rt.cudaEventSynchronize(e1))
            ck(rt.cudaDeviceSynchronize())
            if it >= 2:
                out.append(ck(rt.cudaEventElapsedTime(e0, e1)))
        return out

    def copies(host, nstreams, nchunks):
        def fn():
            per = nbytes // nchunks
            for c in range(nchunks):
                s = streams[c % nstreams]
                if c < nstreams and s is not streams[0]:
                    ck(rt.cudaStreamWaitEvent(s, e0, 0))
                ck(rt.cudaMemcpyAsync(dev.ptr + c * per, host + c * per, per, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, s))
            for k in range(1, nstreams):
                ck(rt.cudaEventRecord(done[k], streams[k]))
                ck(rt.cudaStreamWaitEvent(streams[0], done[k], 0))
        return fn

    for kind, flags in (("pinned", rt.cudaHostAllocDefault), ("pinned write-combined", rt.cudaHostAllocWriteCombined)):
        host = ck(rt.cudaHostAlloc(nbytes, flags))
        t0 = time.perf_counter()
        C.memset(host, 1, nbytes)  # touch every page (first touch decides the NUMA node)
        fill_s = time.perf_counter() - t0
        for ns, nc in ((1, 2), (2, 2), (2, 8), (4, 4), (4, 16)):
            report(f"{kind}: cudaMemcpyAsync, {nc} chunks on {ns} stream(s)", timed(copies(host, ns, nc)), {"host_fill_GBps": round(nbytes / fill_s / 1e9, 2)})
        if kind == "pinned":
            # zero-copy: a grid-stride kernel reads the pinned buffer through its device alias (unified addressing)
            hs = ctx.L.agrs_ctx_create_on_stream
            zctx = _capi.Ctx(local, stream=C.c_void_p(int(streams[0])))

            def zc():
                zctx.call("agrs_scalar_f32", _capi.SMUL, dev.ptr, C.c_void_p(host), C.c_float(1.0), nbytes // 4)
            try:
                report("pinned: kernel reads the host buffer (zero-copy), writes HBM", timed(zc))
            except Exception as ex:  # noqa: BLE001
                if rank == 0:
                    print(json.dumps({"variant": "zero-copy kernel", "error": str(ex)}), flush=True)
            zctx.close()
        ck(rt.cudaFreeHost(host))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
