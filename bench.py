#!/usr/bin/env python
"""Benchmark of the autograd.rs hot path on B200: training samples/s of the BASELINE.json MLP configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config mlp4096|mlp8192|mlp_small] [--impl b200|reference]

A "step" is one training iteration exactly as the reference's examples spell it (examples/fit_sine.rs:64-93):
forward -> MSELoss -> backward -> SGD step -> zero_grad, through the reference-shaped front end
(autograd.rs_b200/frontend.py -> libagrs_engine.so -> libagrs_b200.so).  N > 1: one process per GPU (torchrun),
minibatch rows sharded (weak scaling: 8192 rows per GPU), parameter gradients averaged with NCCL over NVLink.

One JSON line on stdout (rank 0):
  value      samples/s, whole job, inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks
  e2e        the same loop with X and T copied from pinned host memory every step and the loss read back every step
  roofline   the dominant kernel (tcgen05 split GEMM: tf32 hi*hi + bf16 correction terms) timed live with CUDA-event pairs around every launch
  cpu_baseline  the oracle (NumPy restatement of the reference's CPU path) on the host cores, bounded sample
--impl reference times that CPU path alone (the Rust reference cannot be built here: no cargo/rustc in the image).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_samples_per_sec"
UNIT = "samples/s"


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything else any library prints there (NCCL prints its version to stdout when
    NCCL_DEBUG is VERSION or WARN) is re-routed to stderr; emit() writes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


def flops_per_sample(cfg):
    # fwd + dX + dW of every Linear, 2*W*W each (the reference also computes dX of the first layer, operators.rs:159)
    return 6.0 * cfg["layers"] * cfg["width"] ** 2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md).  The sampler runs from
    before the warm-up (nvidia-smi needs ~0.3 s to produce its first line); only samples whose timestamp falls inside a
    timed window are used."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path, self.windows = gpu_index, None, None, []

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout=6.0):
        """Block until nvidia-smi has printed its first sample: its start-up (NVML initialisation, device enumeration) takes
        a few hundred ms and contends for the driver with CUDA launches -- it must be over before anything is timed."""
        if self.proc is None:
            return
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.05)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            with open(self.path) as f:
                for line in f:
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 8 or not c[1].isdigit():
                        continue
                    try:
                        ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except ValueError:
                        continue
                    if not any(t0 - 0.03 <= ts <= t1 + 0.03 for t0, t1 in self.windows):
                        continue
                    sm.append(int(c[1]))
                    mx.append(int(c[2]))
                    try:
                        power.append(float(c[3]))
                    except ValueError:
                        pass
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=int(statistics.median(sm)), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 0) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def cpu_sample_rows(cfg, target_s):
    """rows of the batch the CPU path can train on in ~target_s seconds (calibrated on a 1536^3 f32 GEMM)"""
    n = 1536
    a = np.random.default_rng(0).standard_normal((n, n)).astype(np.float32)
    a @ a
    t0 = time.perf_counter()
    a @ a
    gfs = 2.0 * n ** 3 / (time.perf_counter() - t0) / 1e9
    rows = int(target_s * gfs * 1e9 / flops_per_sample(cfg))
    rows = max(64, min(cfg["rows"], rows // 64 * 64))
    return rows, gfs


def cpu_reference_run(cfg, steps, warmup, target_s_per_step):
    """The reference's CPU path (oracle restatement, NumPy f32 + the host BLAS with all its threads) on a bounded
    sample: `rows` of the batch, same model.  Returns (samples/s, rows, seconds per step, threads, calibration GF/s)."""
    from oracle import oracle_np as O
    from autograd_rs_b200 import workloads as W  # data generation only (NumPy RNG); no device code runs here
    rows, gfs = cpu_sample_rows(cfg, target_s_per_step)
    ws, bs = W.mlp_init(cfg["layers"], cfg["width"], seed=0)
    x, t = W.mlp_batch(rows, cfg["width"], cfg["width"], seed=1)
    layers = [O.OLinearLayer(w, b) for w, b in zip(ws, bs)]
    act = {"relu": O.OpReLU, "sigmoid": O.OpSigmoid}[cfg["act"]]
    params = [p for l in layers for p in l.parameters()]
    X, T = O.OTensor(x), O.OTensor(t, requires_grad=False)

    def step():
        h = X.clone()
        for i, l in enumerate(layers):
            h = l.forward(h)
            if i + 1 < len(layers):
                h = act().forward([h])
        loss = O.OpMSE().forward([h, T.clone()])
        loss.backward()
        O.sgd(params, cfg["lr"])
        O.model_zero_grad(params)
        return float(loss.data[0, 0])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    assert np.isfinite(loss)
    return rows / dt, rows, dt, cpu_threads(), gfs


def cpu_one_core_run(cfg, target_s):
    """The same iteration through the single-threaded C restatement (oracle/oracle_c.c: matrixmultiply-style packed sgemm, the
    reference's own pass structure) -- the reference is single-threaded, so this is the closer stand-in for `cargo run --release`.
    Returns (samples/s, rows, seconds, calibration GF/s)."""
    from oracle import oracle_c as C
    from autograd_rs_b200 import workloads as W  # data generation only
    n = 1024
    a = np.random.default_rng(0).standard_normal((n, n)).astype(np.float32)
    C.dot(a, a)
    t0 = time.perf_counter()
    C.dot(a, a)
    gfs = 2.0 * n ** 3 / (time.perf_counter() - t0) / 1e9
    rows = int(target_s * gfs * 1e9 / flops_per_sample(cfg))
    rows = max(64, min(cfg["rows"], rows // 64 * 64))
    ws, bs = W.mlp_init(cfg["layers"], cfg["width"], seed=0)
    ws, bs = [np.ascontiguousarray(w, np.float32) for w in ws], [np.ascontiguousarray(b, np.float32) for b in bs]
    x, t = W.mlp_batch(rows, cfg["width"], cfg["width"], seed=1)
    t0 = time.perf_counter()
    loss = C.mlp_step(ws, bs, x, t, cfg["act"], cfg["lr"])
    dt = time.perf_counter() - t0
    assert np.isfinite(loss)
    return rows / dt, rows, dt, gfs


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, rows, dt, threads, gfs = cpu_reference_run(cfg, args.steps, args.warmup, target_s_per_step=2.0)
    sample = f"{rows} of {cfg['rows']} rows per step, same {cfg['layers']}x{cfg['width']} {cfg['act']} MLP, NumPy f32 restatement (oracle/), host BLAS {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "the Rust reference cannot be compiled in this image (no cargo/rustc); its ndarray/matrixmultiply path is single-threaded, this port uses every BLAS thread"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, cfg, world):
    return {"workload": f"{args.config}: {cfg['layers']} x Linear({cfg['width']},{cfg['width']}) + {cfg['act']} between, MSELoss, SGD; "
                        f"{cfg['rows']} rows per GPU ({ {'mlp4096': 'BASELINE.json configs[1]', 'mlp8192': 'BASELINE.json configs[4], per-GPU shard'}.get(args.config, 'test-size') })",
            "rows_per_gpu": cfg["rows"], "global_batch": cfg["rows"] * world, "width": cfg["width"], "layers": cfg["layers"],
            "activation": cfg["act"], "lr": cfg["lr"], "parallelism": f"dp{world}",
            "gradient_exchange": ("none" if world == 1 else ("fused over NVLink peer memory: reduce-scatter pushed from the dW GEMM epilogue / scatter kernel, then one slot-sum + SGD + all-gather kernel" if getattr(args, "fused", False)
                                                              else "NCCL all-reduce(avg) per parameter" + ("" if getattr(args, "no_overlap", False) else ", overlapped with backward"))),
            "l2": "working set (activations + weights, > 2 GB) is larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py --gpus {args.gpus} ...")
        args.gpus = world

    import autograd_rs_b200  # noqa: F401
    from autograd_rs_b200 import _capi, frontend as F, workloads as W

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dev = F.Device(local)
    dp = None
    ws, bs = W.mlp_init(cfg["layers"], cfg["width"], seed=0)            # identical on every rank (replicated parameters)
    model = W.build_mlp(dev, ws, bs, cfg["act"])
    del ws, bs
    if world > 1:
        ids = [F.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dp = F.DataParallel(dev, world, rank, ids[0], model.parameters(), overlap=not args.no_overlap)
        if args.fused:
            def all_gather_bytes(h):
                out = [None] * world
                dist.all_gather_object(out, h)
                return out
            if not dp.enable_fused(all_gather_bytes):   # collective decision: all ranks or none
                args.fused = False
                if rank == 0:
                    print(f"fused peer-memory optimiser unavailable ({dp.fused_error}); using the NCCL exchange", file=sys.stderr)
    tr = W.Trainer(model, cfg["lr"], dp=dp)

    x, t = W.mlp_batch(cfg["rows"], cfg["width"], cfg["width"], seed=1, rank=rank)  # this rank's shard of the global batch
    hx, ht = dev.pinned_empty(x.shape), dev.pinned_empty(t.shape)
    hx[...] = x
    ht[...] = t
    del x, t
    X, T = F.Tensor.zeros(dev, hx.shape), F.Tensor.zeros(dev, ht.shape)
    T.requires_grad = False
    X.copy_from_numpy(hx, blocking=True)
    T.copy_from_numpy(ht, blocking=True)

    def barrier():
        dev.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        tt = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident throughput ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    barrier()
    pool_high = tr.reserve_pool(X, T)   # one untimed iteration sizes the memory pool: no pool growth inside the timed regions
    for _ in range(args.warmup):
        loss = tr.step(X, T)
    barrier()
    dev.profile_enable(True)
    dev.profile_gemm(_capi.GEMM_TCGEN05)
    dev.profile_gemm(_capi.GEMM_SIMT)
    n0 = dev.launches()
    w0 = time.time()
    dev.timer_begin()
    for _ in range(args.steps):
        loss = tr.step(X, T)
    ms = dev.timer_end()
    sampler.window(w0, time.time())
    barrier()
    launches = dev.launches() - n0
    dev.profile_enable(False)
    g_n, g_ms, g_fl = dev.profile_gemm(_capi.GEMM_TCGEN05)
    s_n, s_ms, s_fl = dev.profile_gemm(_capi.GEMM_SIMT)
    ms = max_over_ranks(ms)
    final_loss = loss.item()
    assert np.isfinite(final_loss), "training diverged"
    value = cfg["rows"] * world * args.steps / (ms * 1e-3)

    # ---- end to end: host buffers in, loss out, every step ----
    # Every step's X and T come from pinned host memory (268 MB) and its loss goes back to the host.  The copy of step i+1
    # runs on the copy stream while step i computes (two device input slots); the loss read blocks once per step.
    e2e = None
    if not args.no_e2e:
        X2, T2 = F.Tensor.zeros(dev, hx.shape), F.Tensor.zeros(dev, ht.shape)
        T2.requires_grad = False
        slots = [(X, T), (X2, T2)]

        def e2e_loop(n):
            slots[0][0].prefetch_from_numpy(hx)     # step 0's inputs: pinned host -> device
            slots[0][1].prefetch_from_numpy(ht)
            lv = None
            for i in range(n):
                dev.prefetch_join()                 # step i's inputs have landed
                xs, ts = slots[i % 2]
                if i + 1 < n:                       # step i+1's inputs start moving now, under step i's kernels
                    slots[(i + 1) % 2][0].prefetch_from_numpy(hx)
                    slots[(i + 1) % 2][1].prefetch_from_numpy(ht)
                lv = tr.step(xs, ts).item()         # device -> host read of the loss (4 bytes), blocks
            return lv

        e2e_loop(2)
        barrier()
        t0 = time.perf_counter()
        w0 = time.time()
        dev.timer_begin()
        lv = e2e_loop(args.steps)
        ms_e = dev.timer_end()
        wall_e = (time.perf_counter() - t0) * 1e3
        sampler.window(w0, time.time())
        ms_e = max_over_ranks(max(ms_e, wall_e))
        assert np.isfinite(lv)
        e2e = {"value": cfg["rows"] * world * args.steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(hx.nbytes + ht.nbytes),
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e / args.steps,
               "how": "X,T copied from pinned host memory every step on a copy stream (double-buffered, overlapping the previous step), loss read back every step"}
        del X2, T2, slots

    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel ----
    peaks = load_peaks()
    # tf32 tensor rate = 1/2 of bf16 (nominal 1.1 vs 2.25 PFLOP/s dense).  Per algorithmic product the split GEMM issues
    #   cross = 0: three tf32 MMAs                      -> 3 tf32 slots, f32-faithful peak = bf16/6
    #   cross > 0: one tf32 MMA + two bf16 MMAs (1/2 slot each) -> 2 tf32 slots, peak = bf16/4
    # The GEMMs run inside a long step -> sustained figure.
    cross = dev.get_tuning(_capi.TUNE_TC_CROSS)
    slots = {0: 3.0, 3: 1.5, 4: 1.5}.get(cross, 2.0)  # cross = 3 / 4: three bf16 / fp16 MMAs, no tf32 pass
    peak_alg = peaks["bf16_sustained"] / 2.0 / slots
    roofline = None
    if g_n:
        ach = g_fl / (g_ms * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak_alg, "unit": "TFLOP/s", "frac": ach / peak_alg, "traffic": traffic,
                    "kernel": "tc2::k_gemm_3xtf32_2cta (tcgen05.mma.cta_group::2: " +
                              ("kind::tf32 x3)" if cross == 0 else "3 x kind::f16 on bf16 pairs x1 + x2)" if cross == 3 else
                               "3 x kind::f16 on scaled fp16 pairs h + l, plus two absmax kernels per launch)" if cross == 4 else
                               f"kind::tf32 hi*hi + 2 x kind::f16 bf16 cross terms, tile layout {cross})"),
                    "launches": g_n, "avg_launch_ms": g_ms / g_n,
                    "algorithmic_flops_per_launch": g_fl / g_n, "executed_tf32_equivalent_tflops": slots * ach,
                    "gemm_share_of_step": g_ms / ms, "cross_terms": {0: "tf32", 3: "bf16 x3", 4: "fp16 x3 scaled"}.get(cross, "bf16"),
                    # the same fraction against the BURST bf16 figure: what ncu's tensor-pipe-active percentage corresponds to
                    "frac_of_burst_peak": ach / (peaks["bf16_burst"] / 2.0 / slots),
                    "peak_note": f"{peaks['source']} bf16 sustained {peaks['bf16_sustained']} TFLOP/s / 2 (tf32 rate) / {slots:g} (tf32-rate MMA slots per product)"}
        # context: what cuBLAS itself reaches on this pool's B200s for PLAIN tf32 (one MMA per product) -- tools/tf32_cublas_peak.py
        cub = os.path.join(ROOT, "profiles", "r1_cublas_peaks.json")
        if os.path.exists(cub):
            try:
                c = json.load(open(cub))
                roofline["cublas_tf32_tflops"] = {"burst": c["tf32"]["burst_tflops"], "sustained": c["tf32"]["sustained_tflops"]}
                roofline["executed_frac_of_cublas_tf32_sustained"] = slots * ach / c["tf32"]["sustained_tflops"]
                roofline["cublas_f32_simt_tflops"] = c["f32_simt"]["sustained_tflops"]
            except Exception:
                pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, world),
        "algorithmic_tflops_per_gpu": flops_per_sample(cfg) * cfg["rows"] * args.steps / (ms * 1e-3) / 1e12,
        "final_loss": final_loss, "pool_used_high_bytes": pool_high, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "simt_gemm": {"launches": s_n, "ms": s_ms},
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, rows, dt, threads, gfs = cpu_reference_run(cfg, steps=1, warmup=0, target_s_per_step=15.0)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"1 step on {rows} of {cfg['rows']} rows, same model, NumPy f32 restatement of the reference path (oracle/), "
                                          f"host BLAS {threads} threads ({gfs:.0f} GFLOP/s on a 1536^3 sgemm), {dt:.1f} s"}
        try:
            v1, r1, d1, g1 = cpu_one_core_run(cfg, target_s=8.0)
            line["cpu_baseline"]["one_core"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                                                "sample": f"1 step on {r1} of {cfg['rows']} rows, single-threaded C restatement (oracle/oracle_c.c, "
                                                          f"{g1:.0f} GFLOP/s on a 1024^3 sgemm), {d1:.1f} s"}
        except (OSError, RuntimeError, subprocess.CalledProcessError) as e:  # no compiler on the box and no prebuilt library
            line["cpu_baseline"]["one_core"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if dp is not None:
        dp.close()
    if rank == 0:
        emit(line)
    del tr, model, X, T, loss
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="mlp4096", choices=["mlp4096", "mlp8192", "mlp_small"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce gradients after backward instead of during it")
    ap.add_argument("--fused", action="store_true", help="N > 1: one fused reduce-scatter + SGD + all-gather kernel over NVLink peer memory "
                                                         "instead of NCCL all-reduce + SGD (the default)")
    ap.add_argument("--nccl", action="store_true", help="N > 1: NCCL all-reduce per parameter + SGD kernel (the baseline the fused kernel replaces)")
    args = ap.parse_args()
    claim_stdout()
    args.fused = not args.nccl
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    from autograd_rs_b200 import workloads as W
    cfg = dict(W.CONFIGS[args.config])
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_b200_arm(args, cfg)


if __name__ == "__main__":
    main()
