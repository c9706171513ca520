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
    a @ a          # the BLAS thread pool is up (the first calls after a thread-count change run at a fraction of the speed)
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        a @ a
        best = min(best, time.perf_counter() - t0)
    gfs = 2.0 * n ** 3 / best / 1e9
    rows = int(target_s * gfs * 1e9 / flops_per_sample(cfg))
    rows = max(64, min(cfg["rows"], rows // 64 * 64))
    return rows, gfs


def set_blas_threads(n=None):
    """Give the host BLAS every core whatever the launcher exported: torchrun sets OMP_NUM_THREADS=1 for its workers, which
    would make the CPU arm 4-5x slower at N > 1 than at N = 1 for no reason that has to do with the GPUs."""
    n = n or os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n, user_api="blas")
    except Exception:
        pass
    return cpu_threads()


class OracleMLP:
    """The reference's CPU path (oracle restatement, NumPy f32 + host BLAS) for one MLP configuration, one training iteration
    at a time; `hold_grads` keeps the parameter gradients of the last backward so the parity block can compare them."""

    def __init__(self, cfg, ws, bs, x, t):
        from oracle import oracle_np as O
        self.O, self.cfg = O, cfg
        self.layers = [O.OLinearLayer(w.copy(), b.copy()) for w, b in zip(ws, bs)]
        self.act = {"relu": O.OpReLU, "sigmoid": O.OpSigmoid}[cfg["act"]]
        self.params = [p for l in self.layers for p in l.parameters()]
        self.X, self.T = O.OTensor(x), O.OTensor(t, requires_grad=False)
        self.grads = None

    def step(self, hold_grads=False, tie_fix=None):
        """one training iteration; tie_fix(i, z) may patch hidden pre-activation i in place (see oracle_parity); the time it takes
        is kept in self.excluded_s so the caller can leave it out of the CPU timing"""
        O = self.O
        h = self.X.clone()
        for i, l in enumerate(self.layers):
            h = l.forward(h)
            if i + 1 < len(self.layers):
                if tie_fix is not None:
                    t0 = time.perf_counter()
                    tie_fix(i, h.data)
                    self.excluded_s += time.perf_counter() - t0
                h = self.act().forward([h])
        loss = O.OpMSE().forward([h, self.T.clone()])
        loss.backward()
        if hold_grads:
            self.grads = [np.array(p.grad, copy=True) for p in self.params]
        O.sgd(self.params, self.cfg["lr"])
        O.model_zero_grad(self.params)
        return float(loss.data[0, 0])

    excluded_s = 0.0


def cpu_reference_run(cfg, steps, warmup, target_s_per_step, hold_grads=False, tie_fix_factory=None):
    """The reference's CPU path on a bounded sample: `rows` of the batch, same model, every BLAS thread.
    Returns a dict: samples/s, rows, seconds per step, threads, calibration GF/s, first loss (+ the stepper when hold_grads)."""
    from autograd_rs_b200 import workloads as W  # data generation only (NumPy RNG); no device code runs here
    threads = set_blas_threads()
    rows, gfs = cpu_sample_rows(cfg, target_s_per_step)
    ws, bs = W.mlp_init(cfg["layers"], cfg["width"], seed=0)
    x, t = W.mlp_batch(cfg["rows"], cfg["width"], cfg["width"], seed=1)  # rank 0's batch of the timed arm: its first `rows` rows
    om = OracleMLP(cfg, ws, bs, x[:rows], t[:rows])
    tie_fix = tie_fix_factory(ws, bs, x[:rows], t[:rows]) if tie_fix_factory else None
    del ws, bs
    first = None
    for _ in range(warmup):
        l = om.step()
        first = l if first is None else first
    t0 = time.perf_counter()
    for i in range(steps):
        loss = om.step(hold_grads=hold_grads and i == 0 and warmup == 0, tie_fix=tie_fix if i == 0 and warmup == 0 else None)
        first = loss if first is None else first
    dt = (time.perf_counter() - t0 - om.excluded_s) / max(steps, 1)
    assert np.isfinite(loss)
    return {"value": rows / dt, "rows": rows, "dt": dt, "threads": threads, "gfs": gfs, "loss0": first, "oracle": om if hold_grads else None}


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
    r = cpu_reference_run(cfg, args.steps, args.warmup, target_s_per_step=2.0)
    val, rows, dt, threads = r["value"], r["rows"], r["dt"], r["threads"]
    sample = f"{rows} of {cfg['rows']} rows per step, same {cfg['layers']}x{cfg['width']} {cfg['act']} MLP, NumPy f32 restatement (oracle/), host BLAS {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "the Rust reference cannot be compiled in this image (no cargo/rustc); its ndarray/matrixmultiply path is single-threaded, "
                                 "this port uses every BLAS thread (set explicitly: the same count at every --gpus N, whatever OMP_NUM_THREADS the launcher exports)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:  # the single-threaded figure (what `cargo run --release` of the reference would use: Cargo.toml enables no threading)
        v1, r1, d1, g1 = cpu_one_core_run(cfg, target_s=6.0)
        line["cpu_baseline"]["one_core"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                                            "sample": f"1 step on {r1} of {cfg['rows']} rows, single-threaded C restatement (oracle/oracle_c.c, {g1:.0f} GFLOP/s on a 1024^3 sgemm), {d1:.1f} s"}
    except (OSError, RuntimeError, subprocess.CalledProcessError) as e:
        line["cpu_baseline"]["one_core"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    emit(line)


def workload_config(args, cfg, world):
    return workload_config_for(args.config, cfg, world, getattr(args, "fused", False), getattr(args, "no_overlap", False), getattr(args, "fused_mono", False))


# ------------------------------------------------------------------------------------------------ B200 arm
class Job:
    """what every part of the B200 arm needs: device, ranks, the launcher's collectives"""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        from autograd_rs_b200 import frontend as F
        self.dev = F.Device(self.local)
        if getattr(args, "no_forward_fusion", False):
            self.dev.set_forward_fusion(False)  # A/B: Linear and its activation as two engine calls (same bits)
        if getattr(args, "act_warps", False):
            from autograd_rs_b200 import _capi
            self.dev.set_tuning(_capi.TUNE_TC_ACT_WARPS, 1)  # A/B: the activation written inside the forward GEMM (same bits)

    def barrier(self):
        self.dev.sync()
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        import torch
        tt = torch.tensor([v], dtype=torch.float64, device=f"cuda:{self.local}")
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def all_gather_bytes(self, h):
        out = [None] * self.world
        self.dist.all_gather_object(out, h)
        return out

    def data_parallel(self, params, fused, overlap=True):
        """DataParallel over this job's ranks (None on one GPU); returns (dp, fused actually in use)"""
        if self.world == 1:
            return None, False
        from autograd_rs_b200 import frontend as F
        ids = [F.comm_unique_id() if self.rank == 0 else None]   # an NCCL id is good for ONE communicator
        self.dist.broadcast_object_list(ids, src=0)
        dp = F.DataParallel(self.dev, self.world, self.rank, ids[0], params, overlap=overlap)
        if fused and not dp.enable_fused(self.all_gather_bytes):     # collective decision: all ranks or none
            fused = False
            if self.rank == 0:
                print(f"fused peer-memory optimiser unavailable ({dp.fused_error}); using the NCCL exchange", file=sys.stderr)
        return dp, fused

    def replicas_identical(self, arrays):
        """every rank holds the same bytes (MIN == MAX over ranks, element-wise)"""
        import torch
        for a in arrays:
            tt = torch.from_numpy(np.ascontiguousarray(a)).cuda(self.local)
            lo, hi = tt.clone(), tt.clone()
            self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN)
            self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                return False
        return True


def dp_parity(job):
    """N > 1, before anything is timed: the ranks train a small MLP on disjoint shards of ONE global batch through both gradient
    exchanges (NCCL all-reduce overlapped with backward; fused peer-memory reduce-scatter + SGD + all-gather), rank 0 runs the
    oracle on the whole global batch.  Every loss (mean of the shard means) and every final weight must be within 1e-4 of the
    single-device oracle (north star) and the replicas bit-identical -- otherwise the bench stops here."""
    from autograd_rs_b200 import frontend as F, workloads as W
    world, rank, dev = job.world, job.rank, job.dev
    layers, width, rows, steps, lr = 3, 512, 256 * world, 4, 1e-4
    ws, bs = W.mlp_init(layers, width, seed=3)
    x, t = W.mlp_batch(rows, width, width, seed=4)          # the global batch, identical on every rank
    b, e = F.shard_rows(rows, world, rank)
    want = None
    if rank == 0:
        from oracle import oracle_np as O
        want, olayers = O.mlp_train(x, t, ws, bs, "sigmoid", lr, steps, np.float32)
        oparams = [q.data for l in olayers for q in l.parameters()]
    out = {"workload": f"{layers} x Linear({width},{width}) + sigmoid, global batch {rows} in {world} shards, {steps} steps, lr {lr}",
           "tolerance": 1e-4, "paths": {}}
    for name, fused, overlap in (("nccl", False, True), ("fused", True, True), ("fused_one_kernel", True, False)):
        model = W.build_mlp(dev, ws, bs, "sigmoid")
        dp, fused_on = job.data_parallel(model.parameters(), fused)
        tr = W.Trainer(model, lr, dp=dp, fused_overlap=overlap)
        X, T = F.Tensor.from_numpy(dev, x[b:e]), F.Tensor.from_numpy(dev, t[b:e])
        T.requires_grad = False
        losses = []
        for _ in range(steps):
            loss = tr.step(X, T)
            dp.mean_scalar(loss)
            losses.append(loss.item())
        params = [p.data() for p in model.parameters()]
        same = job.replicas_identical(params)
        rec = {"replicas_bit_identical": bool(same), "fused_in_use": bool(fused_on)}
        if rank == 0:
            rec["max_rel_err_loss"] = max(abs(a - c) / abs(c) for a, c in zip(losses, want))
            rec["max_rel_err_weights"] = max(O.rel_err(p, q) for p, q in zip(params, oparams))
            rec["ok"] = bool(same and rec["max_rel_err_loss"] < 1e-4 and rec["max_rel_err_weights"] < 1e-4)
        out["paths"][name] = rec
        dp.close()
        del tr, model, X, T, loss
        job.barrier()
    ok = [all(r.get("ok", False) for r in out["paths"].values()) if rank == 0 else True]
    job.dist.broadcast_object_list(ok, src=0)
    out["ok"] = bool(ok[0])
    return out


def oracle_parity(job, cfg, target_s):
    """N = 1: one iteration of the oracle on a bounded sample of the timed batch (the same call is the cpu_baseline timing), and the
    device path on exactly those rows and weights: step-0 loss and every parameter gradient, max|a-b| / max|b| per tensor <= 1e-4.

    ReLU ties: at this size (3 x 8192 x 4096 hidden pre-activations) a few dozen z land within f32 rounding of 0, where the two
    implementations' last-bit differences decide the gradient mask, and ONE flipped mask element moves a bias gradient (a sum of
    8192 terms of random sign) by ~1 % -- two correct f32 implementations cannot agree to 1e-4 there.  So the oracle takes the
    device's pre-activation at every element whose sign differs, provided both values are within 1e-5 * max|z| of the kink (anything
    else is reported as a real mismatch); the comparison then measures arithmetic, not tie-breaking.  `relu_ties_aligned` counts them."""
    from autograd_rs_b200 import frontend as F, workloads as W
    dev = job.dev
    state = {}

    def factory(ws, bs, x, t):
        # device first: forward layer by layer keeping the hidden pre-activations, MSE, backward
        model = W.build_mlp(dev, ws, bs, cfg["act"])
        X, T = F.Tensor.from_numpy(dev, x), F.Tensor.from_numpy(dev, t)
        T.requires_grad = False
        layers = model.layers()
        h, zs = X.clone(), []
        for i, l in enumerate(layers):
            h = l.forward([h])
            if isinstance(l, F.nn.Linear) and i + 1 < len(layers):
                zs.append(h)
        loss = F.nn.MeanSquaredError().forward([h, T.clone()])
        loss.backward()
        state.update(loss=loss.item(), grads=[p.grad().data() for p in model.parameters()], ties=0, bad=0, model=model)

        def tie_fix(i, z):
            if cfg["act"] != "relu":
                return
            zd = zs[i].data().reshape(z.shape)
            flip = (zd > 0) != (z > 0)
            n = int(flip.sum())
            if n:
                tol = 1e-5 * float(np.abs(z).max())
                legit = flip & (np.abs(z) <= tol) & (np.abs(zd) <= tol)
                state["bad"] += n - int(legit.sum())
                state["ties"] += int(legit.sum())
                z[legit] = zd[legit]
        return tie_fix

    r = cpu_reference_run(cfg, steps=1, warmup=0, target_s_per_step=target_s, hold_grads=True, tie_fix_factory=factory)
    om, rows = r.pop("oracle"), r["rows"]
    errs = [om.O.rel_err(g.reshape(-1), og.reshape(-1)) for g, og in zip(state["grads"], om.grads)]
    el = abs(state["loss"] - r["loss0"]) / abs(r["loss0"])
    par = {"rows": rows, "of_rows": cfg["rows"], "loss_device": state["loss"], "loss_oracle": r["loss0"], "rel_err_loss": el,
           "max_rel_err_param_grads": max(errs), "tensors_compared": len(errs), "tolerance": 1e-4,
           "relu_ties_aligned": state["ties"], "sign_mismatches_beyond_ties": state["bad"],
           "ok": bool(el < 1e-4 and max(errs) < 1e-4 and state["bad"] == 0),
           "what": "step-0 loss and all dW / db of the timed model on the first `rows` rows of the timed batch: device vs the f32 oracle "
                   "(oracle/oracle_np.py); hidden pre-activations within 1e-5 * max|z| of the ReLU kink whose sign differs take the device's value"}
    state.clear()
    return par, r


def gemm_roofline(dev, g_n, g_ms, g_fl, step_ms_total):
    """roofline block of the dominant kernel from the live event-pair timings of every GEMM launch in the timed region"""
    from autograd_rs_b200 import _capi
    peaks = load_peaks()
    # tf32 tensor rate = 1/2 of bf16 (nominal 1.1 vs 2.25 PFLOP/s dense).  Per algorithmic product the split GEMM issues
    #   cross = 0: three tf32 MMAs                      -> 3 tf32 slots, f32-faithful peak = bf16/6
    #   cross = 1/2: one tf32 MMA + two bf16 MMAs (1/2 slot each) -> 2 tf32 slots, peak = bf16/4
    #   cross = 3/4: three bf16 / fp16 MMAs             -> 1.5 tf32 slots, peak = bf16/3
    #   cross = 5 (default): mode 4 for every GEMM whose operand magnitudes the engine announces -- all of them in an MLP step
    # The GEMMs run inside a long step -> sustained figure.
    cross = dev.get_tuning(_capi.TUNE_TC_CROSS)
    slots = {0: 3.0, 3: 1.5, 4: 1.5, 5: 1.5}.get(cross, 2.0)
    peak_alg = peaks["bf16_sustained"] / 2.0 / slots
    if not g_n:
        return None
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
                           "3 x kind::f16 on power-of-two-scaled fp16 pairs h + l; one magnitude per operand, reported by the kernel that produced it)" if cross in (4, 5) else
                           f"kind::tf32 hi*hi + 2 x kind::f16 bf16 cross terms, tile layout {cross})"),
                "launches": g_n, "avg_launch_ms": g_ms / g_n,
                "algorithmic_flops_per_launch": g_fl / g_n, "executed_tf32_equivalent_tflops": slots * ach,
                "gemm_share_of_step": g_ms / step_ms_total, "cross_terms": {0: "tf32", 3: "bf16 x3", 4: "fp16 x3 scaled", 5: "fp16 x3 scaled"}.get(cross, "bf16"),
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
    return roofline


def time_config(job, name, cfg, steps, warmup, sampler, do_e2e):
    """Build the model of `cfg` on every rank, warm up, time `steps` training iterations device-resident (and end to end when
    asked).  Returns the measurements as a dict; rank-independent values are maxima over ranks."""
    from autograd_rs_b200 import _capi, frontend as F, workloads as W
    args, dev, rank, world = job.args, job.dev, job.rank, job.world
    ws, bs = W.mlp_init(cfg["layers"], cfg["width"], seed=0)            # identical on every rank (replicated parameters)
    model = W.build_mlp(dev, ws, bs, cfg["act"])
    del ws, bs
    dp, fused_on = job.data_parallel(model.parameters(), args.fused, overlap=not args.no_overlap)
    tr = W.Trainer(model, cfg["lr"], dp=dp, fused_overlap=not args.fused_mono)

    x, t = W.mlp_batch(cfg["rows"], cfg["width"], cfg["width"], seed=1, rank=rank)  # this rank's shard of the global batch
    hx, ht = dev.pinned_empty(x.shape), dev.pinned_empty(t.shape)
    hx[...] = x
    ht[...] = t
    del x, t
    X, T = F.Tensor.zeros(dev, hx.shape), F.Tensor.zeros(dev, ht.shape)
    T.requires_grad = False
    X.copy_from_numpy(hx, blocking=True)
    T.copy_from_numpy(ht, blocking=True)

    # ---- device-resident throughput ----
    job.barrier()
    pool_high = tr.reserve_pool(X, T)   # one untimed iteration sizes the memory pool: no pool growth inside the timed regions
    for _ in range(warmup):
        loss = tr.step(X, T)
    job.barrier()
    dev.profile_enable(True)
    dev.profile_gemm(_capi.GEMM_TCGEN05)
    dev.profile_gemm(_capi.GEMM_SIMT)
    n0 = dev.launches()
    w0 = time.time()
    dev.timer_begin()
    for _ in range(steps):
        loss = tr.step(X, T)
    ms = dev.timer_end()
    if sampler is not None:
        sampler.window(w0, time.time())
    job.barrier()
    launches = dev.launches() - n0
    dev.profile_enable(False)
    g_n, g_ms, g_fl = dev.profile_gemm(_capi.GEMM_TCGEN05)
    s_n, s_ms, s_fl = dev.profile_gemm(_capi.GEMM_SIMT)
    exposed = dp.fused_exposed_ms() if (dp is not None and fused_on and not args.fused_mono) else None
    ms_local = ms
    ms = job.max_over_ranks(ms)
    final_loss = loss.item()
    assert np.isfinite(final_loss), "training diverged"
    res = {"value": cfg["rows"] * world * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "algorithmic_tflops_per_gpu": flops_per_sample(cfg) * cfg["rows"] * steps / (ms * 1e-3) / 1e12,
           "final_loss": final_loss, "pool_used_high_bytes": pool_high, "gpu_launches": launches,
           "roofline": gemm_roofline(dev, g_n, g_ms, g_fl, ms_local), "simt_gemm": {"launches": s_n, "ms": s_ms},
           "fused": fused_on,
           "exchange_exposed_ms_per_step": (job.max_over_ranks(exposed[0] / max(exposed[1], 1)) if exposed else None),
           "config": workload_config_for(name, cfg, world, fused_on, args.no_overlap, args.fused_mono)}
    res["config"]["forward_activation"] = ("own engine call (A/B)" if getattr(args, "no_forward_fusion", False)
                                           else "inside the forward GEMM, by its activation warps (A/B)" if getattr(args, "act_warps", False)
                                           else "second pass of the forward GEMM's call (agrs_gemm_bias_act_f32)")

    # ---- end to end: host buffers in, loss out, every step ----
    # Every step's X and T come from pinned host memory (268 MB) and its loss goes back to the host.  The copy of step i+1
    # runs on the copy stream while step i computes (two device input slots); the loss read blocks once per step.
    if do_e2e:
        X2, T2 = F.Tensor.zeros(dev, hx.shape), F.Tensor.zeros(dev, ht.shape)
        T2.requires_grad = False
        slots = [(X, T), (X2, T2)]

        hl = [dev.pinned_empty((1,)), dev.pinned_empty((1,))]   # pinned landing slots of the per-step loss
        evs = [F.Event(dev), F.Event(dev)]

        def e2e_loop(n):
            slots[0][0].prefetch_from_numpy(hx)     # step 0's inputs: pinned host -> device
            slots[0][1].prefetch_from_numpy(ht)
            losses = []
            for i in range(n):
                xs, ts = slots[i % 2]               # each tensor carries its own landing event: the first kernel that reads X waits
                                                    # for X's copy only, T may still be on the bus until the loss needs it
                if i + 1 < n:                       # step i+1's inputs start moving once step i-1 (their slot's last reader) is done
                    slots[(i + 1) % 2][0].prefetch_from_numpy(hx)
                    slots[(i + 1) % 2][1].prefetch_from_numpy(ht)
                tr.step(xs, ts).read_async(hl[i % 2], evs[i % 2])   # device -> host copy of the loss, enqueued behind the step
                if i > 0:                           # the host reads step i-1's loss now that step i is in the queue:
                    evs[(i - 1) % 2].wait()         # every step's loss reaches the host, the device never waits for the host
                    losses.append(float(hl[(i - 1) % 2][0]))
            evs[(n - 1) % 2].wait()
            losses.append(float(hl[(n - 1) % 2][0]))
            assert len(losses) == n
            return losses[-1]

        e2e_loop(2)
        job.barrier()
        t0 = time.perf_counter()
        w0 = time.time()
        dev.timer_begin()
        lv = e2e_loop(steps)
        ms_e = dev.timer_end()
        wall_e = (time.perf_counter() - t0) * 1e3
        if sampler is not None:
            sampler.window(w0, time.time())
        ms_e = job.max_over_ranks(max(ms_e, wall_e))
        assert np.isfinite(lv)
        res["e2e"] = {"value": cfg["rows"] * world * steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(hx.nbytes + ht.nbytes),
                      "d2h_bytes_per_step": 4, "ms_per_step": ms_e / steps,
                      "how": "X,T copied from pinned host memory every step on a copy stream (double-buffered, overlapping the previous step; each tensor's first reader waits for that tensor's copy only); every step's loss copied to pinned host memory behind the step and read by the host once the next step is enqueued (agrs_download_async / agrs_event_wait)"}
        for ev in evs:
            ev.close()
        del X2, T2, slots
    if dp is not None:
        dp.close()
    del tr, model, X, T, loss
    job.barrier()
    return res


def workload_config_for(name, cfg, world, fused, no_overlap, fused_mono=False):
    return {"workload": f"{name}: {cfg['layers']} x Linear({cfg['width']},{cfg['width']}) + {cfg['act']} between, MSELoss, SGD; "
                        f"{cfg['rows']} rows per GPU ({ {'mlp4096': 'BASELINE.json configs[1]', 'mlp8192': 'BASELINE.json configs[4], per-GPU shard'}.get(name, 'test-size') })",
            "rows_per_gpu": cfg["rows"], "global_batch": cfg["rows"] * world, "width": cfg["width"], "layers": cfg["layers"],
            "activation": cfg["act"], "lr": cfg["lr"], "parallelism": f"dp{world}",
            "gradient_exchange": ("none" if world == 1 else (("fused over NVLink peer memory: reduce-scatter pushed from the dW GEMM epilogue / scatter kernel, then one slot-sum + SGD + all-gather kernel after backward" if fused_mono else
                                                               "fused over NVLink peer memory, per layer under the backward pass: reduce-scatter pushed from the dW GEMM epilogue / scatter kernel, flag round, slot-sum + SGD kernel on a side stream, all-gather by the copy engines") if fused
                                                              else "NCCL all-reduce(avg) per parameter" + ("" if no_overlap else ", overlapped with backward"))),
            "l2": "working set (activations + weights, > 2 GB) is larger than the 126 MB L2; no flush needed"}


def run_b200_arm(args, cfg):
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if world_env != args.gpus:
        if world_env == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py --gpus {args.gpus} ...")
        args.gpus = world_env

    import autograd_rs_b200  # noqa: F401
    from autograd_rs_b200 import workloads as W
    job = Job(args)
    rank, world = job.rank, job.world

    # ---- parity gates, before anything is timed ----
    parity = {}
    if world > 1 and not args.no_parity:
        parity["data_parallel"] = dp_parity(job)
        if not parity["data_parallel"]["ok"]:
            if rank == 0:
                emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "error": "data-parallel parity failed; nothing was timed", "parity": parity})
            sys.exit(1)

    sampler = ClockSampler(job.local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
        sampler.wait_ready()

    head = time_config(job, args.config, cfg, args.steps, args.warmup, sampler, do_e2e=not args.no_e2e)

    # ---- the north star's scaling configuration next to the headline: mlp8192 (BASELINE.json configs[4]), a few steps ----
    also = {}
    if args.config == "mlp4096" and not args.no_also:
        c5 = dict(W.CONFIGS["mlp8192"])
        job.dev.sync()
        job.dev.pool_trim()   # the 13 GB model starts from the memory pool a fresh process would have, not from mlp4096's blocks
        r5 = time_config(job, "mlp8192", c5, steps=args.also_steps, warmup=3, sampler=sampler, do_e2e=False)
        also["mlp8192"] = {"metric": METRIC, "unit": UNIT, **r5}
    clocks = sampler.stop() if sampler is not None else None

    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": head["config"], "algorithmic_tflops_per_gpu": head["algorithmic_tflops_per_gpu"],
        "final_loss": head["final_loss"], "pool_used_high_bytes": head["pool_used_high_bytes"], "clocks": clocks, "e2e": head.get("e2e"),
        "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "simt_gemm": head["simt_gemm"],
        "exchange_exposed_ms_per_step": head["exchange_exposed_ms_per_step"],
    }
    if also:
        line["also"] = also

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # one oracle iteration on a bounded sample: the CPU baseline AND the checker of the timed model's first step
        par, r = oracle_parity(job, cfg, target_s=15.0)
        parity[args.config] = par
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                "sample": f"1 step on {r['rows']} of {cfg['rows']} rows, same model, NumPy f32 restatement of the reference path (oracle/), "
                                          f"host BLAS {r['threads']} threads ({r['gfs']:.0f} GFLOP/s on a 1536^3 sgemm), {r['dt']:.1f} s"}
        if "mlp8192" in also:
            par5, r5 = oracle_parity(job, dict(W.CONFIGS["mlp8192"]), target_s=4.0)
            parity["mlp8192"] = par5
            also["mlp8192"]["cpu_baseline"] = {"value": r5["value"], "unit": UNIT, "cores": r5["threads"], "kind": "port",
                                               "sample": f"1 step on {r5['rows']} of 8192 rows, {r5['dt']:.1f} s"}
        try:
            v1, r1, d1, g1 = cpu_one_core_run(cfg, target_s=8.0)
            line["cpu_baseline"]["one_core"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                                                "sample": f"1 step on {r1} of {cfg['rows']} rows, single-threaded C restatement (oracle/oracle_c.c, "
                                                          f"{g1:.0f} GFLOP/s on a 1024^3 sgemm), {d1:.1f} s"}
        except (OSError, RuntimeError, subprocess.CalledProcessError) as e:  # no compiler on the box and no prebuilt library
            line["cpu_baseline"]["one_core"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if parity:
        parity["ok"] = all(v.get("ok", False) for v in parity.values() if isinstance(v, dict))
        line["parity"] = parity
    if rank == 0:
        emit(line)
        if parity and not parity["ok"]:
            print("PARITY FAILED: " + json.dumps(parity), file=sys.stderr)
    if job.dist is not None:
        job.dist.barrier()
        job.dist.destroy_process_group()
    if rank == 0 and parity and not parity["ok"]:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="mlp4096", choices=["mlp4096", "mlp8192", "mlp_small"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the data-parallel parity gate that runs before the timing")
    ap.add_argument("--no-also", action="store_true", help="headline only: do not time mlp8192 (BASELINE configs[4]) after mlp4096")
    ap.add_argument("--also-steps", type=int, default=3)
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce gradients after backward instead of during it")
    ap.add_argument("--fused", action="store_true", help="N > 1: one fused reduce-scatter + SGD + all-gather kernel over NVLink peer memory "
                                                         "instead of NCCL all-reduce + SGD (the default)")
    ap.add_argument("--fused-mono", action="store_true", help="N > 1, fused exchange: ONE kernel after backward instead of per-layer exchanges under backward")
    ap.add_argument("--no-forward-fusion", action="store_true", help="A/B: Linear and the activation behind it as two engine calls instead of one agrs_gemm_bias_act_f32")
    ap.add_argument("--act-warps", action="store_true", help="A/B: AGRS_TUNE_TC_ACT_WARPS=1, the pair GEMM's activation warps write act(Z) inside the forward GEMM (default: second pass)")
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
