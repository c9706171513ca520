"""HBM roofline of the elementwise / reduction / im2col kernels: algorithmic bytes / CUDA-event time, per kernel,
at the sizes of BASELINE configs C2 (8192 x 4096 activations, 4096^2 weights) and C3 (1024 x 1 x 28 x 28 images).
Buffers are rotated through a pool larger than the 126 MB L2 so every launch reads from HBM.
  python tools/eltwise_bench.py [--iters N]   -> one JSON line per kernel, then a summary"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import _capi  # noqa: E402

PEAK = 6538.3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    ctx = _capi.Ctx(0)
    n = 8192 * 4096          # one C2 activation: 134 MB
    nset = 6                 # rotate 6 sets of buffers: > 800 MB per operand role, nothing survives in L2
    bufs = {k: [ctx.empty(n) for _ in range(nset)] for k in ("a", "b", "c")}
    for k in bufs:
        for b in bufs[k]:
            ctx.call("agrs_fill_f32", b.ptr, 0.5 if k != "b" else 0.25, n)
    scal = ctx.empty(16)
    ctx.call("agrs_fill_f32", scal.ptr, 1.0, 16)
    small = ctx.empty(8192)
    out = []

    def timed(name, bytes_per_launch, fn):
        for i in range(3):
            fn(i % nset)
        ctx.call("agrs_timer_begin")
        for i in range(a.iters):
            fn(i % nset)
        ms = C.c_float()
        ctx.call("agrs_timer_end", C.byref(ms))
        t = ms.value / a.iters
        gbs = bytes_per_launch / (t * 1e-3) / 1e9
        r = {"kernel": name, "bytes": bytes_per_launch, "ms": t, "GBps": gbs, "frac_of_measured_hbm": gbs / PEAK}
        print(json.dumps(r), flush=True)
        out.append(r)

    A, B, Cc = bufs["a"], bufs["b"], bufs["c"]
    timed("relu_fwd", 8 * n, lambda i: ctx.call("agrs_relu_fwd_f32", A[i].ptr, Cc[i].ptr, n))
    timed("relu_bwd", 12 * n, lambda i: ctx.call("agrs_relu_bwd_f32", A[i].ptr, B[i].ptr, Cc[i].ptr, n))
    timed("sigmoid_fwd", 8 * n, lambda i: ctx.call("agrs_sigmoid_fwd_f32", A[i].ptr, Cc[i].ptr, n))
    timed("sigmoid_bwd", 12 * n, lambda i: ctx.call("agrs_sigmoid_bwd_f32", A[i].ptr, B[i].ptr, Cc[i].ptr, n))
    timed("mse_fwd", 8 * n, lambda i: ctx.call("agrs_mse_fwd_f32", A[i].ptr, B[i].ptr, n, scal.ptr))
    timed("mse_bwd", 12 * n, lambda i: ctx.call("agrs_mse_bwd_f32", A[i].ptr, B[i].ptr, scal.ptr, 8192.0, n, Cc[i].ptr))
    timed("add_inplace (grad +=)", 12 * n, lambda i: ctx.call("agrs_binary_f32", _capi.ADD, A[i].ptr, A[i].ptr, B[i].ptr, n, n, 0))
    timed("bias_add_rowvec", 8 * n, lambda i: ctx.call("agrs_binary_f32", _capi.ADD, A[i].ptr, A[i].ptr, small.ptr, n, 4096, 0))
    timed("sgd_step", 12 * n, lambda i: ctx.call("agrs_sgd_step_f32", A[i].ptr, B[i].ptr, 1e-6, n, n))
    timed("colsum (db, [8192,4096] -> [4096])", 4 * n, lambda i: ctx.call("agrs_sum_axis_f32", A[i].ptr, 1, 8192, 4096, small.ptr))
    # the fused backward + column-sum kernels of the training step (f2): same bytes as the plain backward, db for free
    timed("relu_bwd + colsum", 12 * n, lambda i: ctx.call("agrs_relu_bwd_colsum_f32", A[i].ptr, B[i].ptr, Cc[i].ptr, 8192, 4096, small.ptr))
    timed("sigmoid_bwd + colsum", 12 * n, lambda i: ctx.call("agrs_sigmoid_bwd_colsum_f32", A[i].ptr, B[i].ptr, Cc[i].ptr, 8192, 4096, small.ptr))
    timed("mse_bwd + colsum", 12 * n, lambda i: ctx.call("agrs_mse_bwd_colsum_f32", A[i].ptr, B[i].ptr, scal.ptr, 8192.0, 8192, 4096, Cc[i].ptr, small.ptr))
    timed("sum", 4 * n, lambda i: ctx.call("agrs_sum_f32", A[i].ptr, n, scal.ptr))
    timed("fill_zero (memset)", 4 * n, lambda i: ctx.call("agrs_fill_f32", Cc[i].ptr, 0.0, n))
    timed("fill_one", 4 * n, lambda i: ctx.call("agrs_fill_f32", Cc[i].ptr, 1.0, n))
    timed("scale", 8 * n, lambda i: ctx.call("agrs_scalar_f32", _capi.SMUL, Cc[i].ptr, A[i].ptr, 1.5, n))
    timed("copy (agrs_copy, D2D memcpy)", 8 * n, lambda i: ctx.call("agrs_copy", Cc[i].ptr, A[i].ptr, 4 * n))
    timed("transpose 8192x4096", 8 * n, lambda i: ctx.call("agrs_transpose_f32", A[i].ptr, 8192, 4096, Cc[i].ptr))
    # C3: im[1024,1,28,28] -> col[1024,576,25]; scaled x8 in batch so the launch is not latency-bound (8192 images)
    Bt, k, H = 8192, 5, 28
    P = (H - k + 1) ** 2
    im_n, col_n = Bt * H * H, Bt * P * k * k
    assert col_n <= n * 4
    colbuf = [ctx.empty(col_n) for _ in range(2)]
    timed("im2col 8192x1x28x28 k5", 4 * (im_n + col_n), lambda i: ctx.call("agrs_im2col_f32", A[i].ptr, Bt, 1, H, H, k, 1, 0, colbuf[i % 2].ptr))
    timed("col2im 8192x1x28x28 k5", 4 * (im_n + col_n), lambda i: ctx.call("agrs_col2im_f32", colbuf[i % 2].ptr, Bt, 1, H - k + 1, H - k + 1, k, 1, A[i].ptr))
    B3 = 1024
    timed("im2col 1024x1x28x28 k5 (C3 size)", 4 * (B3 * H * H + B3 * P * k * k), lambda i: ctx.call("agrs_im2col_f32", A[i].ptr, B3, 1, H, H, k, 1, 0, colbuf[i % 2].ptr))
    print(json.dumps({"hbm_peak_GBps_measured": PEAK, "kernels": len(out), "min_frac": min(r["frac_of_measured_hbm"] for r in out),
                      "below_70pct": [r["kernel"] for r in out if r["frac_of_measured_hbm"] < 0.7]}))
    ctx.close()


if __name__ == "__main__":
    main()
