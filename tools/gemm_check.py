"""GPU diagnostic: run agrs_gemm_f32 on one path / operand-major combination and report the error
against an f64 product next to the error of a plain f32 product (the reference's own precision).
Each invocation is its own process so that a trapped kernel cannot poison the next case.

  python tools/gemm_check.py --path tc --ta 0 --tb 0 --m 256 --n 512 --k 128 [--bias] [--explicit-hi] [--bn 128]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autograd_rs_b200  # noqa: E402
from autograd_rs_b200 import _capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", default="tc", choices=["tc", "simt", "auto"])
    ap.add_argument("--ta", type=int, default=0)
    ap.add_argument("--tb", type=int, default=0)
    ap.add_argument("--m", type=int, default=256)
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--bias", action="store_true")
    ap.add_argument("--explicit-hi", action="store_true")
    ap.add_argument("--bn", type=int, default=256)
    ap.add_argument("--kchunk", type=int, default=1024)
    ap.add_argument("--pair", type=int, default=1, help="1: cta_group::2 pair kernel, 0: one-CTA kernel")
    ap.add_argument("--cross", type=int, default=0, help="pair kernel correction terms: 0 tf32, 1 bf16 (64B swizzle), 2 bf16 (no swizzle)")
    ap.add_argument("--iters", type=int, default=0, help="time this many launches after the check")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    M, N, K = a.m, a.n, a.k
    A = rng.uniform(-1, 1, (K, M) if a.ta else (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K) if a.tb else (K, N)).astype(np.float32)
    bias = rng.uniform(-1, 1, N).astype(np.float32) if a.bias else None
    opA = A.T if a.ta else A
    opB = B.T if a.tb else B
    truth = opA.astype(np.float64) @ opB.astype(np.float64)
    f32ref = (opA @ opB).astype(np.float64)
    if bias is not None:
        truth = truth + bias
        f32ref = f32ref + bias
    ctx = _capi.Ctx(0)
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_BN, a.bn)
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_EXPLICIT_HI, int(a.explicit_hi))
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_KCHUNK, a.kchunk)
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_2CTA, a.pair)
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_CROSS, a.cross)
    path = {"tc": _capi.GEMM_TCGEN05, "simt": _capi.GEMM_SIMT, "auto": _capi.GEMM_AUTO}[a.path]
    got = ctx.gemm(A, B, bool(a.ta), bool(a.tb), bias, path).astype(np.float64)
    scale = np.abs(truth).max()
    res = {"path": a.path, "ta": a.ta, "tb": a.tb, "M": M, "N": N, "K": K, "bias": a.bias, "explicit_hi": a.explicit_hi, "bn": a.bn, "kchunk": a.kchunk, "pair": a.pair, "cross": a.cross,
           "rel_err_vs_f64": float(np.abs(got - truth).max() / scale),
           "f32_numpy_rel_err_vs_f64": float(np.abs(f32ref - truth).max() / scale),
           "rel_err_vs_f32": float(np.abs(got - f32ref).max() / scale)}
    if res["rel_err_vs_f64"] > 1e-3:
        bad = np.argwhere(np.abs(got - truth) > 1e-3 * scale)
        res["n_bad"] = int(len(bad))
        res["first_bad"] = [[int(i), int(j), float(got[i, j]), float(truth[i, j])] for i, j in bad[:8]]
        res["bad_rows"] = sorted(set(int(i) for i, _ in bad))[:16]
        res["bad_cols"] = sorted(set(int(j) for _, j in bad))[:16]
    if a.iters:
        dA, dB = ctx.to_device(A), ctx.to_device(B)
        dC = ctx.empty(M * N)
        ctx.call("agrs_set_gemm_path", path)
        args = (int(a.ta), int(a.tb), M, N, K, dA.ptr, A.shape[1], dB.ptr, B.shape[1], dC.ptr, N, None)
        for _ in range(3):
            ctx.call("agrs_gemm_f32", *args)
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(a.iters):
            ctx.call("agrs_gemm_f32", *args)
        ctx.sync()
        dt = (time.perf_counter() - t0) / a.iters
        res["ms"] = dt * 1e3
        res["tflops"] = 2.0 * M * N * K / dt / 1e12
    print(json.dumps(res))


if __name__ == "__main__":
    main()
