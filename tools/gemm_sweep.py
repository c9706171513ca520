"""GEMM sweep on the GPU (BASELINE config C4 and the Linear shapes of C2/C5): for every shape x operand-major combination,
check a sample of output rows against an f64 NumPy product and time the kernel with CUDA events on its stream.
  python tools/gemm_sweep.py [--pair 0|1|both] [--sizes 256,512,...] [--linear] [--iters N] [--kchunk K]
One JSON line per case; the last line is a summary table."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import _capi  # noqa: E402


REUSE_SCALES = False
AUTO = False
GRAPH = False
SIMT = False
ANNOUNCE = False  # announce one magnitude per operand before every call (what the host engine does): the default mode then runs fp16 x3


def run_case(ctx, M, N, K, ta, tb, pair, iters, kchunk, rng, check_rows=48, bias=False, splitk=0):
    A = rng.uniform(-1, 1, (K, M) if ta else (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32)
    bv = rng.uniform(-1, 1, N).astype(np.float32) if bias else None
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_2CTA, pair)
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_KCHUNK, kchunk)
    ctx.call("agrs_set_tuning", _capi.TUNE_TC_SPLITK, splitk)
    ctx.call("agrs_set_gemm_path", _capi.GEMM_SIMT if SIMT else (_capi.GEMM_AUTO if AUTO else _capi.GEMM_TCGEN05))
    dA, dB, dC = ctx.to_device(A), ctx.to_device(B), ctx.empty(M * N)
    dbias = ctx.to_device(bv) if bias else None
    args = (int(ta), int(tb), M, N, K, dA.ptr, A.shape[1], dB.ptr, B.shape[1], dC.ptr, N, dbias.ptr if bias else None)
    if ANNOUNCE:
        sa, sb = ctx.empty(1), ctx.empty(1)
        ctx.call("agrs_absmax_f32", dA.ptr, A.size, sa.ptr)
        ctx.call("agrs_absmax_f32", dB.ptr, B.size, sb.ptr)
        plain = ctx.call

        def call(name, *a):  # the magnitudes belong to the tensors (computed once); every GEMM call is told about them
            if name == "agrs_gemm_f32":
                plain("agrs_gemm_operand_absmax", sa.ptr, sb.ptr)
            plain(name, *a)
    else:
        call = ctx.call
    call("agrs_gemm_f32", *args)
    got = ctx.to_host(dC, (M, N))
    rows = np.unique(np.concatenate([np.arange(min(M, 4)), np.arange(max(M - 4, 0), M), rng.integers(0, M, check_rows)]))
    opA = (A.T if ta else A)[rows].astype(np.float64)
    opB = (B.T if tb else B).astype(np.float64)
    truth = opA @ opB + (bv.astype(np.float64) if bias else 0.0)
    scale = float(np.abs(truth).max())
    err = float(np.abs(got[rows] - truth).max() / scale)
    f32 = ((A.T if ta else A)[rows] @ (B.T if tb else B)).astype(np.float64) + (bv if bias else 0.0)
    err32 = float(np.abs(f32 - truth).max() / scale)
    res = {"M": M, "N": N, "K": K, "ta": int(ta), "tb": int(tb), "pair": pair, "announced_magnitudes": ANNOUNCE, "kchunk": kchunk, "splitk": splitk, "bias": bias, "rel_err_vs_f64": err,
           "numpy_f32_rel_err_vs_f64": err32, "finite": bool(np.isfinite(got).all())}
    if iters:
        for _ in range(3):
            call("agrs_gemm_f32", *args)
        ctx.call("agrs_timer_begin")
        for _ in range(iters):
            call("agrs_gemm_f32", *args)
        import ctypes as C
        ms = C.c_float()
        ctx.call("agrs_timer_end", C.byref(ms))
        res["ms"] = ms.value / iters
        res["tflops"] = 2.0 * M * N * K / (res["ms"] * 1e-3) / 1e12
        if GRAPH:  # the same launches replayed from a CUDA graph: device time per launch without the host's per-call work
            g = C.c_void_p()
            ctx.call("agrs_capture_begin")
            try:
                for _ in range(20):
                    call("agrs_gemm_f32", *args)
            finally:
                ctx.call("agrs_capture_end", C.byref(g))
            ctx.call("agrs_graph_launch", g, 2)
            ctx.sync()
            ctx.call("agrs_timer_begin")
            ctx.call("agrs_graph_launch", g, 5)
            ctx.call("agrs_timer_end", C.byref(ms))
            ctx.L.agrs_graph_destroy(ctx.h, g)
            res["us_in_graph"] = ms.value * 1e3 / 100
            res["tflops_in_graph"] = 2.0 * M * N * K / (res["us_in_graph"] * 1e-6) / 1e12
        if REUSE_SCALES:  # cross mode 4 without its absmax passes: the magnitudes of the launches above are still in the workspace
            ctx.call("agrs_set_tuning", _capi.TUNE_TC_REUSE_SCALES, 1)
            ctx.call("agrs_timer_begin")
            for _ in range(iters):
                ctx.call("agrs_gemm_f32", *args)
            ctx.call("agrs_timer_end", C.byref(ms))
            ctx.call("agrs_set_tuning", _capi.TUNE_TC_REUSE_SCALES, 0)
            res["ms_reused_scales"] = ms.value / iters
            res["tflops_reused_scales"] = 2.0 * M * N * K / (res["ms_reused_scales"] * 1e-3) / 1e12
            res["rel_err_reused_scales_vs_first"] = float(np.abs(ctx.to_host(dC, (M, N)) - got).max() / max(np.abs(got).max(), 1e-30))
    ctx.call("agrs_set_gemm_path", _capi.GEMM_AUTO)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pair", default="both")
    ap.add_argument("--sizes", default="256,512,1024,2048,4096,8192")
    ap.add_argument("--linear", action="store_true", help="also the Linear fwd/dX/dW shapes of C2 (8192x4096x4096) and C5 (8192x8192x8192)")
    ap.add_argument("--linear-c2", action="store_true", help="only the three Linear shapes of C2")
    ap.add_argument("--reuse-scales", action="store_true", help="cross mode 4: also time the kernel without its absmax passes")
    ap.add_argument("--announce", action="store_true", help="announce one magnitude per operand before every call, as the host engine does")
    ap.add_argument("--auto", action="store_true", help="let AUTO path selection choose the kernel instead of forcing the tcgen05 path")
    ap.add_argument("--simt", action="store_true", help="force the FP32 SIMT kernel (what small problems fall back to)")
    ap.add_argument("--graph", action="store_true", help="also time 20 launches captured into a CUDA graph (device time per launch, no host work)")
    ap.add_argument("--odd", action="store_true", help="ragged shapes (partial tiles in M, N and K)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--kchunk", type=int, default=1024)
    ap.add_argument("--splitk", default="0", help="comma list of forced K splits for the pair kernel (0 = automatic)")
    ap.add_argument("--majors", default="00,01,10", help="ta tb pairs for the square sizes: 00 = X.W, 01 = G.W^T, 10 = X^T.G, 11")
    a = ap.parse_args()
    global REUSE_SCALES, ANNOUNCE, AUTO, GRAPH, SIMT
    GRAPH = a.graph
    SIMT = a.simt
    REUSE_SCALES = a.reuse_scales
    ANNOUNCE = a.announce
    AUTO = a.auto
    pairs = [0, 1] if a.pair == "both" else [int(a.pair)]
    ctx = _capi.Ctx(0)
    rng = np.random.default_rng(0)
    cases = []
    for n in [int(s) for s in a.sizes.split(",") if s]:
        for ta, tb in [(int(m[0]), int(m[1])) for m in a.majors.split(",")]:  # fwd X.W (NN), dX = G.W^T (NT), dW = X^T.G (TN)
            cases.append((n, n, n, ta, tb, False))
    if a.linear or a.linear_c2:
        for (m, n, k) in (((8192, 4096, 4096),) if a.linear_c2 else ((8192, 4096, 4096), (8192, 8192, 8192))):
            cases += [(m, n, k, 0, 0, True), (m, k, n, 0, 1, False), (k, n, m, 1, 0, False)]
    if a.odd:
        cases += [(136, 72, 40, 0, 0, True), (200, 300, 100, 1, 1, False), (1000, 520, 3000, 1, 0, True), (257, 264, 36, 0, 1, False),
                  (129, 8, 8, 0, 0, False), (384, 1000, 520, 0, 1, True), (2048, 2048, 2048, 1, 1, False), (132, 260, 4100, 1, 1, True), (4100, 132, 260, 1, 0, False)]
    out = []
    for (M, N, K, ta, tb, bias) in cases:
        for p in pairs:
            for sk in ([int(x) for x in a.splitk.split(",")] if p else [0]):
                r = run_case(ctx, M, N, K, ta, tb, p, a.iters, a.kchunk, rng, bias=bias, splitk=sk)
                print(json.dumps(r), flush=True)
                out.append(r)
    bad = [r for r in out if not r["finite"] or r["rel_err_vs_f64"] > 2e-5]
    print(json.dumps({"cases": len(out), "bad": len(bad), "max_rel_err": max(r["rel_err_vs_f64"] for r in out)}))
    ctx.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
