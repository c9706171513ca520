"""C3's Conv2D layer at the C-ABI level: the implicit-GEMM kernels against the literal lowering they replace, kernel by kernel
(CUDA-event timing on the library's stream, 50 launches each).  One JSON line per kernel."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import _capi  # noqa: E402


def main():
    ctx = _capi.Ctx(0)
    B, H, k, co = 1024, 28, 5, 6
    Ho = H - k + 1
    P, kk = Ho * Ho, k * k
    rng = np.random.default_rng(0)
    im = ctx.to_device(rng.uniform(0, 1, (B, 1, H, H)).astype(np.float32))
    filt = ctx.to_device(rng.uniform(-0.3, 0.3, (k, k, co)).astype(np.float32))
    bias = ctx.to_device(rng.uniform(-0.1, 0.1, (co,)).astype(np.float32))
    g = ctx.to_device(rng.standard_normal((B, P, co)).astype(np.float32))
    out, col, dxcol = ctx.empty(B * P * co), ctx.empty(B * P * kk), ctx.empty(B * P * kk)
    dx, dw, db = ctx.empty(B * H * H), ctx.empty(kk * co), ctx.empty(co)

    def timed(name, fn, bytes_moved, iters=50):
        for _ in range(3):
            fn()
        ctx.call("agrs_timer_begin")
        for _ in range(iters):
            fn()
        ms = C.c_float()
        ctx.call("agrs_timer_end", C.byref(ms))
        us = ms.value * 1e3 / iters
        print(json.dumps({"kernel": name, "us": round(us, 2), "algorithmic_MB": round(bytes_moved / 1e6, 2), "GBps": round(bytes_moved / us / 1e3, 1)}), flush=True)
        return us

    img_b, out_b, col_b = B * H * H * 4, B * P * co * 4, B * P * kk * 4
    t_impl = timed("implicit forward (agrs_conv2d_fwd_f32)", lambda: ctx.call("agrs_conv2d_fwd_f32", im.ptr, filt.ptr, bias.ptr, B, H, H, k, co, 1, 0, out.ptr), img_b + out_b)
    y = ctx.empty(B * P * co)
    timed("implicit forward + ReLU in one launch (agrs_conv2d_fwd_act_f32)",
          lambda: ctx.call("agrs_conv2d_fwd_act_f32", im.ptr, filt.ptr, bias.ptr, B, H, H, k, co, 1, 0, out.ptr, 1, y.ptr), img_b + 2 * out_b)
    timed("ReLU forward on the conv output, its own launch (agrs_relu_fwd_f32)", lambda: ctx.call("agrs_relu_fwd_f32", out.ptr, y.ptr, B * P * co), 2 * out_b)
    t_impl += timed("implicit backward (agrs_conv2d_bwd_f32: dx, dw, db)", lambda: ctx.call("agrs_conv2d_bwd_f32", im.ptr, filt.ptr, g.ptr, B, H, H, k, co, 1, 0, dx.ptr, dw.ptr, db.ptr),
                    img_b + out_b + img_b)
    t_lit = timed("literal: im2col", lambda: ctx.call("agrs_im2col_f32", im.ptr, B, 1, H, H, k, 1, 0, col.ptr), img_b + col_b)
    t_lit += timed("literal: col . K + bias", lambda: ctx.call("agrs_gemm_f32", 0, 0, B * P, co, kk, col.ptr, kk, filt.ptr, co, out.ptr, co, bias.ptr), col_b + out_b)
    t_lit += timed("literal: db = g.sum_axis(0)", lambda: ctx.call("agrs_sum_axis_f32", g.ptr, 1, B * P, co, db.ptr), out_b)
    t_lit += timed("literal: dxcol = g . K^T", lambda: ctx.call("agrs_gemm_f32", 0, 1, B * P, kk, co, g.ptr, co, filt.ptr, co, dxcol.ptr, kk, None), out_b + col_b)
    t_lit += timed("literal: col2im", lambda: ctx.call("agrs_col2im_f32", dxcol.ptr, B, 1, Ho, Ho, k, 1, dx.ptr), col_b + img_b)
    t_lit += timed("literal: dw = col^T . g", lambda: ctx.call("agrs_gemm_f32", 1, 0, kk, co, B * P, col.ptr, kk, g.ptr, co, dw.ptr, co, None), col_b + out_b)
    print(json.dumps({"layer_total_us": {"implicit": round(t_impl, 1), "literal": round(t_lit, 1)}}))
    ctx.close()


if __name__ == "__main__":
    main()
