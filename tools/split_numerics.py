"""CPU study of operand-splitting schemes for an f32-faithful GEMM on 16/19-bit tensor-core inputs (no GPU needed).
Every scheme's products are formed exactly (float64) from the rounded pieces and summed in float64, so what is measured is the
REPRESENTATION error of the scheme alone (the tensor core's truncating accumulator adds ~5e-6 on top).
  python tools/split_numerics.py [--k 4096] [--m 256] [--n 256]
slots = tf32-rate MMA slots per product (tf32 MMA = 1, bf16/fp16 MMA = 1/2)."""
import argparse
import json

import numpy as np


def trunc_tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rn_bf16(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def rn_fp16(x):
    return x.astype(np.float16).astype(np.float32)


def schemes(a, b):
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    out = {}
    # plain f32 GEMM (what the reference's CPU path computes)
    out["f32 sgemm (numpy)"] = ((a @ b).astype(np.float64), None)
    # single tf32 product
    ah, bh = trunc_tf32(a), trunc_tf32(b)
    out["tf32 x1"] = (ah.astype(np.float64) @ bh.astype(np.float64), 1.0)
    # 3xTF32: lo truncated to tf32 by the tensor core
    al, bl = trunc_tf32(a - ah), trunc_tf32(b - bh)
    f = lambda x: x.astype(np.float64)
    out["tf32 x3 (cross 0)"] = (f(ah) @ f(bh) + f(al) @ f(bh) + f(ah) @ f(bl), 3.0)
    # tf32 hi*hi + bf16 cross terms (the default)
    al16, bl16, ax16, bx16 = rn_bf16(a - ah), rn_bf16(b - bh), rn_bf16(a), rn_bf16(b)
    out["tf32 + 2 x bf16 (cross 1, default)"] = (f(ah) @ f(bh) + f(al16) @ f(bx16) + f(ax16) @ f(bl16), 2.0)
    # bf16 x3: a = a1 + a2
    a1, b1 = rn_bf16(a), rn_bf16(b)
    a2, b2 = rn_bf16(a - a1), rn_bf16(b - b1)
    out["bf16 x3"] = (f(a1) @ f(b1) + f(a1) @ f(b2) + f(a2) @ f(b1), 1.5)
    # fp16 x3 with a power-of-two scale per row of A / column of B (row max brought to [2^13, 2^14))
    def pow2_scale(mx):
        e = np.floor(np.log2(np.maximum(mx, np.float32(1e-37))))
        return np.exp2(13.0 - e).astype(np.float32)
    sa = pow2_scale(np.abs(a).max(axis=1, keepdims=True))
    sb = pow2_scale(np.abs(b).max(axis=0, keepdims=True))
    as_, bs_ = a * sa, b * sb
    h_a, h_b = rn_fp16(as_), rn_fp16(bs_)
    l_a, l_b = rn_fp16(as_ - h_a), rn_fp16(bs_ - h_b)
    c = f(h_a) @ f(h_b) + f(l_a) @ f(h_b) + f(h_a) @ f(l_b)
    out["fp16 x3, row/column scaled"] = (c / f(sa) / f(sb), 1.5)
    return out, a64 @ b64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=256)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--k", type=int, default=4096)
    a_ = ap.parse_args()
    rng = np.random.default_rng(0)
    cases = {
        "uniform(-1,1)": (rng.uniform(-1, 1, (a_.m, a_.k)), rng.uniform(-1, 1, (a_.k, a_.n))),
        "relu activations x kaiming weights": (np.maximum(rng.standard_normal((a_.m, a_.k)), 0), rng.uniform(-1, 1, (a_.k, a_.n)) / np.sqrt(a_.k)),
        "gradients 1e-7 x weights, rows spread over 2^20": (rng.standard_normal((a_.m, a_.k)) * 1e-7 * np.exp2(rng.integers(-20, 1, (a_.m, 1))),
                                                          rng.uniform(-1, 1, (a_.k, a_.n))),
    }
    for name, (a, b) in cases.items():
        a, b = a.astype(np.float32), b.astype(np.float32)
        res, truth = schemes(a, b)
        scale = np.abs(truth).max()
        rowscale = np.abs(truth).max(axis=1, keepdims=True)
        print(f"== {name}  (M={a_.m} N={a_.n} K={a_.k})")
        for k, (c, slots) in res.items():
            err = np.abs(c - truth)
            print(json.dumps({"scheme": k, "slots": slots, "max_err_over_max_C": float(err.max() / scale),
                              "max_err_over_row_max": float((err / rowscale).max())}))


if __name__ == "__main__":
    main()
