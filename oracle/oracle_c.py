"""ctypes wrapper around oracle_c.c -- the single-threaded C restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.build() (to compile it) and bench.py's CPU legs.
The library is compiled in-tree into oracle/_build/liboracle_c.so (git-ignored, travels to the GPU box)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle_c.c")
OUT = os.path.join(HERE, "_build", "liboracle_c.so")

_lib = None


def build(force=False):
    """gcc -O2 -mavx2 -mfma, no -ffast-math (summation order is part of the restatement; a*b+c may contract to one FMA, as
    matrixmultiply's x86 kernel does when the CPU has FMA) -> oracle/_build/liboracle_c.so"""
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = OUT + f".{os.getpid()}.tmp"
    subprocess.run(["gcc", "-O2", "-mavx2", "-mfma", "-fno-fast-math", "-ffp-contract=fast", "-std=gnu11", "-shared", "-fPIC",
                    "-fvisibility=hidden", "-o", tmp, SRC, "-lm"], check=True)
    os.replace(tmp, OUT)
    return OUT


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        f32p, lng, f32 = ctypes.POINTER(ctypes.c_float), ctypes.c_long, ctypes.c_float
        sig = {
            "oc_relu_fwd": (None, [f32p, f32p, lng]),
            "oc_relu_bwd": (None, [f32p, f32p, f32p, lng]),
            "oc_sigmoid_fwd": (None, [f32p, f32p, lng]),
            "oc_sigmoid_bwd": (None, [f32p, f32p, f32p, lng]),
            "oc_mse_fwd": (f32, [f32p, f32p, f32p, lng]),
            "oc_mse_bwd": (None, [f32p, f32p, f32, lng, lng, f32p]),
            "oc_transpose": (None, [f32p, lng, lng, f32p]),
            "oc_sgemm": (None, [lng, lng, lng, f32p, lng, f32p, lng, f32p, lng]),
            "oc_add_rowvec": (None, [f32p, f32p, lng, lng]),
            "oc_colsum": (None, [f32p, lng, lng, f32p]),
            "oc_sgd": (None, [f32p, f32p, f32, lng, lng, f32p]),
            "oc_im2col": (None, [f32p, lng, lng, lng, lng, lng, lng, lng, f32p]),
            "oc_col2im": (None, [f32p, lng, lng, lng, lng, lng, lng, f32p]),
            "oc_mlp_step": (f32, [ctypes.c_int, ctypes.POINTER(lng), ctypes.POINTER(f32p), ctypes.POINTER(f32p), f32p, f32p, lng,
                                  ctypes.c_int, f32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def relu_fwd(x):
    x = _c(x); y = np.empty_like(x); lib().oc_relu_fwd(_p(x), _p(y), x.size); return y


def relu_bwd(x, g):
    x, g = _c(x), _c(g); y = np.empty_like(x); lib().oc_relu_bwd(_p(x), _p(g), _p(y), x.size); return y


def sigmoid_fwd(x):
    x = _c(x); y = np.empty_like(x); lib().oc_sigmoid_fwd(_p(x), _p(y), x.size); return y


def sigmoid_bwd(x, g):
    x, g = _c(x), _c(g); y = np.empty_like(x); lib().oc_sigmoid_bwd(_p(x), _p(g), _p(y), x.size); return y


def mse_fwd(p, t):
    p, t = _c(p), _c(t); s = np.empty_like(p)
    return np.float32(lib().oc_mse_fwd(_p(p), _p(t), _p(s), p.size))


def mse_bwd(p, t, upstream=1.0):
    p, t = _c(p), _c(t); out = np.empty_like(p)
    lib().oc_mse_bwd(_p(p), _p(t), float(upstream), p.shape[0], p.size, _p(out)); return out


def dot(a, b):
    a, b = _c(a), _c(b)
    (m, k), (k2, n) = a.shape, b.shape
    assert k == k2
    c = np.empty((m, n), np.float32)
    lib().oc_sgemm(m, n, k, _p(a), k, _p(b), n, _p(c), n)
    return c


def linear_fwd(x, w, b):
    y = dot(x, w); b = _c(b).reshape(-1); lib().oc_add_rowvec(_p(y), _p(b), y.shape[0], y.shape[1]); return y


def linear_bwd(x, w, g):
    """(dx, dw, db) with both transposes materialised as operators.rs:159-161 does"""
    x, w, g = _c(x), _c(w), _c(g)
    wt = np.empty((w.shape[1], w.shape[0]), np.float32); lib().oc_transpose(_p(w), w.shape[0], w.shape[1], _p(wt))
    xt = np.empty((x.shape[1], x.shape[0]), np.float32); lib().oc_transpose(_p(x), x.shape[0], x.shape[1], _p(xt))
    db = np.empty((1, g.shape[1]), np.float32); lib().oc_colsum(_p(g), g.shape[0], g.shape[1], _p(db))
    return dot(g, wt), dot(xt, g), db


def sgd_step(w, g, lr):
    """in place on w (must be a C-contiguous f32 array); g broadcasts by tiling as the shapes of the reference allow"""
    g = _c(g); tmp = np.empty(g.size, np.float32)
    assert w.size % g.size == 0
    lib().oc_sgd(_p(w), _p(g), float(lr), w.size, g.size, _p(tmp)); return w


def im2col(im, k, stride, pad):
    im = _c(im); B, C, H, W = im.shape
    ho, wo = (H - k + 2 * pad) // stride + 1, (W - k + 2 * pad) // stride + 1
    col = np.empty((B, ho * wo, C * k * k), np.float32)
    lib().oc_im2col(_p(im), B, C, H, W, k, stride, pad, _p(col)); return col


def col2im(col, c, h_out, w_out, stride, k):
    col = _c(col); B = col.shape[0]
    im = np.empty((B, c, (h_out - 1) * stride + k, (w_out - 1) * stride + k), np.float32)
    lib().oc_col2im(_p(col), B, c, h_out, w_out, k, stride, _p(im)); return im


def mlp_step(weights, biases, x, t, act, lr):
    """One full training iteration in C, updating `weights` / `biases` (lists of C-contiguous f32 arrays) in place."""
    L = len(weights)
    dims = [weights[0].shape[0]] + [w.shape[1] for w in weights]
    x, t = _c(x), _c(t)
    f32p = ctypes.POINTER(ctypes.c_float)
    Wp = (f32p * L)(*[_p(w) for w in weights])
    Bp = (f32p * L)(*[_p(b) for b in biases])
    D = (ctypes.c_long * (L + 1))(*dims)
    code = {None: 0, "relu": 1, "sigmoid": 2}[act]
    return float(lib().oc_mlp_step(L, D, Wp, Bp, _p(x), _p(t), x.shape[0], code, float(lr)))
