"""GPU parity tests of the kernel-level C ABI (libagrs_b200.so) against the oracle.
Bit-exact where the arithmetic is elementwise IEEE f32; tolerance (written per test) where a
summation order or a transcendental differs.  Run with: pytest -m gpu (under gpurun)."""
import ctypes

import numpy as np
import pytest

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import _capi as K
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 3, 4, 5, 1023, 4096, (1 << 20) + 3]


@pytest.fixture(scope="module")
def ctx():
    c = K.Ctx(0)
    yield c
    c.close()


def rnd(rng, *shape):
    return rng.uniform(-2, 2, shape).astype(np.float32)


def unary(ctx, name, x):
    dx = ctx.to_device(x)
    dy = ctx.empty(max(x.size, 1))
    ctx.call(name, dx.ptr, dy.ptr, x.size)
    return ctx.to_host(dy, x.shape)


# ------------------------------------------------------------------ reference KATs through the C ABI
def test_kat_relu_on_mat(ctx, kats):
    k = kats["relu_on_mat"]
    x = np.array(k["x"], np.float32)
    assert np.array_equal(unary(ctx, "agrs_relu_fwd_f32", x), np.array(k["y"], np.float32))


def test_kat_linear(ctx, kats):
    k = kats["linear"]
    y = ctx.gemm(np.array(k["x"]), np.array(k["w"]), bias=np.array(k["b"]).reshape(-1))
    assert np.array_equal(y, np.array(k["y"], np.float32))


@pytest.mark.parametrize("name", ["im2col_nopad", "im2col_pad"])
def test_kat_im2col(ctx, kats, name):
    k = kats[name]
    im = np.arange(*k["im_range"], dtype=np.float32).reshape(k["im_shape"])
    B, C, H, W = im.shape
    dim = ctx.to_device(im)
    n = int(np.prod(k["col_shape"]))
    dcol = ctx.empty(n)
    ctx.call("agrs_im2col_f32", dim.ptr, B, C, H, W, k["k"], k["stride"], k["pad"], dcol.ptr)
    assert np.array_equal(ctx.to_host(dcol, k["col_shape"]), np.array(k["col"], np.float32).reshape(k["col_shape"]))


def test_kat_conv2d_ones(ctx, kats):
    k = kats["conv2d_ones"]
    im = np.ones(k["im_shape"], np.float32)
    filt = np.ones(k["filt_shape"], np.float32)
    bias = np.ones(k["bias_shape"], np.float32)
    B, C, H, W = im.shape
    kk, co = k["k"] * k["k"], k["c_out"]
    P = 9
    dim, dfilt = ctx.to_device(im), ctx.to_device(filt)
    dcol, dkm = ctx.empty(B * P * C * kk), ctx.empty(C * kk * co)
    ctx.call("agrs_im2col_f32", dim.ptr, B, C, H, W, k["k"], 1, 0, dcol.ptr)
    ctx.call("agrs_broadcast_filter_f32", dfilt.ptr, kk, co, C, dkm.ptr)
    col = ctx.to_host(dcol, (B * P, C * kk))
    km = ctx.to_host(dkm, (C * kk, co))
    assert np.array_equal(km, O.conv_kernel_matrix(filt, C))
    out = ctx.gemm(col, km, bias=bias).reshape(k["out_shape"])
    assert np.all(out == k["out_value"])


def test_kat_relu_backward_single_node(ctx, kats):
    k = kats["single_node_graph"]
    x = np.array(k["x"], np.float32)
    g = np.ones_like(x)
    dx, dg, do = ctx.to_device(x), ctx.to_device(g), ctx.empty(x.size)
    ctx.call("agrs_relu_bwd_f32", dx.ptr, dg.ptr, do.ptr, x.size)
    assert np.array_equal(ctx.to_host(do, x.shape), np.array(k["x_grad"], np.float32))


# ------------------------------------------------------------------ elementwise: bit-exact
@pytest.mark.parametrize("n", SIZES)
def test_relu_fwd_bwd_exact(ctx, n):
    rng = np.random.default_rng(n)
    x = rnd(rng, n)
    if n > 4:
        x[:5] = [0.0, -0.0, np.nan, np.inf, -np.inf]
    g = rnd(rng, n)
    y = unary(ctx, "agrs_relu_fwd_f32", x)
    want = O.relu_fwd(x)
    assert np.array_equal(y, want) and np.array_equal(np.signbit(y), np.signbit(want))
    dx, dg, do = ctx.to_device(x), ctx.to_device(g), ctx.empty(max(n, 1))
    ctx.call("agrs_relu_bwd_f32", dx.ptr, dg.ptr, do.ptr, n)
    assert np.array_equal(ctx.to_host(do, (n,)), O.relu_bwd(x, g))


@pytest.mark.parametrize("n", SIZES)
def test_sigmoid(ctx, n):
    rng = np.random.default_rng(100 + n)
    x = (rnd(rng, n) * 5).astype(np.float32)
    g = rnd(rng, n)
    y = unary(ctx, "agrs_sigmoid_fwd_f32", x)
    # tolerance: CUDA expf is <= 2 ulp, Rust's f32::exp <= 1 ulp: 4e-7 relative on the output
    assert np.allclose(y, O.sigmoid_fwd(x), rtol=4e-7, atol=1e-37)
    dx, dg, do = ctx.to_device(x), ctx.to_device(g), ctx.empty(max(n, 1))
    ctx.call("agrs_sigmoid_bwd_f32", dx.ptr, dg.ptr, do.ptr, n)
    got = ctx.to_host(do, (n,))
    want64 = O.sigmoid_bwd(x.astype(np.float64), g.astype(np.float64))
    # The literal formula ((1 - s) * s) * g (operators.rs:219-223) cancels when s -> 1: an error of e ulp(1) in the f32
    # sigmoid becomes e * 6e-8 * |s * g| ABSOLUTE in dx, in the reference's ndarray path exactly as here.  Bound:
    # 2e-6 relative (roundings of the three products) + 3 ulp(1) * |g| absolute (expf <= 2 ulp, divide, 1 - s).
    assert np.all(np.abs(got - want64) <= 2e-6 * np.abs(want64) + 3 * 5.97e-8 * np.abs(g))
    # and against the f32 literal restatement (numpy exp vs CUDA expf differ by <= 2 ulp): same bound
    want32 = O.sigmoid_bwd(x, g)
    assert np.all(np.abs(got - want32) <= 2e-6 * np.abs(want32) + 4 * 5.97e-8 * np.abs(g))


@pytest.mark.parametrize("n", SIZES)
def test_binary_scalar_exact(ctx, n):
    rng = np.random.default_rng(200 + n)
    a, b = rnd(rng, n), rnd(rng, n)
    da, db, do = ctx.to_device(a), ctx.to_device(b), ctx.empty(max(n, 1))
    for op, f in ((K.ADD, np.add), (K.SUB, np.subtract), (K.MUL, np.multiply)):
        ctx.call("agrs_binary_f32", op, do.ptr, da.ptr, db.ptr, n, n, K.BCAST_TRAILING)
        assert np.array_equal(ctx.to_host(do, (n,)), f(a, b))
    s = np.float32(1.7)
    for op, want in ((K.SMUL, a * s), (K.SDIV, a / s), (K.SRSUB, s - a), (K.SADD, a + s)):
        ctx.call("agrs_scalar_f32", op, do.ptr, da.ptr, float(s), n)
        assert np.array_equal(ctx.to_host(do, (n,)), want)
    ctx.call("agrs_neg_f32", do.ptr, da.ptr, n)
    assert np.array_equal(ctx.to_host(do, (n,)), -a)
    ctx.call("agrs_powi_f32", do.ptr, da.ptr, 2, n)
    assert np.array_equal(ctx.to_host(do, (n,)), a * a)
    ctx.call("agrs_powi_f32", do.ptr, da.ptr, 3, n)
    assert np.allclose(ctx.to_host(do, (n,)), a.astype(np.float64) ** 3, rtol=3e-7)
    ctx.call("agrs_fill_f32", do.ptr, 2.5, n)
    assert np.all(ctx.to_host(do, (n,)) == 2.5)
    ctx.call("agrs_fill_f32", do.ptr, 0.0, n)
    assert np.all(ctx.to_host(do, (n,)) == 0.0)


@pytest.mark.parametrize("rows,cols", [(1, 1), (7, 5), (33, 8), (513, 1024), (64, 3)])
def test_broadcast_exact(ctx, rows, cols):
    rng = np.random.default_rng(rows * 31 + cols)
    a = rnd(rng, rows, cols)
    row, colv, sc = rnd(rng, cols), rnd(rng, rows, 1), rnd(rng, 1, 1)
    da, do = ctx.to_device(a), ctx.empty(a.size)
    n = a.size
    d = ctx.to_device(row)
    ctx.call("agrs_binary_f32", K.ADD, do.ptr, da.ptr, d.ptr, n, cols, K.BCAST_TRAILING)  # bias add / grad += [O]
    assert np.array_equal(ctx.to_host(do, a.shape), a + row)
    d = ctx.to_device(colv)
    ctx.call("agrs_binary_f32", K.MUL, do.ptr, da.ptr, d.ptr, n, rows, K.BCAST_LEADING)
    assert np.array_equal(ctx.to_host(do, a.shape), a * colv)
    d = ctx.to_device(sc)
    ctx.call("agrs_binary_f32", K.MUL, do.ptr, da.ptr, d.ptr, n, 1, K.BCAST_TRAILING)     # * grad[1,1]
    assert np.array_equal(ctx.to_host(do, a.shape), a * sc)
    # in place (out aliases a), and a shape that cannot broadcast
    d = ctx.to_device(row)
    ctx.call("agrs_binary_f32", K.SUB, da.ptr, da.ptr, d.ptr, n, cols, K.BCAST_TRAILING)
    assert np.array_equal(ctx.to_host(da, a.shape), a - row)
    if n % 7:
        with pytest.raises(K.AgrsError) as e:
            ctx.call("agrs_binary_f32", K.ADD, do.ptr, da.ptr, d.ptr, n, 7, K.BCAST_TRAILING)
        assert e.value.code == 3


@pytest.mark.parametrize("n,ng", [(1, 1), (12, 12), (12, 4), (4096 * 33, 4096), (1 << 20, 1 << 20), (1000003, 1000003)])
def test_sgd_step_exact(ctx, n, ng):
    rng = np.random.default_rng(n)
    w, g = rnd(rng, n), rnd(rng, ng)
    dw, dg = ctx.to_device(w), ctx.to_device(g)
    ctx.call("agrs_sgd_step_f32", dw.ptr, dg.ptr, 1e-3, n, ng)
    want = w.reshape(-1, ng).copy()
    O.sgd_step(want, g, 1e-3)  # two roundings: tmp = g*lr; w -= tmp (optim.rs:31-33)
    assert np.array_equal(ctx.to_host(dw, (n,)), want.reshape(-1))


def test_sgd_step_multi_matches_the_single_tensor_kernel(ctx):
    """agrs_sgd_step_multi_f32: SGD::step's loop over the parameter list as one launch per 16 parameters -- the same bits as
    agrs_sgd_step_f32 per tensor (two roundings), for aligned, ragged, tiny and empty tensors, and the magnitude words it leaves
    are what agrs_absmax_f32 finds on the updated tensors"""
    rng = np.random.default_rng(321)
    sizes = [1, 3, 4, 17, 256, 1000003, 0, 4096 * 33, 5, 10, 6 * 25, 3456 * 10, 1 << 20, 12, 7, 64, 33, 2, 100]   # 19 tensors: two launches
    ws = [rnd(rng, n) for n in sizes]
    gs = [rnd(rng, n) for n in sizes]
    dw = [ctx.to_device(w) for w in ws]
    dg = [ctx.to_device(g) for g in gs]
    mx = [ctx.empty(1) if n >= 256 else None for n in sizes]
    for m in mx:
        if m is not None:
            ctx.call("agrs_fill_f32", m.ptr, ctypes.c_float(3.0), 1)   # stale contents must not survive
    cnt = len(sizes)
    PW, PS = ctypes.c_void_p * cnt, ctypes.c_size_t * cnt
    ctx.call("agrs_sgd_step_multi_f32", cnt, PW(*[b.ptr for b in dw]), PW(*[b.ptr for b in dg]), PS(*sizes),
             PW(*[m.ptr if m is not None else None for m in mx]), ctypes.c_float(1e-3))
    for n, w, g, b, m in zip(sizes, ws, gs, dw, mx):
        if n == 0:
            continue
        want = w.reshape(-1, n).copy()
        O.sgd_step(want, g, 1e-3)
        got = ctx.to_host(b, (n,))
        assert np.array_equal(got, want.reshape(-1)), n
        if m is not None:
            bits = ctx.to_host(m, (1,)).view(np.uint32)[0]
            assert bits == np.abs(got).max().view(np.uint32), n
    with pytest.raises(K.AgrsError):
        ctx.call("agrs_sgd_step_multi_f32", 2, PW(dw[0].ptr, None), PW(dg[0].ptr, dg[1].ptr), PS(1, 3), None, ctypes.c_float(1e-3))


def test_async_download_and_events_at_the_c_abi(ctx):
    """agrs_download_async / agrs_event_wait / agrs_upload_prefetch_event / agrs_wait_event: the calls of INTEGRATION.md section 6"""
    n = 1 << 20
    host_in, host_out = ctypes.c_void_p(), ctypes.c_void_p()
    ctx.call("agrs_host_alloc", n * 4, ctypes.byref(host_in))
    ctx.call("agrs_host_alloc", n * 4, ctypes.byref(host_out))
    a = np.frombuffer((ctypes.c_float * n).from_address(host_in.value), dtype=np.float32)
    b = np.frombuffer((ctypes.c_float * n).from_address(host_out.value), dtype=np.float32)
    a[:] = np.arange(n, dtype=np.float32)
    b[:] = -1
    landed, read = ctypes.c_void_p(), ctypes.c_void_p()
    ctx.call("agrs_event_create", ctypes.byref(landed))
    ctx.call("agrs_event_create", ctypes.byref(read))
    ctx.call("agrs_event_wait", read)                      # never recorded: returns at once
    dev, dev2 = ctx.empty(n), ctx.empty(n)
    ctx.call("agrs_upload_prefetch_event", dev.ptr, host_in, n * 4, landed)
    ctx.call("agrs_wait_event", landed)                    # the stream, not the host, waits for the copy
    ctx.call("agrs_scalar_f32", K.SMUL, dev2.ptr, dev.ptr, ctypes.c_float(2.0), n)
    ctx.call("agrs_download_async", host_out, dev2.ptr, n * 4, read)
    ctx.call("agrs_event_wait", read)
    assert np.array_equal(b, 2 * a)
    with pytest.raises(K.AgrsError):
        ctx.call("agrs_download_async", host_out, dev2.ptr, n * 4, None)
    ctx.call("agrs_event_destroy", landed)
    ctx.call("agrs_event_destroy", read)
    ctx.call("agrs_host_free", host_in)
    ctx.call("agrs_host_free", host_out)


def test_pool_trim_returns_memory(ctx):
    """agrs_pool_trim: unused blocks go back to the driver, high-water marks restart"""
    def stats():
        v = [ctypes.c_size_t() for _ in range(4)]
        ctx.call("agrs_pool_stats", *[ctypes.byref(x) for x in v])
        return [int(x.value) for x in v]
    ctx.call("agrs_pool_reserve", 512 << 20)
    assert stats()[0] >= 512 << 20
    ctx.call("agrs_pool_trim")
    reserved, used, reserved_high, used_high = stats()
    assert reserved < 512 << 20 and reserved_high <= reserved + (64 << 20) and used_high <= used + (64 << 20)
    buf = ctx.empty(1 << 20)                               # the pool still serves allocations afterwards
    ctx.call("agrs_fill_f32", buf.ptr, ctypes.c_float(1.5), 1 << 20)
    assert np.all(ctx.to_host(buf, (1 << 20,)) == 1.5)


@pytest.mark.parametrize("B,D", [(1, 1), (5, 3), (2000, 1), (256, 4096), (1023, 77)])
def test_mse_fwd_bwd(ctx, B, D):
    rng = np.random.default_rng(B * 7 + D)
    p, t = rnd(rng, B, D), rnd(rng, B, D)
    dp, dt, dl = ctx.to_device(p), ctx.to_device(t), ctx.empty(1)
    ctx.call("agrs_mse_fwd_f32", dp.ptr, dt.ptr, p.size, dl.ptr)
    loss = ctx.to_host(dl, (1, 1))
    want64 = O.mse_fwd(p.astype(np.float64), t.astype(np.float64))
    # tolerance: tree vs sequential f32 summation, both within 1e-6 of the f64 value
    assert abs(loss[0, 0] - want64[0, 0]) <= 1e-6 * want64[0, 0]
    up = np.array([[1.0]], np.float32)
    dup, dg = ctx.to_device(up), ctx.empty(p.size)
    ctx.call("agrs_mse_bwd_f32", dp.ptr, dt.ptr, dup.ptr, float(B), p.size, dg.ptr)
    assert np.array_equal(ctx.to_host(dg, p.shape), O.mse_bwd(p, t, up))  # IEEE sub/mul/div/mul: bit-exact
    up = np.array([[0.37]], np.float32)
    dup = ctx.to_device(up)
    ctx.call("agrs_mse_bwd_f32", dp.ptr, dt.ptr, dup.ptr, float(B), p.size, dg.ptr)
    assert np.array_equal(ctx.to_host(dg, p.shape), O.mse_bwd(p, t, up))


@pytest.mark.parametrize("n", [1, 3, 1000, 4097, (1 << 22) + 5])
def test_sum_mean(ctx, n):
    rng = np.random.default_rng(n)
    x = rnd(rng, n)
    dx, do = ctx.to_device(x), ctx.empty(1)
    ctx.call("agrs_sum_f32", dx.ptr, n, do.ptr)
    s = float(ctx.to_host(do, (1,))[0])
    ref = float(x.astype(np.float64).sum())
    tol = 1e-6 * float(np.abs(x).astype(np.float64).sum()) + 1e-30  # tolerance: relative to sum|x|
    assert abs(s - ref) <= tol
    ctx.call("agrs_mean_f32", dx.ptr, n, do.ptr)
    assert abs(float(ctx.to_host(do, (1,))[0]) - ref / n) <= tol / n + 1e-12
    # deterministic: same bits on a second launch
    ctx.call("agrs_sum_f32", dx.ptr, n, do.ptr)
    assert float(ctx.to_host(do, (1,))[0]) == s


@pytest.mark.parametrize("outer,length,inner", [(1, 1, 1), (1, 8192, 4096), (1, 2000, 1), (1, 1000, 10), (1, 777, 6),
                                                (3, 50, 129), (5000, 3, 1), (37, 200, 1), (9 * 25, 3, 1), (1, 100000, 8), (1, 589824, 6), (1, 5000, 31),
                                                (1, 600, 3), (1, 512, 25)])
def test_sum_axis(ctx, outer, length, inner):
    rng = np.random.default_rng(outer + length + inner)
    x = rnd(rng, outer, length, inner)
    dx, do = ctx.to_device(x), ctx.empty(outer * inner)
    ctx.call("agrs_sum_axis_f32", dx.ptr, outer, length, inner, do.ptr)
    got = ctx.to_host(do, (outer, inner))
    ref = x.astype(np.float64).sum(axis=1)
    bound = np.abs(x).astype(np.float64).sum(axis=1)
    assert np.all(np.abs(got - ref) <= 1e-6 * bound + 1e-30)  # tolerance relative to sum|x| per output


def test_transpose_equal(ctx):
    rng = np.random.default_rng(5)
    for r, c in [(1, 1), (3, 5), (64, 64), (1000, 37)]:
        x = rnd(rng, r, c)
        dx, do = ctx.to_device(x), ctx.empty(x.size)
        ctx.call("agrs_transpose_f32", dx.ptr, r, c, do.ptr)
        assert np.array_equal(ctx.to_host(do, (c, r)), x.T)
        import ctypes
        eq = ctypes.c_int(-1)
        ctx.call("agrs_equal_f32", dx.ptr, dx.ptr, x.size, ctypes.byref(eq))
        assert eq.value == 1
        if x.size > 1:
            ctx.call("agrs_equal_f32", dx.ptr, do.ptr, x.size, ctypes.byref(eq))
            assert eq.value == int(np.array_equal(x.reshape(-1), x.T.reshape(-1)))


# ------------------------------------------------------------------ conv lowering
@pytest.mark.parametrize("B,C,H,W,k,s,p", [(1, 1, 5, 5, 3, 1, 0), (2, 3, 8, 7, 3, 2, 1), (4, 1, 28, 28, 5, 1, 0), (3, 2, 9, 9, 2, 2, 0),
                                           (2, 2, 6, 6, 3, 1, 2), (1, 3, 24, 24, 3, 1, 0), (300, 1, 28, 28, 5, 1, 0), (5, 2, 11, 13, 4, 3, 2),
                                           (2, 1, 230, 230, 3, 1, 0)])  # the last sample does not fit in shared memory: generic kernels
def test_im2col_col2im(ctx, B, C, H, W, k, s, p):
    rng = np.random.default_rng(B + C + H + k)
    im = rnd(rng, B, C, H, W)
    want = O.im2col(im, k, s, p) if im.size < 20000 else O.im2col_fast(im, k, s, p)
    ho, wo = O.conv_out_hw(H, W, k, s, p)
    dim, dcol = ctx.to_device(im), ctx.empty(want.size)
    ctx.call("agrs_im2col_f32", dim.ptr, B, C, H, W, k, s, p, dcol.ptr)
    assert np.array_equal(ctx.to_host(dcol, want.shape), want)
    col = rnd(rng, *want.shape)
    want_im = O.col2im(col, ho, wo, s, k)
    dcol, dout = ctx.to_device(col), ctx.empty(want_im.size)
    ctx.call("agrs_col2im_f32", dcol.ptr, B, C, ho, wo, k, s, dout.ptr)
    got = ctx.to_host(dout, want_im.shape)
    assert np.array_equal(got, want_im)  # same patch order as the reference's overlap-add: bit-exact


def test_im2col_bad_geometry(ctx):
    d = ctx.empty(16)
    with pytest.raises(K.AgrsError) as e:
        ctx.call("agrs_im2col_f32", d.ptr, 1, 1, 2, 2, 5, 1, 0, d.ptr)
    assert e.value.code == 3


# ------------------------------------------------------------------ SIMT GEMM
def gemm_case(ctx, rng, M, N, Kd, ta, tb, bias, path, tol):
    A = rnd(rng, *((Kd, M) if ta else (M, Kd)))
    B = rnd(rng, *((N, Kd) if tb else (Kd, N)))
    bv = rnd(rng, N) if bias else None
    got = ctx.gemm(A, B, ta, tb, bv, path)
    opA, opB = (A.T if ta else A).astype(np.float64), (B.T if tb else B).astype(np.float64)
    truth = opA @ opB + (bv if bias else 0)
    err = O.rel_err(got, truth) if truth.size else 0.0
    assert err <= tol, (M, N, Kd, ta, tb, bias, err)
    return err


@pytest.mark.parametrize("M,N,Kd", [(1, 1, 1), (2000, 1, 3), (3, 2000, 1), (65, 63, 17), (128, 128, 128), (1024, 10, 3456),
                                   (25, 6, 100000), (300, 200, 100), (5, 7, 0)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_simt(ctx, M, N, Kd, ta, tb):
    rng = np.random.default_rng(M * 3 + N * 5 + Kd)
    # tolerance: f32 accumulation over K terms in [-4,4]: 1e-5 of max|C| is ~10x the observed error
    gemm_case(ctx, rng, M, N, Kd, ta, tb, bool((M + N) % 2), K.GEMM_SIMT, 1e-5)
    assert ctx.L.agrs_last_gemm_path(ctx.h) == K.GEMM_SIMT


# ------------------------------------------------------------------ tcgen05 split GEMMs (one-CTA and pair kernels, K splits)
def tuning(ctx, key):
    v = ctypes.c_int64()
    ctx.call("agrs_get_tuning", key, ctypes.byref(v))
    return int(v.value)


@pytest.mark.parametrize("M,N,Kd,ta,tb", [(256, 256, 32, 0, 1), (384, 1000, 520, 0, 1), (1000, 520, 3000, 1, 0), (200, 300, 100, 1, 1),
                                         (2048, 1024, 1024, 0, 0), (132, 260, 4100, 1, 1), (129, 8, 8, 0, 0), (4100, 132, 260, 1, 0)])
@pytest.mark.parametrize("pair,splitk", [(0, 0), (1, 0), (1, 1), (1, 2), (1, 3)])
def test_gemm_tcgen05(ctx, M, N, Kd, ta, tb, pair, splitk):
    """f32-faithful: max error <= 2e-5 of max|C| against the f64 product (measured ~8e-6; plain TF32 would be ~5e-4)"""
    rng = np.random.default_rng(M + 3 * N + 5 * Kd + pair)
    ctx.call("agrs_set_tuning", K.TUNE_TC_2CTA, pair)
    ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, splitk)
    try:
        for bias in (False, True):
            gemm_case(ctx, rng, M, N, Kd, ta, tb, bias, K.GEMM_TCGEN05, 2e-5)
            assert ctx.L.agrs_last_gemm_path(ctx.h) == K.GEMM_TCGEN05
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_2CTA, 1)
        ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 0)


@pytest.mark.parametrize("M,N,Kd,ta,tb", [(256, 256, 32, 0, 1), (384, 1000, 520, 0, 1), (1000, 520, 3000, 1, 0), (200, 300, 100, 1, 1),
                                         (2048, 1024, 1024, 0, 0), (132, 260, 4100, 1, 1), (4100, 132, 260, 1, 0), (512, 384, 96, 0, 0)])
@pytest.mark.parametrize("cross", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("splitk,epi_tma", [(0, 1), (2, 1), (0, 0), (2, 0)])
def test_gemm_pair_cross_modes_and_epilogues(ctx, M, N, Kd, ta, tb, cross, splitk, epi_tma):
    """pair kernel with the two correction terms on tf32 operands (AGRS_TUNE_TC_CROSS 0) or on bf16 operands (1: 64-byte-swizzled
    tiles, the default; 2: unswizzled), as three bf16 products (3) or three fp16 products of row/column-scaled operands (4),
    with the staged TMA epilogue (tensor store + reduce-add across K chunks) and with the direct
    one: same 2e-5 bar on every operand-major combination (the splitter warps transpose MN-major operands), ragged edges included"""
    rng = np.random.default_rng(M + 3 * N + 5 * Kd + cross)
    default_cross = tuning(ctx, K.TUNE_TC_CROSS)
    ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, cross)
    ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, splitk)
    ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, epi_tma)
    try:
        for bias in (False, True):
            gemm_case(ctx, rng, M, N, Kd, ta, tb, bias, K.GEMM_TCGEN05, 2e-5)
            assert ctx.L.agrs_last_gemm_path(ctx.h) == K.GEMM_TCGEN05
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, default_cross)
        ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 0)
        ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, 1)


@pytest.mark.parametrize("M,N,Kd,ta,tb", [(4096, 4096, 512, 1, 0), (4096, 4096, 1024, 0, 0), (3900, 4000, 800, 0, 1)])
def test_gemm_pair_splits_only_the_last_wave(ctx, M, N, Kd, ta, tb):
    """256 tiles on 74 CTA pairs = 3.46 waves: the automatic plan cuts only the 34 tiles of the last wave along K and adds their
    partial planes tile by tile (k_add_partials_tiles).  Same 2e-5 bar against f64 as every other plan, and within 1e-5 of the
    unsplit result; a forced uniform split still works."""
    rng = np.random.default_rng(M + N + Kd)
    A = rng.uniform(-1, 1, (Kd, M) if ta else (M, Kd)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, Kd) if tb else (Kd, N)).astype(np.float32)
    bias = rng.uniform(-1, 1, N).astype(np.float32)
    truth = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64) + bias
    scale = np.abs(truth).max()
    got = {}
    try:
        for splitk in (0, 1, 2):
            ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, splitk)
            got[splitk] = ctx.gemm(A, B, bool(ta), bool(tb), bias, K.GEMM_TCGEN05)
            assert np.abs(got[splitk] - truth).max() / scale < 2e-5
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 0)
    assert np.abs(got[0] - got[1]).max() / scale < 1e-5 and np.abs(got[2] - got[1]).max() / scale < 1e-5
    assert not np.array_equal(got[0], got[1])  # the automatic plan did split something at these shapes


def test_gemm_pair_wave_alignment_changes_no_bit(ctx):
    """AGRS_TUNE_TC_WAVE_SYNC: the producers of a wave wait (bounded) for each other before they load a tile -- a scheduling hint for
    L2 reuse on very large problems; forced on for a 7-wave problem it must give the bits it gives when off"""
    rng = np.random.default_rng(77)
    M, N, Kd = 8192, 4096, 512
    A = rng.uniform(-1, 1, (M, Kd)).astype(np.float32)
    B = rng.uniform(-1, 1, (Kd, N)).astype(np.float32)
    assert tuning(ctx, K.TUNE_TC_WAVE_SYNC) == 2
    out = []
    try:
        for mode in (0, 1):
            ctx.call("agrs_set_tuning", K.TUNE_TC_WAVE_SYNC, mode)
            out.append(ctx.gemm(A, B, False, False, None, K.GEMM_TCGEN05))
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_WAVE_SYNC, 2)
    assert np.array_equal(out[0], out[1])


def test_gemm_pair_epilogues_agree_bitwise(ctx):
    """the staged TMA epilogue (store, then reduce-add per K chunk in L2) and the direct one (read-modify-write from registers)
    perform the same round-to-nearest f32 adds in the same order: identical bits, K chunks, bias and ragged edges included"""
    rng = np.random.default_rng(99)
    M, N, Kd = 700, 1000, 3000
    A = rng.uniform(-1, 1, (M, Kd)).astype(np.float32)
    B = rng.uniform(-1, 1, (Kd, N)).astype(np.float32)
    bias = rng.uniform(-1, 1, N).astype(np.float32)
    out = []
    try:
        for epi in (1, 0):
            ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, epi)
            out.append(ctx.gemm(A, B, False, False, bias, K.GEMM_TCGEN05))
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, 1)
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("cross", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("lo_side", ["a", "b"])
def test_gemm_pair_cross_terms_exact(ctx, ta, tb, cross, lo_side):
    """Known-answer test of the split itself.  One operand is s * (1 + c * 2^-12) with s in {+-1, +-2}, c in 0..3: its tf32 part is
    s and its remainder s*c*2^-12 survives only through the correction term; the other operand holds small integers.  Every partial
    sum is exactly representable, so the result must equal the f64 product BIT FOR BIT -- any misplaced element of a lo / bf16 tile
    (swizzle, transpose, K slice, CTA half) changes it."""
    M, N, Kd = 512, 384, 160
    rng = np.random.default_rng(17 * cross + 2 * ta + tb)
    s = rng.choice(np.array([-2.0, -1.0, 1.0, 2.0]), (M, Kd) if lo_side == "a" else (Kd, N))
    c = rng.integers(0, 4, s.shape)
    split = (s * (1.0 + c * 2.0 ** -12)).astype(np.float32)
    ints = rng.integers(-3, 4, (Kd, N) if lo_side == "a" else (M, Kd)).astype(np.float32)
    opA, opB = (split, ints) if lo_side == "a" else (ints, split)
    truth = opA.astype(np.float64) @ opB.astype(np.float64)
    assert np.array_equal(truth.astype(np.float32).astype(np.float64), truth)  # representable: the test is exact by construction
    A = np.ascontiguousarray(opA.T) if ta else opA
    B = np.ascontiguousarray(opB.T) if tb else opB
    default_cross = tuning(ctx, K.TUNE_TC_CROSS)
    ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, cross)
    try:
        got = ctx.gemm(A, B, bool(ta), bool(tb), None, K.GEMM_TCGEN05)
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, default_cross)
    assert np.array_equal(got.astype(np.float64), truth), float(np.abs(got - truth).max())


def test_gemm_tcgen05_special_values(ctx):
    """NaN and zeros propagate as in f32 arithmetic.  An inf operand gives a NON-FINITE result (inf or NaN, where plain f32
    gives inf): its product with the other operand's lo part is inf * 0 or -inf -- inherent to any hi/lo split scheme."""
    A = np.ones((256, 64), np.float32)
    B = np.ones((64, 256), np.float32)
    A[3, 5] = np.inf
    A[7, 1] = np.nan
    A[9, :] = 0.0
    for pair in (0, 1):
        ctx.call("agrs_set_tuning", K.TUNE_TC_2CTA, pair)
        got = ctx.gemm(A, B, path=K.GEMM_TCGEN05)
        assert not np.any(np.isfinite(got[3])) and np.all(np.isnan(got[7])) and np.all(got[9] == 0.0)
        assert np.all(got[0] == 64.0)
    ctx.call("agrs_set_tuning", K.TUNE_TC_2CTA, 1)


# ------------------------------------------------------------------ skinny (HBM-bound) GEMMs: the LeNet / conv-lowering shapes
@pytest.mark.parametrize("M,N,Kd,ta,tb", [
    (20000, 6, 25, 0, 0),     # conv forward: col . K + bias      (rows kernel)
    (20000, 25, 6, 0, 1),     # conv dxcol: g . K^T               (rows kernel, B transposed)
    (1024, 32, 64, 0, 0), (1500, 1, 1, 0, 1), (4097, 17, 33, 0, 0), (1024, 9, 64, 0, 1),
    (1024, 10, 3456, 0, 0),   # Linear(3456,10) forward           (dot kernel)
    (7, 16, 65, 0, 1), (300, 3, 1000, 0, 0), (1, 1, 100, 0, 0),
    (25, 6, 200000, 1, 0),    # conv dw: col^T . g                (tn kernel, split K + ticket fold)
    (3456, 10, 1024, 1, 0),   # Linear dW: x^T . g                (tn kernel)
    (33, 16, 7, 1, 0), (1, 1, 1, 1, 0), (100, 9, 64, 1, 0),
    (100, 1, 8192, 1, 0),     # dW of Linear(100, 1), batch 8192: M*N <= 160 but too wide for tn_small's staging
    (160, 1, 4096, 1, 0), (40, 4, 5000, 1, 0), (45, 1, 4096, 1, 0), (46, 1, 4096, 1, 0),
])
def test_gemm_skinny(ctx, M, N, Kd, ta, tb):
    rng = np.random.default_rng(M + 7 * N + Kd)
    for bias in (False, True):
        gemm_case(ctx, rng, M, N, Kd, ta, tb, bias, K.GEMM_SKINNY, 1e-5)
        assert ctx.L.agrs_last_gemm_path(ctx.h) == K.GEMM_SKINNY
    # AUTO takes the same path for these shapes, twice in a row (the split-K tickets reset themselves)
    for _ in range(2):
        gemm_case(ctx, rng, M, N, Kd, ta, tb, True, K.GEMM_AUTO, 1e-5)
        assert ctx.L.agrs_last_gemm_path(ctx.h) == K.GEMM_SKINNY


def test_gemm_skinny_refuses_wide_outputs(ctx):
    d = ctx.empty(256 * 256)
    ctx.call("agrs_set_gemm_path", K.GEMM_SKINNY)
    try:
        with pytest.raises(K.AgrsError) as e:
            ctx.call("agrs_gemm_f32", 0, 0, 256, 256, 16, d.ptr, 16, d.ptr, 256, d.ptr, 256, None)
        assert e.value.code == 4
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


def test_gemm_bad_args(ctx):
    d = ctx.empty(64)
    with pytest.raises(K.AgrsError) as e:
        ctx.call("agrs_gemm_f32", 0, 0, 4, 4, 4, d.ptr, 2, d.ptr, 4, d.ptr, 4, None)  # lda < K
    assert e.value.code == 3
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    try:
        with pytest.raises(K.AgrsError) as e:
            ctx.call("agrs_gemm_f32", 0, 0, 4, 3, 3, d.ptr, 3, d.ptr, 3, d.ptr, 3, None)  # ld % 4 != 0: not TMA-mappable
        assert e.value.code == 4
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


# ------------------------------------------------------------------ fused exchange primitives on one rank (world = 1)
def test_fused_scatter_and_step_single_rank(ctx):
    """agrs_peer_alloc / agrs_fused_create / agrs_fused_scatter_f32 / agrs_fused_sgd_step with world = 1: a ragged gradient
    (n % 4 != 0) pushed at a non-zero offset lands in slot 0 at the same offsets (one rank owns every granule), untouched
    elsewhere, and the step applies w -= lr * g exactly (two roundings) on the whole region."""
    import ctypes as C
    L = ctx.L
    flag = int(L.agrs_fused_flag_floats())
    total = 32 * 9
    w_off, g_off = flag, flag + total
    arena = C.c_void_p()
    ctx.call("agrs_peer_alloc", (g_off + total) * 4, C.byref(arena))
    arenas = (C.c_void_p * 1)(arena.value)
    f = C.c_void_p()
    ctx.call("agrs_fused_create", 1, 0, arenas, w_off, 0, g_off, total, 5, C.byref(f))   # 32-float granules, single weight region
    rng = np.random.default_rng(5)
    w = rng.standard_normal(total).astype(np.float32)
    g = rng.standard_normal(37).astype(np.float32)
    ctx.call("agrs_upload", arena.value + 4 * w_off, w.ctypes.data, w.nbytes)
    dg = ctx.to_device(g)
    assert L.agrs_fused_scatter_f32(f, dg.ptr, 64, 37) == 0
    slot = np.empty(total, np.float32)
    ctx.call("agrs_download", slot.ctypes.data, arena.value + 4 * g_off, slot.nbytes)
    want_slot = np.zeros(total, np.float32)
    want_slot[64:64 + 37] = g
    assert np.array_equal(slot, want_slot)
    assert L.agrs_fused_sgd_step(f, C.c_float(0.25)) == 0
    assert L.agrs_fused_sgd_step(f, C.c_float(0.0)) == 0   # a second generation of the barriers; lr = 0 leaves w alone
    got = np.empty(total, np.float32)
    ctx.call("agrs_download", got.ctypes.data, arena.value + 4 * w_off, got.nbytes)
    assert np.array_equal(got, w - want_slot * np.float32(0.25))
    with pytest.raises(K.AgrsError):
        ctx.check(L.agrs_fused_scatter_f32(f, dg.ptr, 2, 37))  # offsets must be multiples of 4
    # the bucket protocol on the same exchange: two buckets signalled in backward order, each updates only its own range
    w1 = got.copy()
    g2 = rng.standard_normal(total).astype(np.float32)
    dg2 = ctx.to_device(g2)
    assert L.agrs_fused_begin_step(f, C.c_float(0.5)) == 0
    assert L.agrs_fused_scatter_f32(f, dg2.ptr, 0, total) == 0
    assert L.agrs_fused_bucket_ready(f, 1, 32 * 4, 32 * 5) == 0    # [128, 288)
    assert L.agrs_fused_bucket_ready(f, 0, 0, 32 * 4) == 0        # [0, 128)
    with pytest.raises(K.AgrsError):
        ctx.check(L.agrs_fused_bucket_ready(f, 0, 0, 32 * 4))     # a bucket is signalled once per step
    assert L.agrs_fused_finish_step(f) == 0
    ctx.call("agrs_download", got.ctypes.data, arena.value + 4 * w_off, got.nbytes)
    assert np.array_equal(got, w1 - g2 * np.float32(0.5))
    with pytest.raises(K.AgrsError):
        ctx.check(L.agrs_fused_finish_step(f))                    # no open step
    memops, streams, agk = C.c_int(), C.c_int(), C.c_int()
    assert L.agrs_fused_info(f, C.byref(memops), C.byref(streams), C.byref(agk)) == 0
    print("fused exchange: stream memory operations", memops.value, "copy streams", streams.value, "all-gather by kernel", agk.value)
    L.agrs_fused_destroy(f)
    ctx.call("agrs_peer_free", arena)


# ------------------------------------------------------------------ operand magnitudes for the GEMM's fp16 split mode
def _bits_of_max(a):
    a = np.asarray(a, np.float32).reshape(-1)
    fin = np.abs(a[np.isfinite(a)])
    return int(np.float32(fin.max() if fin.size else 0.0).view(np.uint32))


def _read_u32(ctx, buf):
    return int(ctx.to_host(buf, (1,)).view(np.uint32)[0])


@pytest.mark.parametrize("n", [1, 5, 1024, 4099, 1 << 20])
def test_absmax_pass_and_producer_reports(ctx, n):
    """agrs_absmax_f32 and the by-product of agrs_next_output_absmax equal the bit pattern of max finite |x| (exact: a maximum
    has no rounding), ignore inf / NaN, and one request serves exactly one launch."""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 3).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    if n > 4:
        x[1], x[3] = np.inf, np.nan
    dx, dg, dy, slot = ctx.to_device(x), ctx.to_device(g), ctx.empty(n), ctx.empty(1)
    ctx.call("agrs_absmax_f32", dx.ptr, n, slot.ptr)
    assert _read_u32(ctx, slot) == _bits_of_max(x)
    with np.errstate(all="ignore"):
        cases = [
            ("agrs_relu_fwd_f32", (dx.ptr, dy.ptr, n), np.where(x > 0, x, 0).astype(np.float32)),
            ("agrs_relu_bwd_f32", (dx.ptr, dg.ptr, dy.ptr, n), np.where(x <= 0, 0, g).astype(np.float32)),
            ("agrs_neg_f32", (dy.ptr, dx.ptr, n), -x),
            ("agrs_scalar_f32", (K.SMUL if hasattr(K, "SMUL") else 0, dy.ptr, dx.ptr, 0.5, n), x * np.float32(0.5)),
            ("agrs_sgd_step_f32", (dg.ptr, dx.ptr, 0.25, n, n), None),
        ]
    for name, args, want in cases:
        ctx.call("agrs_fill_f32", slot.ptr, 123.0, 1)      # garbage: the library zeroes the word itself
        ctx.call("agrs_next_output_absmax", slot.ptr)
        ctx.call(name, *args)
        got_bits = _read_u32(ctx, slot)
        out = ctx.to_host(dg if name == "agrs_sgd_step_f32" else dy, (n,))
        if want is not None:
            assert np.array_equal(out, want, equal_nan=True), name
        assert got_bits == _bits_of_max(out), name
        # consumed: the next launch reports nowhere
        ctx.call("agrs_fill_f32", slot.ptr, 7.0, 1)
        ctx.call("agrs_relu_fwd_f32", dx.ptr, dy.ptr, n)
        assert float(ctx.to_host(slot, (1,))[0]) == 7.0


@pytest.mark.parametrize("M,N,Kd,ta,tb", [(512, 512, 512, 0, 0), (1000, 520, 3000, 1, 0), (384, 1000, 520, 0, 1), (2048, 256, 4100, 1, 1)])
def test_gemm_fp16_mode_with_caller_magnitudes(ctx, M, N, Kd, ta, tb):
    """cross mode 4 with one magnitude per operand (agrs_gemm_operand_absmax): <= 1e-5 of max|C| against the f64 product, rows of
    very different size in one operand included; the announcement serves one call (the next one finds its own per-row scales)."""
    rng = np.random.default_rng(M + N)
    A = rng.uniform(-1, 1, (Kd, M) if ta else (M, Kd)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, Kd) if tb else (Kd, N)).astype(np.float32)
    A[:: 7] *= np.float32(1e-3)          # strips 1000x smaller than the tensor maximum share its scale
    B *= np.float32(37.5)
    dA, dB, dC, sa, sb = ctx.to_device(A), ctx.to_device(B), ctx.empty(M * N), ctx.empty(1), ctx.empty(1)
    ctx.call("agrs_absmax_f32", dA.ptr, A.size, sa.ptr)
    ctx.call("agrs_absmax_f32", dB.ptr, B.size, sb.ptr)
    truth = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    import ctypes as C
    before = C.c_int64()
    ctx.call("agrs_get_tuning", K.TUNE_TC_CROSS, C.byref(before))
    ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, 4)
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    try:
        assert ctx.L.agrs_gemm_wants_absmax(ctx.h, M, N, Kd) == 1
        for announce in (True, False):
            if announce:
                ctx.call("agrs_gemm_operand_absmax", sa.ptr, sb.ptr)
            ctx.call("agrs_gemm_f32", ta, tb, M, N, Kd, dA.ptr, A.shape[1], dB.ptr, B.shape[1], dC.ptr, N, None)
            got = ctx.to_host(dC, (M, N))
            assert np.abs(got - truth).max() / np.abs(truth).max() < 1e-5, announce
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)
        ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, before.value)
    assert ctx.L.agrs_gemm_wants_absmax(ctx.h, M, N, Kd) == (1 if before.value in (4, 5) else 0)
    ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, 1)
    assert ctx.L.agrs_gemm_wants_absmax(ctx.h, M, N, Kd) == 0
    ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, before.value)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("lo_side", ["a", "b"])
def test_gemm_fp16_mode_exact_with_announced_magnitudes(ctx, ta, tb, lo_side):
    """The known-answer construction of test_gemm_pair_cross_terms_exact through the DEFAULT tuning (mode 5) with one announced
    magnitude per operand: the fp16 split must reproduce the f64 product bit for bit (the scales are exact powers of two)."""
    M, N, Kd = 512, 384, 160
    rng = np.random.default_rng(5 + 2 * ta + tb)
    s = rng.choice(np.array([-2.0, -1.0, 1.0, 2.0]), (M, Kd) if lo_side == "a" else (Kd, N))
    c = rng.integers(0, 4, s.shape)
    split = (s * (1.0 + c * 2.0 ** -12)).astype(np.float32)
    ints = rng.integers(-3, 4, (Kd, N) if lo_side == "a" else (M, Kd)).astype(np.float32)
    opA, opB = (split, ints) if lo_side == "a" else (ints, split)
    truth = opA.astype(np.float64) @ opB.astype(np.float64)
    A = np.ascontiguousarray(opA.T) if ta else opA
    B = np.ascontiguousarray(opB.T) if tb else opB
    dA, dB, dC, sa, sb = ctx.to_device(A), ctx.to_device(B), ctx.empty(M * N), ctx.empty(1), ctx.empty(1)
    ctx.call("agrs_absmax_f32", dA.ptr, A.size, sa.ptr)
    ctx.call("agrs_absmax_f32", dB.ptr, B.size, sb.ptr)
    if tuning(ctx, K.TUNE_TC_CROSS) not in (4, 5):
        pytest.skip("the context's default split mode does not use magnitudes")
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    try:
        ctx.call("agrs_gemm_operand_absmax", sa.ptr, sb.ptr)
        ctx.call("agrs_gemm_f32", ta, tb, M, N, Kd, dA.ptr, A.shape[1], dB.ptr, B.shape[1], dC.ptr, N, None)
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)
    got = ctx.to_host(dC, (M, N))
    assert np.array_equal(got.astype(np.float64), truth), float(np.abs(got - truth).max())


# ------------------------------------------------------------------ f2: backward + column sum in one pass
@pytest.mark.parametrize("rows,cols", [(64, 32), (1000, 520), (8192, 4096), (4097, 36)])
def test_backward_colsum_fusions_bit_identical(ctx, rows, cols):
    """agrs_{relu,sigmoid,mse}_bwd_colsum_f32 == the unfused pair (elementwise backward, then agrs_sum_axis_f32 over rows), bit for
    bit, for the result and for the column sums; the magnitude by-product works through the fused kernels too."""
    rng = np.random.default_rng(rows + cols)
    n = rows * cols
    x = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    up = np.array([0.75], np.float32)
    dx, dg, dup = ctx.to_device(x), ctx.to_device(g), ctx.to_device(up)
    o1, o2, c1, c2, slot = ctx.empty(n), ctx.empty(n), ctx.empty(cols), ctx.empty(cols), ctx.empty(1)
    cases = [
        ("relu", lambda o: ctx.call("agrs_relu_bwd_f32", dx.ptr, dg.ptr, o.ptr, n), lambda o, c: ctx.call("agrs_relu_bwd_colsum_f32", dx.ptr, dg.ptr, o.ptr, rows, cols, c.ptr)),
        ("sigmoid", lambda o: ctx.call("agrs_sigmoid_bwd_f32", dx.ptr, dg.ptr, o.ptr, n), lambda o, c: ctx.call("agrs_sigmoid_bwd_colsum_f32", dx.ptr, dg.ptr, o.ptr, rows, cols, c.ptr)),
        ("mse", lambda o: ctx.call("agrs_mse_bwd_f32", dx.ptr, dg.ptr, dup.ptr, float(rows), n, o.ptr),
         lambda o, c: ctx.call("agrs_mse_bwd_colsum_f32", dx.ptr, dg.ptr, dup.ptr, float(rows), rows, cols, o.ptr, c.ptr)),
    ]
    for name, plain, fused in cases:
        plain(o1)
        ctx.call("agrs_sum_axis_f32", o1.ptr, 1, rows, cols, c1.ptr)
        ctx.call("agrs_next_output_absmax", slot.ptr)
        fused(o2, c2)
        a, b = ctx.to_host(o1, (n,)), ctx.to_host(o2, (n,))
        assert np.array_equal(a, b), name
        assert np.array_equal(ctx.to_host(c1, (cols,)), ctx.to_host(c2, (cols,))), name
        assert _read_u32(ctx, slot) == _bits_of_max(a), name


def test_mse_backward_power_of_two_batch_rounds_like_the_division(ctx):
    """With a power-of-two batch the fused kernel multiplies by 1 / batch instead of dividing (reduce.cu): the same single rounding,
    also where the result is subnormal or the intermediate overflows -- checked bit for bit against the dividing elementwise kernel
    and against NumPy's f32 division (operators.rs:249-253: ((p - t) * 2.0 / B) * grad)."""
    rows, cols = 256, 64
    n = rows * cols
    rng = np.random.default_rng(77)
    scale = np.float32(10.0) ** rng.integers(-44, 38, n).astype(np.float32)       # from deep subnormal results to overflow of (p - t) * 2
    p = (rng.standard_normal(n).astype(np.float32) * scale).astype(np.float32)
    t = (rng.standard_normal(n).astype(np.float32) * scale * np.float32(0.5)).astype(np.float32)
    up = np.array([1.0], np.float32)
    dp, dt, dup = ctx.to_device(p), ctx.to_device(t), ctx.to_device(up)
    o1, o2, c2 = ctx.empty(n), ctx.empty(n), ctx.empty(cols)
    ctx.call("agrs_mse_bwd_f32", dp.ptr, dt.ptr, dup.ptr, float(rows), n, o1.ptr)
    ctx.call("agrs_mse_bwd_colsum_f32", dp.ptr, dt.ptr, dup.ptr, float(rows), rows, cols, o2.ptr, c2.ptr)
    a, b = ctx.to_host(o1, (n,)), ctx.to_host(o2, (n,))
    with np.errstate(over="ignore", invalid="ignore"):
        want = (((p - t) * np.float32(2.0)) / np.float32(rows)) * up[0]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(b.view(np.uint32), want.astype(np.float32).view(np.uint32))
    assert np.count_nonzero((np.abs(want) < np.float32(1.1754944e-38)) & (want != 0)) > 100   # the subnormal range is exercised


def test_backward_colsum_refuses_unaligned(ctx):
    d = ctx.empty(64 * 6)
    with pytest.raises(K.AgrsError) as e:
        ctx.call("agrs_relu_bwd_colsum_f32", d.ptr, d.ptr, d.ptr, 64, 6, d.ptr)   # cols % 4 != 0
    assert e.value.code == 4
