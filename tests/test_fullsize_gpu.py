"""GPU tests at BASELINE.json's FULL sizes (C2: 8192 x 4096 activations, 4096^2 weights; C3: 1024 x 1 x 28 x 28), where running
the oracle on everything would take minutes: sampled rows against an f64 product, and size-independent properties of the
domain (linearity of the GEMM, idempotence of ReLU, exact integer sums, SGD fixed points, im2col -> col2im tap counts)."""
import numpy as np
import pytest

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import _capi as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = K.Ctx(0)
    yield c
    c.close()


def dev_gemm(ctx, dA, dB, dC, M, N, Kd, ta, tb, lda, ldb, bias=None):
    ctx.call("agrs_gemm_f32", int(ta), int(tb), M, N, Kd, dA.ptr, lda, dB.ptr, ldb, dC.ptr, N, bias.ptr if bias else None)


@pytest.mark.parametrize("M,N,Kd,ta,tb", [(8192, 4096, 4096, 0, 0), (8192, 4096, 4096, 0, 1), (4096, 4096, 8192, 1, 0)])
def test_linear_gemms_full_size(ctx, M, N, Kd, ta, tb):
    """forward X.W, dX = G.W^T, dW = X^T.G of the 4096-wide layer at batch 8192: 64 sampled rows within 2e-5 of the f64 product
    (contract: 1e-4), and linearity in B: gemm(A, B1 + B2) == gemm(A, B1) + gemm(A, B2) to the same tolerance, full matrices."""
    rng = np.random.default_rng(M + N + Kd + ta)
    A = rng.uniform(-1, 1, (Kd, M) if ta else (M, Kd)).astype(np.float32)
    B1 = rng.uniform(-1, 1, (N, Kd) if tb else (Kd, N)).astype(np.float32)
    B2 = rng.uniform(-1, 1, B1.shape).astype(np.float32)
    dA, dB1, dB2, dB12 = ctx.to_device(A), ctx.to_device(B1), ctx.to_device(B2), ctx.to_device(B1 + B2)
    dC1, dC2, dC12 = ctx.empty(M * N), ctx.empty(M * N), ctx.empty(M * N)
    for dB, dC in ((dB1, dC1), (dB2, dC2), (dB12, dC12)):
        dev_gemm(ctx, dA, dB, dC, M, N, Kd, ta, tb, A.shape[1], B1.shape[1])
    assert ctx.L.agrs_last_gemm_path(ctx.h) == K.GEMM_TCGEN05
    C1, C2, C12 = (ctx.to_host(d, (M, N)) for d in (dC1, dC2, dC12))
    rows = np.unique(np.concatenate([[0, 1, M - 1], rng.integers(0, M, 61)]))
    opA = (A.T if ta else A)[rows].astype(np.float64)
    truth = opA @ (B1.T if tb else B1).astype(np.float64)
    scale = np.abs(truth).max()
    assert np.abs(C1[rows] - truth).max() / scale < 2e-5
    lin = np.abs(C12.astype(np.float64) - (C1.astype(np.float64) + C2.astype(np.float64))).max() / np.abs(C12).max()
    assert lin < 4e-5, lin
    assert np.isfinite(C12).all()


def test_elementwise_properties_full_size(ctx):
    n = 8192 * 4096
    rng = np.random.default_rng(7)
    x = rng.standard_normal(n).astype(np.float32)
    dx, dy, dz = ctx.to_device(x), ctx.empty(n), ctx.empty(n)
    ctx.call("agrs_relu_fwd_f32", dx.ptr, dy.ptr, n)
    ctx.call("agrs_relu_fwd_f32", dy.ptr, dz.ptr, n)
    y = ctx.to_host(dy, (n,))
    assert np.array_equal(y, np.maximum(x, 0)) and np.array_equal(ctx.to_host(dz, (n,)), y)  # exact, and idempotent
    # relu backward with an all-ones upstream is the indicator of x > 0
    ctx.call("agrs_fill_f32", dz.ptr, 1.0, n)
    ctx.call("agrs_relu_bwd_f32", dx.ptr, dz.ptr, dy.ptr, n)
    assert np.array_equal(ctx.to_host(dy, (n,)), (x > 0).astype(np.float32))
    # sums of ones are exact integers (2^25 is representable): the reduction tree loses nothing
    one = ctx.empty(1)
    ctx.call("agrs_sum_f32", dz.ptr, n, one.ptr)
    assert float(ctx.to_host(one, (1,))[0]) == float(n)
    col = ctx.empty(4096)
    ctx.call("agrs_sum_axis_f32", dz.ptr, 1, 8192, 4096, col.ptr)
    assert np.all(ctx.to_host(col, (4096,)) == 8192.0)
    # mse(x, x) == 0 and its gradient is 0; SGD with lr = 0 is a fixed point, with g = 0 too
    ctx.call("agrs_mse_fwd_f32", dx.ptr, dx.ptr, n, one.ptr)
    assert float(ctx.to_host(one, (1,))[0]) == 0.0
    ctx.call("agrs_fill_f32", one.ptr, 1.0, 1)
    ctx.call("agrs_mse_bwd_f32", dx.ptr, dx.ptr, one.ptr, 8192.0, n, dy.ptr)
    assert not ctx.to_host(dy, (n,)).any()
    ctx.call("agrs_sgd_step_f32", dx.ptr, dz.ptr, 0.0, n, n)
    assert np.array_equal(ctx.to_host(dx, (n,)), x)
    ctx.call("agrs_sgd_step_f32", dx.ptr, dy.ptr, 0.5, n, n)
    assert np.array_equal(ctx.to_host(dx, (n,)), x)
    # sigmoid(-x) + sigmoid(x) == 1 to 2 ulp, and sigmoid is within (0, 1]
    ctx.call("agrs_sigmoid_fwd_f32", dx.ptr, dy.ptr, n)
    ctx.call("agrs_neg_f32", dz.ptr, dx.ptr, n)
    ctx.call("agrs_sigmoid_fwd_f32", dz.ptr, dz.ptr, n)
    s, sn = ctx.to_host(dy, (n,)), ctx.to_host(dz, (n,))
    assert np.abs(s + sn - 1.0).max() <= 3e-7 and s.min() > 0 and s.max() <= 1


def test_conv_lowering_properties_c3_size(ctx):
    """C3: im[1024,1,28,28], k=5: col2im(im2col(im)) multiplies every pixel by the number of patches that cover it -- a pattern
    that depends only on the geometry -- and the column sums of col equal tap-shifted window sums of the image."""
    B, H, k = 1024, 28, 5
    Ho = H - k + 1
    rng = np.random.default_rng(11)
    im = rng.integers(-8, 9, (B, 1, H, H)).astype(np.float32)  # small integers: every sum below is exact in f32
    dim, dcol, dback = ctx.to_device(im), ctx.empty(B * Ho * Ho * k * k), ctx.empty(B * H * H)
    ctx.call("agrs_im2col_f32", dim.ptr, B, 1, H, H, k, 1, 0, dcol.ptr)
    ctx.call("agrs_col2im_f32", dcol.ptr, B, 1, Ho, Ho, k, 1, dback.ptr)
    cover1 = np.minimum(np.minimum(np.arange(H) + 1, k), np.minimum(H - np.arange(H), k)).astype(np.float32)
    cover1 = np.minimum(cover1, Ho)
    cover = np.outer(cover1, cover1)
    assert np.array_equal(ctx.to_host(dback, (B, 1, H, H)), im * cover)
    col = ctx.to_host(dcol, (B, Ho * Ho, k * k))
    for i, j in ((0, 0), (2, 3), (4, 4)):
        assert np.array_equal(col[:, :, i * k + j].reshape(B, Ho, Ho), im[:, 0, i:i + Ho, j:j + Ho])


# ------------------------------------------------------------------ full-size configurations against the oracle
from autograd_rs_b200 import frontend as F  # noqa: E402
from autograd_rs_b200 import workloads as W  # noqa: E402
from oracle import oracle_np as O  # noqa: E402

TOL = 1e-4  # BASELINE.json north star


@pytest.fixture(scope="module")
def dev():
    d = F.Device(0)
    yield d
    d.close()


@pytest.mark.parametrize("n,ta,tb", [(8192, 0, 0), (8192, 1, 0), (16384, 0, 0), (16384, 0, 1)])
def test_gemm_cube_8192_and_16384_sampled_vs_f64(ctx, n, ta, tb):
    """C4's upper end: n^3 through the default path, 40 sampled rows of C within 2e-5 of the f64 product (contract 1e-4) and the
    whole result finite.  16384^3 is 2 x 1 GB of operands and 8.8 TFLOP."""
    rng = np.random.default_rng(n + ta + 2 * tb)
    A = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    B = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    dA, dB, dC = ctx.to_device(A), ctx.to_device(B), ctx.empty(n * n)
    dev_gemm(ctx, dA, dB, dC, n, n, n, ta, tb, n, n)
    assert ctx.L.agrs_last_gemm_path(ctx.h) == K.GEMM_TCGEN05
    C = ctx.to_host(dC, (n, n))
    rows = np.unique(np.concatenate([[0, 1, n - 1], rng.integers(0, n, 37)]))
    opA = (A.T if ta else A)[rows].astype(np.float64)
    truth = opA @ (B.T if tb else B).astype(np.float64)
    assert np.abs(C[rows] - truth).max() / np.abs(truth).max() < 2e-5
    assert np.isfinite(C).all()


def test_lenet_c3_full_batch_tracks_oracle(dev):
    """C3 at its full batch (BASELINE configs[2]: im[1024,1,28,28] -> Conv2D(1,6,5) -> ReLU -> Flatten -> Linear(3456,10) -> MSE):
    two training iterations, every loss, every parameter and the input gradient within 1e-4 of the oracle."""
    rng = np.random.default_rng(31)
    B = 1024
    im = rng.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)
    t = rng.uniform(0, 1, (B, 10)).astype(np.float32)
    fb = W.kaiming_bound((5, 5, 6))
    filt = rng.uniform(-fb, fb, (5, 5, 6)).astype(np.float32)
    cbias = rng.uniform(-0.1, 0.1, (6,)).astype(np.float32)
    w = rng.uniform(-0.05, 0.05, (3456, 10)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (1, 10)).astype(np.float32)
    lr, steps = 1e-3, 2
    model = W.build_lenet(dev, filt, cbias, w, b)
    tr = W.Trainer(model, lr)
    X, Tt = F.Tensor.from_numpy(dev, im), F.Tensor.from_numpy(dev, t)
    Tt.requires_grad = False
    got = [tr.step(X, Tt).item() for _ in range(steps)]
    conv, lin = O.OConvLayer(filt.copy(), cbias.copy(), 1), O.OLinearLayer(w.copy(), b.copy())
    params = conv.parameters() + lin.parameters()
    OX, OT = O.OTensor(im), O.OTensor(t, requires_grad=False)
    want = []
    for _ in range(steps):
        h = O.OpFlatten().forward([O.OpReLU().forward([conv.forward(OX.clone())])])
        loss = O.OpMSE().forward([lin.forward(h), OT.clone()])
        want.append(float(loss.data[0, 0]))
        loss.backward()
        O.sgd(params, lr)
        O.model_zero_grad(params)
    assert max(abs(a - c) / abs(c) for a, c in zip(got, want)) < TOL
    for p, op in zip(model.parameters(), params):
        assert O.rel_err(p.data(), op.data) < TOL
    assert O.rel_err(X.grad().data(), OX.grad) < TOL


def test_sigmoid_layer_8192_wide_fwd_bwd_vs_oracle(dev):
    """One layer pair of C5 at its full width (Linear(8192,8192) -> Sigmoid -> Linear(8192,8192) -> MSE) on 512 rows: prediction,
    loss, input gradient and all four parameter gradients within 1e-4 of the f32 oracle."""
    ws, bs = W.mlp_init(2, 8192, seed=41)
    x, t = W.mlp_batch(512, 8192, 8192, seed=42)
    model = W.build_mlp(dev, ws, bs, "sigmoid")
    X, Tt = F.Tensor.from_numpy(dev, x), F.Tensor.from_numpy(dev, t)
    Tt.requires_grad = False
    pred = model.forward([X.clone()])
    loss = F.nn.MeanSquaredError().forward([pred, Tt.clone()])
    loss.backward()
    ol = [O.OLinearLayer(w, b) for w, b in zip(ws, bs)]
    OX, OT = O.OTensor(x), O.OTensor(t, requires_grad=False)
    h = O.OpSigmoid().forward([ol[0].forward(OX.clone())])
    opred = ol[1].forward(h)
    oloss = O.OpMSE().forward([opred, OT.clone()])
    oloss.backward()
    assert abs(loss.item() - float(oloss.data[0, 0])) / abs(float(oloss.data[0, 0])) < TOL
    assert O.rel_err(pred.data(), opred.data) < TOL
    assert O.rel_err(X.grad().data(), OX.grad) < TOL
    for p, op in zip(model.parameters(), [q for l in ol for q in l.parameters()]):
        assert O.rel_err(p.grad().data(), op.grad) < TOL
