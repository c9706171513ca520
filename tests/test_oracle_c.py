"""Pins the single-threaded C restatement (oracle/oracle_c.c) against the reference's own golden vectors and against the
NumPy oracle: two independent restatements of the same operator decomposition must agree to rounding.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_c as C
from oracle import oracle_np as O


def A(v):
    return np.array(v, dtype=np.float32)


def test_library_builds_and_exports():
    L = C.lib()
    for name in ("oc_sgemm", "oc_mlp_step", "oc_im2col", "oc_col2im", "oc_sgd", "oc_mse_fwd", "oc_mse_bwd"):
        assert hasattr(L, name)


def test_relu_and_linear_kats(kats):
    k = kats["relu_on_mat"]
    assert np.array_equal(C.relu_fwd(A(k["x"])), A(k["y"]))
    k = kats["linear"]
    assert np.array_equal(C.linear_fwd(A(k["x"]), A(k["w"]), A(k["b"])), A(k["y"]))


@pytest.mark.parametrize("name", ["im2col_nopad", "im2col_pad"])
def test_im2col_kats(kats, name):
    k = kats[name]
    im = np.arange(*k["im_range"], dtype=np.float32).reshape(k["im_shape"])
    assert np.array_equal(C.im2col(im, k["k"], k["stride"], k["pad"]), A(k["col"]).reshape(k["col_shape"]))


def test_simple_graph_kat_through_c_ops(kats):
    """linear(relu(x)) with a ones seed (autograd.rs:121-171 simple_graph), evaluated with the C operators"""
    k = kats["simple_graph"]
    x, w, b = A(k["x"]), A(k["w"]), A(k["b"])
    h = C.relu_fwd(x)
    y = C.linear_fwd(h, w, b)
    assert np.array_equal(y, A(k["y"]))
    dh, dw, db = C.linear_bwd(h, w, np.ones_like(y))
    dx = C.relu_bwd(x, dh)
    assert np.array_equal(dx, A(k["x_grad"])) and np.array_equal(dw, A(k["w_grad"]))
    assert np.array_equal(db.reshape(-1), A(k["b_grad"]).reshape(-1))


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 1, 2), (8, 8, 8), (9, 257, 17), (70, 300, 45), (129, 513, 1031), (65, 64, 1025)])
def test_sgemm_blocking_edges(shape):
    """panel/block edges of the packed loop nest (MR = NR = 8, MC = 64, KC = 256, NC = 1024) against float64"""
    m, k, n = shape
    rng = np.random.default_rng(m * 1000 + k)
    a, b = rng.standard_normal((m, k), dtype=np.float32), rng.standard_normal((k, n), dtype=np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    assert O.rel_err(C.dot(a, b), want) < 2e-6
    ints = rng.integers(-4, 5, (m, k)).astype(np.float32)  # exactly representable: any summation order gives the same bits
    intb = rng.integers(-4, 5, (k, n)).astype(np.float32)
    assert np.array_equal(C.dot(ints, intb), ints @ intb)


def test_elementwise_ops_match_numpy_oracle():
    rng = np.random.default_rng(1)
    x, g = rng.standard_normal((37, 53), dtype=np.float32) * 4, rng.standard_normal((37, 53), dtype=np.float32)
    assert np.array_equal(C.relu_fwd(x), O.relu_fwd(x)) and np.array_equal(C.relu_bwd(x, g), O.relu_bwd(x, g))
    assert O.rel_err(C.sigmoid_fwd(x), O.sigmoid_fwd(x)) < 3e-7  # libm expf vs numpy's exp: last-ulp differences only
    assert O.rel_err(C.sigmoid_bwd(x, g), O.sigmoid_bwd(x, g)) < 1e-6
    t = rng.standard_normal((37, 53), dtype=np.float32)
    assert abs(float(C.mse_fwd(x, t)) - float(np.asarray(O.mse_fwd(x, t)).reshape(-1)[0])) < 1e-5 * float(C.mse_fwd(x, t))
    assert O.rel_err(C.mse_bwd(x, t), O.mse_bwd(x, t, np.ones((1, 1), np.float32))) < 3e-7


@pytest.mark.parametrize("geom", [(1, 1, 5, 5, 3, 1, 0), (2, 3, 8, 7, 3, 2, 1), (3, 2, 6, 6, 2, 2, 0), (1, 4, 9, 9, 5, 1, 2)])
def test_im2col_col2im_match_numpy_oracle(geom):
    B, Cc, H, W, k, s, p = geom
    rng = np.random.default_rng(7)
    im = rng.standard_normal((B, Cc, H, W), dtype=np.float32)
    col = C.im2col(im, k, s, p)
    assert np.array_equal(col, O.im2col(im, k, s, p))
    ho, wo = O.conv_out_hw(H, W, k, s, p)
    g = rng.standard_normal(col.shape, dtype=np.float32)
    got = C.col2im(g, Cc, ho, wo, s, k)
    want = O.col2im(g, ho, wo, s, k)
    assert got.shape == want.shape and np.array_equal(got, want)  # same patch order -> same bits


def test_sgd_two_roundings_and_bias_broadcast():
    rng = np.random.default_rng(3)
    w, g = rng.standard_normal((6, 5), dtype=np.float32), rng.standard_normal((6, 5), dtype=np.float32)
    want = w.copy(); O.sgd_step(want, g, 0.1)
    got = w.copy(); C.sgd_step(got, g, 0.1)
    assert np.array_equal(got, want)
    b, gb = rng.standard_normal((1, 5), dtype=np.float32), rng.standard_normal((1, 5), dtype=np.float32)
    want = b.copy(); O.sgd_step(want, gb, 0.1)
    got = b.copy(); C.sgd_step(got, gb, 0.1)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("act", [None, "relu", "sigmoid"])
def test_mlp_training_matches_numpy_oracle(act):
    """the whole iteration (forward, MSE, backward, SGD, zero_grad) in C against the graph-walking NumPy oracle"""
    rng = np.random.default_rng(5)
    dims = [16, 40, 33, 8]
    W = [rng.uniform(-0.3, 0.3, (dims[i], dims[i + 1])).astype(np.float32) for i in range(3)]
    Bs = [rng.uniform(-0.3, 0.3, (1, dims[i + 1])).astype(np.float32) for i in range(3)]
    x, t = rng.standard_normal((70, 16), dtype=np.float32), rng.standard_normal((70, 8), dtype=np.float32)
    losses, layers = O.mlp_train(x, t, W, Bs, act, 0.05, 8)
    Wc, Bc = [w.copy() for w in W], [b.copy() for b in Bs]
    lc = [C.mlp_step(Wc, Bc, x, t, act, 0.05) for _ in range(8)]
    assert max(abs(a - b) / abs(a) for a, b in zip(losses, lc)) < 1e-5
    assert losses[-1] < losses[0]
    for wc, bc, l in zip(Wc, Bc, layers):
        assert O.rel_err(wc, l.w.data) < 1e-5 and O.rel_err(bc, l.b.data) < 1e-5


def test_fit_sine_curve_in_c():
    """examples/fit_sine.rs with the seed-independent start used by the NumPy pin: same loss curve from the C restatement"""
    x, y = O.fit_sine_data()
    rng = np.random.default_rng(0)
    w, b = rng.uniform(-0.5, 0.5, (3, 1)).astype(np.float32), rng.uniform(-0.5, 0.5, (1, 1)).astype(np.float32)
    losses, _ = O.mlp_train(x, y, [w], [b], None, 1e-3, 200)
    Wc, Bc = [w.copy()], [b.copy()]
    lc = [C.mlp_step(Wc, Bc, x, y, None, 1e-3) for _ in range(200)]
    assert max(abs(a - c) / abs(a) for a, c in zip(losses, lc)) < 1e-4
