"""Pins the NumPy oracle against every golden vector the reference's own unit tests hold for
the hot path (tests/golden/reference_kats.json) and against the derived fit_sine curve of
SURVEY.md Appendix E.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_np as O


def A(v, dt=np.float32):
    return np.array(v, dtype=dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_relu_on_mat(kats, dt):
    k = kats["relu_on_mat"]
    assert np.array_equal(O.relu_fwd(A(k["x"], dt)), A(k["y"], dt))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_linear(kats, dt):
    k = kats["linear"]
    assert np.array_equal(O.linear_fwd(A(k["x"], dt), A(k["w"], dt), A(k["b"], dt)), A(k["y"], dt))


@pytest.mark.parametrize("name", ["im2col_nopad", "im2col_pad"])
def test_im2col(kats, name):
    k = kats[name]
    im = np.arange(*k["im_range"], dtype=np.float32).reshape(k["im_shape"])
    want = A(k["col"]).reshape(k["col_shape"])
    assert np.array_equal(O.im2col(im, k["k"], k["stride"], k["pad"]), want)
    assert np.array_equal(O.im2col_fast(im, k["k"], k["stride"], k["pad"]), want)


def test_conv2d_ones(kats):
    k = kats["conv2d_ones"]
    im, filt, bias = np.ones(k["im_shape"], np.float32), np.ones(k["filt_shape"], np.float32), np.ones(k["bias_shape"], np.float32)
    out, col2d = O.conv2d_fwd(im, filt, bias, k["c_in"], k["k"], k["stride"], k["pad"])
    assert list(out.shape) == k["out_shape"] and np.all(out == k["out_value"])
    g = np.ones(k["out_shape"], np.float32)
    for lit in (True, False):
        grads = O.conv2d_bwd(im, filt, col2d, g, k["c_in"], k["k"], k["stride"], k["pad"], literal=lit)
        assert [list(x.shape) for x in grads] == k["bwd_shapes"]


def test_conv2d_literal_equals_generalised_at_b1_and_d1_panics():
    rng = np.random.default_rng(0)
    im = rng.standard_normal((1, 3, 6, 6)).astype(np.float32)
    filt = rng.standard_normal((3, 3, 4)).astype(np.float32)
    bias = rng.standard_normal(4).astype(np.float32)
    out, col2d = O.conv2d_fwd(im, filt, bias, 3, 3, 1, 0)
    g = rng.standard_normal(out.shape).astype(np.float32)
    a = O.conv2d_bwd(im, filt, col2d, g, 3, 3, 1, 0, literal=True)
    b = O.conv2d_bwd(im, filt, col2d, g, 3, 3, 1, 0, literal=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    im2 = np.concatenate([im, im])
    out2, col2 = O.conv2d_fwd(im2, filt, bias, 3, 3, 1, 0)
    with pytest.raises(IndexError):  # defect D1: the reference panics for B > 1
        O.conv2d_bwd(im2, filt, col2, np.concatenate([g, g]), 3, 3, 1, 0, literal=True)
    # generalised batch-2 == per-sample literal calls (dx stacked, dw/db summed)
    dx2, dw2, db2 = O.conv2d_bwd(im2, filt, col2, np.concatenate([g, g]), 3, 3, 1, 0, literal=False)
    assert np.allclose(dx2[0], a[0][0]) and np.allclose(dx2[1], a[0][0])
    assert np.allclose(dw2, 2 * a[1], rtol=1e-5, atol=1e-5) and np.allclose(db2, 2 * a[2], rtol=1e-5, atol=1e-5)


def test_single_node_graph(kats):
    k = kats["single_node_graph"]
    x = O.OTensor(A(k["x"]))
    res = O.OpReLU().forward([x.clone()])
    assert np.array_equal(res.data, A(k["y"]))
    res.backward()
    assert np.array_equal(x.grad, A(k["x_grad"]))


def test_simple_graph(kats):
    k = kats["simple_graph"]
    x = O.OTensor(A(k["x"]))
    res = O.OpReLU().forward([x.clone()])
    w, b = O.OTensor(A(k["w"])), O.OTensor(A(k["b"]))
    res = O.OpLinear().forward([res, w.clone(), b.clone()])
    assert np.array_equal(res.data, A(k["y"]))
    res.backward()
    assert np.array_equal(x.grad, A(k["x_grad"]))
    assert np.array_equal(w.grad, A(k["w_grad"]))
    assert b.grad.ndim == 1 and np.array_equal(b.grad, A(k["b_grad"]))


def test_mse_d4_ratio_and_sigmoid_vs_f64():
    rng = np.random.default_rng(1)
    p, t = rng.standard_normal((8, 5)), rng.standard_normal((8, 5))
    g = O.mse_bwd(p, t, np.ones((1, 1)))
    analytic = 2 * (p - t) / p.size
    assert np.allclose(g, analytic * 5)  # D4: D x the analytic gradient
    x = rng.standard_normal((64, 7)).astype(np.float32)
    up = rng.standard_normal((64, 7)).astype(np.float32)
    assert O.rel_err(O.sigmoid_bwd(x, up), O.sigmoid_bwd(x.astype(np.float64), up.astype(np.float64))) < 1e-6


def test_matmul_backward_literal_d3():
    rng = np.random.default_rng(2)
    x, w, g = (rng.standard_normal((4, 4)).astype(np.float32) for _ in range(3))
    dx, dw = O.matmul_bwd(x, w, g, literal=True)
    tdx, tdw = O.matmul_bwd(x, w, g, literal=False)
    assert np.allclose(dx, g @ w) and np.allclose(dw, tdw.T, atol=1e-5)
    assert not np.allclose(dx, tdx)


def test_sgd_broadcast_and_zero_grad_shapes():
    w = np.ones((1, 4), np.float32)
    O.sgd_step(w, np.array([1, 2, 3, 4], np.float32), 0.5)
    assert np.array_equal(w, A([[0.5, 0.0, -0.5, -1.0]]))
    p = O.OTensor(np.ones((1, 4), np.float32))
    O.model_zero_grad([p])
    assert p.grad is None  # lazy: untouched when grad is None (model.rs:16-19)
    p.cell["grad"] = np.ones(4, np.float32)
    O.model_zero_grad([p])
    assert p.grad.shape == (1, 4)  # zeros(shape of data)
    O.accumulate(p.cell["grad"], np.ones(4, np.float32))  # [1,4] += [4]
    assert np.array_equal(p.grad, np.ones((1, 4), np.float32))


def test_fit_sine_curve_matches_survey_appendix_e():
    """Derived golden (W=0,b=0 init, lr=1e-3, D2 fixed)."""
    want = {0: 0.49974999, 1: 0.44477817, 2: 0.41598400, 99: 0.31669122, 999: 0.05300604, 1999: 0.01055911}
    for dt, tol in ((np.float32, 2e-5), (np.float64, 2e-5)):
        x, y = O.fit_sine_data(2000, dt)
        losses, layers = O.mlp_train(x, y, [np.zeros((3, 1))], [np.zeros((1, 1))], None, 1e-3, 2000, dt)
        for s, v in want.items():
            assert abs(losses[s] - v) / v < tol, (dt, s, losses[s], v)
    w = layers[0].w.data.reshape(-1)
    assert abs(w[0] - 0.748815654) < 1e-5 and abs(w[2] + 0.0779789703) < 1e-5


def test_kaiming_bounds():
    # init.rs:19-45: fan_in uses shape[1]
    assert abs(O.kaiming_bound([20, 10]) - np.sqrt(3.0) * np.sqrt(2.0 / 6.0) / np.sqrt(10)) < 1e-6
    assert abs(O.kaiming_bound([5, 5, 6]) - np.sqrt(3.0) * np.sqrt(2.0 / 6.0) / np.sqrt(5 * 6)) < 1e-6
