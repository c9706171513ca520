"""An independent cross-check of the oracle rows the reference itself does not pin (Sigmoid, MSE, Linear / Conv2D backward
values, SGD, whole training loops have no golden vector in the reference's tests, and the reference cannot be built here).  The
oracle's f64 flavour -- the arbiter of every GPU parity test -- is compared with PyTorch's CPU autograd, a third implementation
that shares no code with either side.  Where the reference deviates from the textbook on purpose (Appendix B of SURVEY.md: MSE's
gradient is scaled by 1/B instead of 1/(B*D); Conv2D writes NHWC and shares one filter over input channels) the test states the
deviation as arithmetic on torch's result, so the literal rule itself is what is checked.  CPU only; torch is the checker here, the
product path never sees it."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O

TOL = 1e-12


def t64(a, grad=False):
    return torch.tensor(np.asarray(a, np.float64), requires_grad=grad)


def close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_sigmoid_forward_and_backward():
    """operators.rs:189-225: s = 1 / (1 + exp(-x)); dx = ((1 - s) * s) * g"""
    rng = np.random.default_rng(1)
    x, g = rng.standard_normal((37, 11)) * 4, rng.standard_normal((37, 11))
    X = t64(x, True)
    Y = torch.sigmoid(X)
    Y.backward(t64(g))
    assert close(O.sigmoid_fwd(x), Y.detach().numpy())
    assert close(O.sigmoid_bwd(x, g), X.grad.numpy())


def test_relu_forward_and_backward_away_from_the_kink():
    """functional_ndarray.rs:14-34 (the reference pins the forward; the backward only through one graph test)"""
    rng = np.random.default_rng(2)
    x, g = rng.standard_normal((19, 23)), rng.standard_normal((19, 23))
    X = t64(x, True)
    Y = torch.relu(X)
    Y.backward(t64(g))
    assert close(O.relu_fwd(x), Y.detach().numpy())
    assert close(O.relu_bwd(x, g), X.grad.numpy())


@pytest.mark.parametrize("B,D", [(5, 1), (64, 10), (33, 48)])
def test_mse_forward_and_the_literal_backward_scale(B, D):
    """operators.rs:230-254: the loss is the mean over ALL elements, the gradient (p - t) * 2 / B -- D times the analytic one (D4)"""
    rng = np.random.default_rng(B + D)
    p, t = rng.standard_normal((B, D)), rng.standard_normal((B, D))
    P = t64(p, True)
    loss = ((P - t64(t)) ** 2).mean()
    loss.backward()
    assert close(O.mse_fwd(p, t), loss.detach().numpy().reshape(1, 1))
    up = np.array([[1.0]])
    assert close(O.mse_bwd(p, t, up), D * P.grad.numpy())
    assert close(O.mse_bwd(p, t, np.array([[0.25]])), 0.25 * D * P.grad.numpy())


def test_linear_backward_and_the_one_dimensional_bias_gradient():
    """operators.rs:148-163: dX = G W^T, dW = X^T G, db = sum over rows, shape [O]"""
    rng = np.random.default_rng(3)
    x, w, b, g = rng.standard_normal((17, 9)), rng.standard_normal((9, 13)), rng.standard_normal((1, 13)), rng.standard_normal((17, 13))
    X, Wt, Bt = t64(x, True), t64(w, True), t64(b, True)
    Y = X @ Wt + Bt
    Y.backward(t64(g))
    assert close(O.linear_fwd(x, w, b), Y.detach().numpy())
    dx, dw, db = O.linear_bwd(x, w, g)
    assert close(dx, X.grad.numpy()) and close(dw, Wt.grad.numpy())
    assert db.shape == (13,) and close(db, Bt.grad.numpy().reshape(13))


@pytest.mark.parametrize("B,H,W,co,k,s", [(1, 7, 7, 3, 3, 1), (4, 12, 10, 6, 5, 1), (3, 9, 11, 2, 3, 2)])
def test_conv2d_single_channel_forward_and_backward(B, H, W, co, k, s):
    """operators.rs:350-449 with one input channel and no padding, where the literal lowering IS a convolution: the output in NHWC,
    dx (for B > 1 the per-sample generalisation of SURVEY A.7), dw as [k, k, Co], db as [Co]"""
    rng = np.random.default_rng(B * H + k)
    im, filt, bias = rng.standard_normal((B, 1, H, W)), rng.standard_normal((k, k, co)), rng.standard_normal((co,))
    out, col2d = O.conv2d_fwd(im, filt, bias, 1, k, s, 0)
    g = rng.standard_normal(out.shape)
    dx, dw, db = O.conv2d_bwd(im, filt, col2d, g, 1, k, s, 0, literal=False)
    IM, Fw, Bs = t64(im, True), t64(np.transpose(filt, (2, 0, 1))[:, None], True), t64(bias, True)   # torch: [Co, 1, k, k]
    Y = torch.nn.functional.conv2d(IM, Fw, Bs, stride=s)                                               # NCHW
    Y.backward(t64(np.transpose(g, (0, 3, 1, 2))))
    assert close(out, np.transpose(Y.detach().numpy(), (0, 2, 3, 1)))
    ho, wo = out.shape[1], out.shape[2]
    hh, ww = (ho - 1) * s + k, (wo - 1) * s + k          # col2im's extent (functional_ndarray.rs:107-138): rows / columns the strides never reach are not there
    assert dx.shape == (B, 1, hh, ww) and close(dx, IM.grad.numpy()[:, :, :hh, :ww])
    assert np.abs(IM.grad.numpy()[:, :, hh:, :]).max(initial=0.0) == 0.0 and np.abs(IM.grad.numpy()[:, :, :, ww:]).max(initial=0.0) == 0.0
    assert close(dw, np.transpose(Fw.grad.numpy()[:, 0], (1, 2, 0))) and close(db, Bs.grad.numpy())


def test_conv2d_literal_backward_equals_the_generalisation_at_batch_one():
    """D1: the literal col2im call only works for B = 1, where it must equal the per-sample generalisation the device implements"""
    rng = np.random.default_rng(9)
    im, filt, bias = rng.standard_normal((1, 1, 8, 8)), rng.standard_normal((3, 3, 4)), rng.standard_normal((4,))
    out, col2d = O.conv2d_fwd(im, filt, bias, 1, 3, 1, 0)
    g = rng.standard_normal(out.shape)
    for a, b in zip(O.conv2d_bwd(im, filt, col2d, g, 1, 3, 1, 0, literal=True), O.conv2d_bwd(im, filt, col2d, g, 1, 3, 1, 0, literal=False)):
        assert np.array_equal(a, b)


def test_sgd_step_and_gradient_accumulation():
    """optim.rs:25-34 (w -= g * lr, g broadcast onto w) and tensor.rs:154-163 (None -> copy, Some -> +=)"""
    rng = np.random.default_rng(4)
    w, g = rng.standard_normal((1, 6)), rng.standard_normal((6,))
    Wt = t64(w, True)
    Wt.grad = t64(g).reshape(1, 6)
    torch.optim.SGD([Wt], lr=0.05).step()
    assert close(O.sgd_step(w.copy(), g, 0.05), Wt.detach().numpy())
    acc = O.accumulate(None, g)
    acc = O.accumulate(acc, g)
    assert close(acc, 2 * g)


@pytest.mark.parametrize("act", ["relu", "sigmoid", None])
def test_training_loop_tracks_torch_autograd_step_for_step(act):
    """examples/fit_sine.rs:64-93 on a three-layer MLP, f64: every loss and every final parameter against a torch loop that
    reports the mean loss and differentiates sum((p - t)^2) / B -- the literal MSE gradient of operators.rs:247-249"""
    rng = np.random.default_rng(5)
    widths, B, lr, steps = [7, 12, 9, 4], 25, 1e-2, 6
    ws = [rng.uniform(-0.5, 0.5, (a, b)) for a, b in zip(widths[:-1], widths[1:])]
    bs = [rng.uniform(-0.5, 0.5, (1, b)) for b in widths[1:]]
    x, t = rng.uniform(-1, 1, (B, widths[0])), rng.uniform(-1, 1, (B, widths[-1]))
    want, layers = O.mlp_train(x, t, ws, bs, act, lr, steps, np.float64)
    Ws, Bs = [t64(w, True) for w in ws], [t64(b, True) for b in bs]
    X, T = t64(x), t64(t)
    f = {"relu": torch.relu, "sigmoid": torch.sigmoid, None: None}[act]
    got = []
    for _ in range(steps):
        h = X
        for i, (w, b) in enumerate(zip(Ws, Bs)):
            h = h @ w + b
            if f is not None and i + 1 < len(Ws):
                h = f(h)
        d2 = (h - T) ** 2
        got.append(float(d2.mean().detach()))
        (d2.sum() / B).backward()
        with torch.no_grad():
            for p in Ws + Bs:
                p -= lr * p.grad
                p.grad = None
    assert max(abs(a - b) / abs(b) for a, b in zip(want, got)) < 1e-12
    for l, w, b in zip(layers, Ws, Bs):
        assert close(l.w.data, w.detach().numpy()) and close(l.b.data, b.detach().numpy())


def test_conv2d_three_input_channels_share_the_filter():
    """D5 (operators.rs:362-376): the [k, k, Co] filter is broadcast over the input channels -- a convolution whose weight is the same
    for every channel; the forward, dx and db follow torch with that weight (dw keeps the reference's own [KK, C] regrouping, which
    is a convolution gradient only for one channel and is pinned by shape in the reference's test_conv2d instead)"""
    rng = np.random.default_rng(12)
    B, C, H, W, co, k = 2, 3, 9, 8, 4, 3
    im, filt, bias = rng.standard_normal((B, C, H, W)), rng.standard_normal((k, k, co)), rng.standard_normal((co,))
    out, col2d = O.conv2d_fwd(im, filt, bias, C, k, 1, 0)
    g = rng.standard_normal(out.shape)
    dx, dw, db = O.conv2d_bwd(im, filt, col2d, g, C, k, 1, 0, literal=False)
    IM, Bs = t64(im, True), t64(bias, True)
    Wt = t64(np.broadcast_to(np.transpose(filt, (2, 0, 1))[:, None], (co, C, k, k)).copy())
    Y = torch.nn.functional.conv2d(IM, Wt, Bs)
    Y.backward(t64(np.transpose(g, (0, 3, 1, 2))))
    assert close(out, np.transpose(Y.detach().numpy(), (0, 2, 3, 1)))
    assert close(dx, IM.grad.numpy()) and close(db, Bs.grad.numpy())
    assert dw.shape == (k, k, co)
