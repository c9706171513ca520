"""GPU parity tests of the host engine (libagrs_engine.so over libagrs_b200.so) against the oracle.

The first half are the reference's own 17 unit tests, ported line by line to the front end (file:line in each
docstring); small-integer data, so they are exact.  The second half compares whole training loops with the
oracle on seeded inputs; tolerance: max relative error <= 1e-4 per tensor (BASELINE north star), in practice ~1e-6.
Run with: pytest -m gpu (under gpurun)."""
import numpy as np
import pytest

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import _capi as K
from autograd_rs_b200 import frontend as F
from autograd_rs_b200 import workloads as W
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

TOL = 1e-4  # BASELINE.json north star: max relative error on outputs and gradients


@pytest.fixture(scope="module")
def dev():
    d = F.Device(0)
    yield d
    d.close()


def T(dev, a):
    return F.Tensor.from_numpy(dev, np.array(a, dtype=np.float32))


# ====================================================================== the reference's unit tests
def test_relu_on_mat(dev, kats):
    """operators.rs:486-506"""
    k = kats["relu_on_mat"]
    x = T(dev, k["x"])
    res = F.ReLU().forward([x])
    assert res.requires_grad
    assert np.array_equal(res.data(), np.array(k["y"], np.float32))


def test_linear(dev, kats):
    """operators.rs:508-521"""
    k = kats["linear"]
    res = F.Linear().forward([T(dev, k["x"]), T(dev, k["w"]), T(dev, k["b"])])
    assert np.array_equal(res.data(), np.array(k["y"], np.float32))


@pytest.mark.parametrize("name", ["im2col_nopad", "im2col_pad"])
def test_im2col(dev, kats, name):
    """operators.rs:523-566"""
    k = kats[name]
    im = T(dev, np.arange(*k["im_range"], dtype=np.float32).reshape(k["im_shape"]))
    col = F.Conv2D.im2col(im, k["k"], k["stride"], k["pad"])
    assert col.shape() == k["col_shape"]
    assert np.array_equal(col.data(), np.array(k["col"], np.float32).reshape(k["col_shape"]))


def test_im2col_rejects_non_bchw(dev):
    """functional_ndarray.rs:100-103: Err("Expected image of shape BCHW...")"""
    with pytest.raises(F.Panic, match="BCHW"):
        F.Conv2D.im2col(T(dev, np.zeros((3, 4, 4))), 2, 1, 0)


def test_conv2d(dev, kats):
    """operators.rs:568-602: forward all 9.0 (NHWC), backward gradient shapes equal input shapes"""
    k = kats["conv2d_ones"]
    conv = F.Conv2D(k["c_in"], k["c_out"], k["k"], k["stride"], k["pad"])
    xs = [F.Tensor.ones(dev, k["im_shape"]), F.Tensor.ones(dev, k["filt_shape"]), F.Tensor.ones(dev, k["bias_shape"])]
    res = conv.forward(xs)
    assert res.shape() == k["out_shape"]
    assert np.all(res.data() == k["out_value"])
    # the test calls backward with the op inputs + the cached col, which lives in the graph node
    col = F.Conv2D.im2col(xs[0], k["k"], k["stride"], k["pad"])
    col.reshape([col.shape()[0] * col.shape()[1], col.shape()[2]])
    grads = conv.backward(xs + [col], F.Tensor.ones(dev, k["out_shape"]))
    assert [g.shape() for g in grads] == k["bwd_shapes"]


def test_single_node_graph(dev, kats):
    """autograd.rs:94-119"""
    k = kats["single_node_graph"]
    x = T(dev, k["x"])
    res = F.ReLU().forward([x.clone()])
    assert np.array_equal(res.data(), np.array(k["y"], np.float32))
    name, nvars, nparents = res.graph_info()
    assert name == "ReLU" and nvars == 1 and nparents == 0
    res.backward()
    assert np.array_equal(x.grad().data(), np.array(k["x_grad"], np.float32))


def test_simple_graph(dev, kats):
    """autograd.rs:121-171: ReLU -> Linear; x.grad, W.grad, b.grad (1-D)"""
    k = kats["simple_graph"]
    x = T(dev, k["x"])
    res = F.ReLU().forward([x.clone()])
    w, b = T(dev, k["w"]), T(dev, k["b"])
    res = F.Linear().forward([res, w.clone(), b.clone()])
    assert np.array_equal(res.data(), np.array(k["y"], np.float32))
    assert res.graph_info() == ("Linear", 3, 1)
    res.backward()
    assert np.array_equal(x.grad().data(), np.array(k["x_grad"], np.float32))
    assert np.array_equal(w.grad().data(), np.array(k["w_grad"], np.float32))
    assert b.grad().shape() == [2] and np.array_equal(b.grad().data(), np.array(k["b_grad"], np.float32))


def test_basic_from(dev, kats):
    """tensor.rs:464-472"""
    k = kats["basic_from"]
    t = T(dev, np.zeros(k["shape"]))
    assert t.ndim() == k["ndim"] and t.len() == k["len"] and t.shape() == k["shape"] and not t.is_empty()


def test_clone_tensor(dev):
    """tensor.rs:474-492: a clone points at the same storage"""
    a = F.Tensor.zeros(dev, (2, 2))
    b = a.clone()
    assert a.data_ptr() == b.data_ptr()
    b.fill(3.0)
    assert np.all(a.data() == 3.0)


def test_reshape(dev, kats):
    """tensor.rs:493-502: reshape through a clone is seen by the original"""
    k = kats["reshape_alias"]
    a = F.Tensor.zeros(dev, k["from"])
    b = a.clone()
    b.reshape(k["to"])
    assert a.shape() == k["to"]
    with pytest.raises(F.Panic):
        a.reshape([5])


def test_adds(dev, kats):
    """tensor_arithmetic.rs:348-365"""
    k = kats["arith_adds"]
    a, b = T(dev, k["a"]), T(dev, k["b"])
    a_ptr = a.data_ptr()
    a = a.add_(b)  # a + &b: in place
    assert np.array_equal(a.data(), np.array(k["a_plus_b"], np.float32)) and a.data_ptr() == a_ptr
    c = a.clone()
    c.fill(1.0)
    assert np.all(a.data() == 1.0)  # visible through the alias
    d = a + b  # &a + &b: new storage
    assert d.data_ptr() != a.data_ptr() and np.all(d.data() == 2.0) and np.all(a.data() == 1.0)
    assert d.grad() is None


def test_subs(dev):
    """tensor_arithmetic.rs:366-383"""
    a, b = T(dev, [[1, 2], [3, 4]]), F.Tensor.ones(dev, (2, 2))
    ptr = a.data_ptr()
    a.sub_(b)
    assert np.array_equal(a.data(), np.array([[0, 1], [2, 3]], np.float32)) and a.data_ptr() == ptr
    d = a - b
    assert d.data_ptr() != ptr and np.array_equal(d.data(), np.array([[-1, 0], [1, 2]], np.float32))
    a -= b
    assert np.array_equal(a.data(), d.data()) and a.data_ptr() == ptr


def test_scalar_ops(dev, kats):
    """tensor_arithmetic.rs:385-399"""
    k = kats["arith_scalar"]
    a = T(dev, k["a"])
    assert np.array_equal((a * 2).data(), np.array(k["mul2"], np.float32))
    assert np.array_equal((a / 2).data(), np.array(k["div2"], np.float32))
    assert np.array_equal((3.0 - a).data(), np.array(k["three_minus_a"], np.float32))
    a *= 3
    assert np.array_equal(a.data(), np.array(k["mul_assign3"], np.float32))


def test_mul(dev):
    """tensor_arithmetic.rs:401-437: &a * &b is new storage; a * &b keeps a's storage address"""
    a, b = T(dev, [[1, 2], [3, 4]]), T(dev, [[2, 2], [2, 2]])
    c = a * b
    assert c.data_ptr() != a.data_ptr() and np.array_equal(c.data(), np.array([[2, 4], [6, 8]], np.float32))
    ptr = a.data_ptr()
    a = a.mul_(b)
    assert a.data_ptr() == ptr and np.array_equal(a.data(), c.data())
    assert a.equals(c) and not a.equals(b)
    assert np.array_equal((-a).data(), -c.data())  # Neg is todo!() in the reference (Q4): implemented


def test_layers_mlp(dev):
    """layers.rs:194-206: Linear(20,10) -> ReLU -> Linear(10,1); backward; x.grad != None"""
    dev.seed(7)
    x = F.Tensor.ones(dev, (1, 20))
    l1, act, l2 = F.nn.Linear(dev, 20, 10), F.nn.ReLU(), F.nn.Linear(dev, 10, 1)
    res = l2.forward([act.forward([l1.forward([x.clone()])])])
    res.backward()
    assert x.grad() is not None and x.grad().shape() == [1, 20]


def test_layers_conv(dev):
    """layers.rs:207-217: Conv2D(3,16,3) on 1x3x24x24 -> ReLU -> backward; x.grad != None (B=1, C_in=3)"""
    dev.seed(8)
    x = F.Tensor.ones(dev, (1, 3, 24, 24))
    conv, act = F.nn.Conv2D(dev, 3, 16, 3), F.nn.ReLU()
    res = act.forward([conv.forward([x.clone()])])
    assert res.shape() == [1, 22, 22, 16]
    res.backward()
    assert x.grad() is not None and x.grad().shape() == [1, 3, 24, 24]


def test_model_mlp_shapes(dev, kats):
    """model.rs:31-62"""
    k = kats["mlp_param_shapes"]
    dev.seed(9)
    model = F.nn.Sequential([F.nn.Linear(dev, 20, 10), F.nn.ReLU(), F.nn.Linear(dev, 10, 1)])
    assert [p.shape() for p in model.parameters()] == k["shapes"]
    out = model.forward([F.Tensor.ones(dev, (1, 20))])
    assert out.shape() == k["out_shape"]
    # init bounds (init.rs:19-49; layers.rs:82-89)
    w = model.parameters()[0].data()
    assert np.abs(w).max() <= W.kaiming_bound((20, 10)) and np.abs(w).max() > 0.5 * W.kaiming_bound((20, 10))
    assert np.abs(model.parameters()[1].data()).max() <= 1 / np.sqrt(10)


# ====================================================================== semantics the reference leaves unpinned
def test_backward_needs_graph(dev):
    """tensor.rs:337: expect("Variable has no attached graph")"""
    with pytest.raises(F.Panic, match="no attached graph"):
        F.Tensor.ones(dev, (2, 2)).backward()


def test_dot_panics(dev):
    """tensor.rs:303-305"""
    with pytest.raises(F.Panic, match="1d/2d matmul"):
        F.Tensor.ones(dev, (2, 2, 2)).dot(F.Tensor.ones(dev, (2, 2)))
    with pytest.raises(F.Panic, match="not compatible"):
        F.Tensor.ones(dev, (2, 3)).dot(F.Tensor.ones(dev, (2, 3)))


def test_transpose_views_feed_gemm(dev):
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((5, 7)).astype(np.float32), rng.standard_normal((5, 3)).astype(np.float32)
    ta = T(dev, a)
    ta.t()  # in place on the shared storage (tensor.rs:165-169)
    assert ta.shape() == [7, 5] and not ta.is_contiguous()
    assert O.rel_err(ta.dot(T(dev, b)).data(), a.T @ b) < 1e-6
    assert np.array_equal(ta.data(), a.T)
    tc = T(dev, b).t_clone()
    assert tc.shape() == [3, 5] and np.array_equal(tc.data(), b.T)


def test_accumulate_twice_and_zero_grad(dev):
    """tensor.rs:154-163 (`+=` is todo!() at HEAD, defect D2) and model.rs:14-21"""
    x = T(dev, [[1, -2], [3, 4]])
    for _ in range(2):
        F.ReLU().forward([x.clone()]).backward()
    assert np.array_equal(x.grad().data(), np.array([[2, 0], [2, 2]], np.float32))
    x.zero_grad()
    assert np.all(x.grad().data() == 0)
    # SGD broadcasts a [O] gradient onto a [1,O] bias and zero_grad then makes the grad [1,O]
    w, b = T(dev, np.ones((2, 3))), T(dev, np.ones((1, 3)))
    F.Linear().forward([T(dev, [[1, 2]]), w.clone(), b.clone()]).backward()
    assert b.grad().shape() == [3]
    F.nn.SGD([w, b], 0.5).step()
    assert np.array_equal(b.data(), np.full((1, 3), 0.5, np.float32))
    assert np.array_equal(w.data(), np.array([[0.5] * 3, [0.0] * 3], np.float32))
    model = F.nn.Sequential([F.nn.Linear(dev, 2, 3, w=w, bias=b)])
    model.zero_grad()
    assert b.grad().shape() == [1, 3] and np.all(b.grad().data() == 0)
    with pytest.raises(F.Panic, match="None"):
        F.nn.SGD([T(dev, [[1.0]])], 0.1).step()  # optim.rs:30 unwrap on a None grad


def test_matmul_backward_literal_and_fixed(dev):
    """operators.rs:183-184 (defect D3): literal dx = g.w, dw = g^T.x; the fixed flavour behind a switch"""
    rng = np.random.default_rng(3)
    x, w, g = (rng.standard_normal((6, 6)).astype(np.float32) for _ in range(3))
    for literal in (True, False):
        dev.set_matmul_backward_literal(literal)
        dx, dw = F.MatMul().backward([T(dev, x), T(dev, w)], T(dev, g))
        wdx, wdw = O.matmul_bwd(x, w, g, literal=literal)
        assert O.rel_err(dx.data(), wdx) < 1e-6 and O.rel_err(dw.data(), wdw) < 1e-6
    dev.set_matmul_backward_literal(True)


@pytest.mark.parametrize("op,ofwd,obwd", [("sigmoid", O.sigmoid_fwd, O.sigmoid_bwd), ("relu", O.relu_fwd, O.relu_bwd)])
def test_activation_ops(dev, op, ofwd, obwd):
    rng = np.random.default_rng(4)
    x, g = rng.standard_normal((33, 17)).astype(np.float32) * 3, rng.standard_normal((33, 17)).astype(np.float32)
    o = {"sigmoid": F.Sigmoid, "relu": F.ReLU}[op]()
    assert O.rel_err(o.forward([T(dev, x)]).data(), ofwd(x)) < 1e-6
    assert O.rel_err(o.backward([T(dev, x)], T(dev, g))[0].data(), obwd(x, g)) < 1e-6


def test_mse_and_mean_ops(dev):
    rng = np.random.default_rng(5)
    p, t = rng.standard_normal((64, 10)).astype(np.float32), rng.standard_normal((64, 10)).astype(np.float32)
    loss = F.MeanSquaredError().forward([T(dev, p), T(dev, t)])
    assert loss.shape() == [1, 1] and abs(loss.item() - O.mse_fwd(p, t)[0, 0]) < 1e-6 * abs(loss.item())
    g = F.MeanSquaredError().backward([T(dev, p), T(dev, t)], F.Tensor.ones(dev, (1, 1)))
    assert len(g) == 1  # the target gets no gradient (autograd.rs:73 zips to the shorter list)
    assert O.rel_err(g[0].data(), O.mse_bwd(p, t, np.ones((1, 1), np.float32))) < 1e-6  # literal 2/B scaling (D4)
    m = F.Mean().forward([T(dev, p)])
    assert m.shape() == [1] and abs(m.item() - p.mean()) < 1e-6
    gm = F.Mean().backward([T(dev, p)], F.Tensor.ones(dev, (1,)))
    assert O.rel_err(gm[0].data(), O.mean_bwd(p, np.ones(1, np.float32))) < 1e-6


@pytest.mark.parametrize("B,C,H,Wd,co,k,s,p", [(1, 3, 9, 8, 4, 3, 1, 0),   # C_in = 3: the literal shared-filter / [KK, C] grouping quirks (D5)
                                              (1, 2, 7, 7, 5, 2, 2, 1),   # stride 2, pad 1: dx keeps the padded size (TODO at operators.rs:427)
                                              (5, 3, 10, 9, 6, 3, 1, 1),  # B > 1: the per-sample generalisation of col2im (D1)
                                              (64, 1, 28, 28, 6, 5, 1, 0)])  # the LeNet geometry
def test_conv2d_values_match_literal_oracle(dev, B, C, H, Wd, co, k, s, p):
    """Conv2D forward AND backward values against the oracle's literal restatement of operators.rs:350-449, quirks included
    (the literal lowering: im2col + GEMM with the col matrix cached in the node)"""
    dev.set_conv_implicit(False)
    try:
        _conv2d_literal_case(dev, B, C, H, Wd, co, k, s, p)
    finally:
        dev.set_conv_implicit(True)


def _conv2d_literal_case(dev, B, C, H, Wd, co, k, s, p):
    rng = np.random.default_rng(B + C + H + k + p)
    im = rng.standard_normal((B, C, H, Wd)).astype(np.float32)
    filt = rng.standard_normal((k, k, co)).astype(np.float32)
    bias = rng.standard_normal((co,)).astype(np.float32)
    conv = F.Conv2D(C, co, k, s, p)
    X, Fl, Bi = T(dev, im), T(dev, filt), T(dev, bias)
    out = conv.forward([X.clone(), Fl.clone(), Bi.clone()])
    want_out, col2d = O.conv2d_fwd(im, filt, bias, C, k, s, p)
    assert out.shape() == list(want_out.shape)
    assert O.rel_err(out.data(), want_out) < 1e-5
    assert out.graph_info() == ("Conv2D", 4, 0)  # the cached col is the 4th variable (operators.rs:386-388)
    g = rng.standard_normal(want_out.shape).astype(np.float32)
    col_t = T(dev, col2d)
    dx, dw, db = conv.backward([X, Fl, Bi, col_t], T(dev, g))
    wdx, wdw, wdb = O.conv2d_bwd(im, filt, col2d, g, C, k, s, p, literal=False)
    assert dx.shape() == list(wdx.shape) and dw.shape() == [k, k, co] and db.shape() == [co]
    assert O.rel_err(dx.data(), wdx) < 1e-5 and O.rel_err(dw.data(), wdw) < 1e-5 and O.rel_err(db.data(), wdb) < 1e-5
    if B == 1:  # at B = 1 the generalisation IS the literal code
        ldx, ldw, ldb = O.conv2d_bwd(im, filt, col2d, g, C, k, s, p, literal=True)
        assert np.array_equal(ldx, wdx) and np.array_equal(ldw, wdw)


def test_identity_mean_graph_and_broadcast_arithmetic(dev):
    """Identity shares storage and passes the gradient (operators.rs:274-286); Mean's backward is ones/size (:267-271);
    ndarray broadcasting in the arithmetic: [B,O] op [O], [B,O] op [1,O], [B,O] op [B,1], [B,O] op [1,1]"""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((6, 5)).astype(np.float32)
    X = T(dev, x)
    y = F.Identity().forward([X.clone()])
    assert y.data_ptr() == X.data_ptr() and y.graph_info()[0] == "Identity"
    m = F.Mean().forward([y])
    m.backward()
    # Identity's output is a CLONE of its input: same data cell AND same grad cell (operators.rs:279).  The walk accumulates
    # 1/size into it once as Mean's input and once more as Identity's input -- the literal reference gives 2/size, and so
    # does the oracle.
    OX = O.OTensor(x)
    O.OpMean().forward([O.OpIdentity().forward([OX.clone()])]).backward()
    assert np.allclose(OX.grad, 2.0 / 30) and np.allclose(X.grad().data(), OX.grad, rtol=1e-7)
    a = T(dev, x)
    for other in (rng.standard_normal(5), rng.standard_normal((1, 5)), rng.standard_normal((6, 1)), rng.standard_normal((1, 1))):
        o = other.astype(np.float32)
        assert np.array_equal((a + T(dev, o)).data(), x + o)
        assert np.array_equal((a - T(dev, o)).data(), x - o)
        assert np.array_equal((a * T(dev, o)).data(), x * o)
    with pytest.raises(F.Panic, match="broadcast"):
        a + T(dev, np.zeros((4,), np.float32))  # ndarray: could not broadcast [4] onto [6, 5]
    b = T(dev, x)
    b.add_(T(dev, np.ones(5, np.float32)))      # in place with a broadcast right-hand side keeps the address
    assert np.array_equal(b.data(), x + 1)


# ====================================================================== whole training loops vs the oracle
def run_device_mlp(dev, x, t, ws, bs, act, lr, steps):
    model = W.build_mlp(dev, ws, bs, act)
    tr = W.Trainer(model, lr)
    X, Tt = T(dev, x), T(dev, t)
    Tt.requires_grad = False
    losses = [tr.step(X, Tt).item() for _ in range(steps)]
    return losses, [p.data() for p in model.parameters()], X


def test_fit_sine_tracks_oracle_step_for_step(dev):
    """C1, examples/fit_sine.rs:17-102 (2000 samples, Linear(3,1), lr 1e-3); W=b=0 init as SURVEY Appendix E"""
    x, y = O.fit_sine_data(2000, np.float32)
    steps = 300
    want, layers = O.mlp_train(x, y, [np.zeros((3, 1))], [np.zeros((1, 1))], None, 1e-3, steps, np.float32)
    got, params, _ = run_device_mlp(dev, x, y, [np.zeros((3, 1), np.float32)], [np.zeros((1, 1), np.float32)], None, 1e-3, steps)
    assert max(abs(a - b) / abs(b) for a, b in zip(got, want)) < TOL
    assert abs(got[0] - 0.49974999) < 1e-6 and abs(got[1] - 0.44477817) < 1e-6 and abs(got[99] - 0.31669122) < 2e-6
    # the trained bias is ~1e-10 (pure rounding residue of a symmetric problem): compare both parameters on the weights' scale
    scale = float(np.abs(layers[0].w.data).max())
    assert np.abs(params[0] - layers[0].w.data).max() < TOL * scale and np.abs(params[1] - layers[0].b.data).max() < TOL * scale


@pytest.mark.parametrize("act,width,rows,layers", [("relu", 256, 512, 3), ("sigmoid", 128, 300, 4), ("relu", 520, 1000, 2)])
def test_mlp_training_tracks_oracle(dev, act, width, rows, layers):
    """C2/C5 in miniature: every loss value and every final weight within 1e-4 of the oracle, f32 vs f32;
    the device is also no further from the f64 restatement than the f32 oracle is (x4)."""
    ws, bs = W.mlp_init(layers, width, seed=11)
    x, t = W.mlp_batch(rows, width, width, seed=12)
    lr, steps = 1e-4, 5
    want, olayers = O.mlp_train(x, t, ws, bs, act, lr, steps, np.float32)
    want64, olayers64 = O.mlp_train(x, t, ws, bs, act, lr, steps, np.float64)
    got, params, X = run_device_mlp(dev, x, t, ws, bs, act, lr, steps)
    assert max(abs(a - b) / abs(b) for a, b in zip(got, want)) < TOL
    oparams = [p.data for l in olayers for p in l.parameters()]
    oparams64 = [p.data for l in olayers64 for p in l.parameters()]
    for g, w32, w64 in zip(params, oparams, oparams64):
        assert O.rel_err(g, w32) < TOL
        assert O.rel_err(g, w64) < 4 * O.rel_err(w32, w64) + 1e-6
    # x.grad keeps accumulating over the 5 iterations (the input handle is reused and never zeroed, fit_sine.rs:65)
    assert X.grad() is not None and X.grad().shape() == list(x.shape)


def test_gradients_match_oracle_one_step(dev):
    """outputs AND gradients of one forward/backward (north star tolerance), tcgen05 GEMM shapes"""
    ws, bs = W.mlp_init(2, 384, seed=21)
    x, t = W.mlp_batch(640, 384, 384, seed=22)
    model = W.build_mlp(dev, ws, bs, "relu")
    X, Tt = T(dev, x), T(dev, t)
    Tt.requires_grad = False
    pred = model.forward([X.clone()])
    loss = F.nn.MeanSquaredError().forward([pred, Tt.clone()])
    loss.backward()
    ol = [O.OLinearLayer(w, b) for w, b in zip(ws, bs)]
    OX, OT = O.OTensor(x), O.OTensor(t, requires_grad=False)
    h = O.OpReLU().forward([ol[0].forward(OX.clone())])
    opred = ol[1].forward(h)
    oloss = O.OpMSE().forward([opred, OT.clone()])
    oloss.backward()
    assert O.rel_err(pred.data(), opred.data) < TOL
    assert O.rel_err(X.grad().data(), OX.grad) < TOL
    for p, op in zip(model.parameters(), [q for l in ol for q in l.parameters()]):
        assert p.grad().shape() == list(op.grad.shape)
        assert O.rel_err(p.grad().data(), op.grad) < TOL


def test_lenet_training_tracks_oracle(dev):
    """C3 in miniature (B=32): Conv2D(1,6,5) -> ReLU -> Flatten -> Linear(3456,10) -> MSE; conv backward at B>1
    is the per-sample generalisation of the literal code (defect D1)."""
    rng = np.random.default_rng(31)
    B = 32
    im = rng.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)
    t = rng.uniform(0, 1, (B, 10)).astype(np.float32)
    fb = W.kaiming_bound((5, 5, 6))
    filt = rng.uniform(-fb, fb, (5, 5, 6)).astype(np.float32)
    cbias = rng.uniform(-0.1, 0.1, (6,)).astype(np.float32)
    w = rng.uniform(-0.05, 0.05, (3456, 10)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (1, 10)).astype(np.float32)
    lr, steps = 1e-3, 4
    model = W.build_lenet(dev, filt, cbias, w, b)
    tr = W.Trainer(model, lr)
    X, Tt = T(dev, im), T(dev, t)
    Tt.requires_grad = False
    got = [tr.step(X, Tt).item() for _ in range(steps)]
    # oracle loop
    conv, lin = O.OConvLayer(filt.copy(), cbias.copy(), 1), O.OLinearLayer(w.copy(), b.copy())
    params = conv.parameters() + lin.parameters()
    OX, OT = O.OTensor(im), O.OTensor(t, requires_grad=False)
    want = []
    for _ in range(steps):
        h = O.OpReLU().forward([conv.forward(OX.clone())])
        h = O.OpFlatten().forward([h])
        loss = O.OpMSE().forward([lin.forward(h), OT.clone()])
        want.append(float(loss.data[0, 0]))
        loss.backward()
        O.sgd(params, lr)
        O.model_zero_grad(params)
    assert max(abs(a - c) / abs(c) for a, c in zip(got, want)) < TOL
    for p, op in zip(model.parameters(), params):
        assert O.rel_err(p.data(), op.data) < TOL
    assert O.rel_err(X.grad().data(), OX.grad) < TOL


def test_no_host_sync_inside_a_step(dev):
    """forward .. zero_grad enqueue only: the step returns while the device is still busy (no host round trip)."""
    import time
    # big enough that the device needs ~10 ms per step while the host needs well under 1 ms to enqueue it
    ws, bs = W.mlp_init(3, 4096, seed=41)
    x, t = W.mlp_batch(8192, 4096, 4096, seed=42)
    model = W.build_mlp(dev, ws, bs, "relu")
    tr = W.Trainer(model, 1e-6)
    X, Tt = T(dev, x), T(dev, t)
    Tt.requires_grad = False
    tr.reserve_pool(X, Tt)  # the memory pool is sized once: afterwards a step never calls into the driver for memory
    for _ in range(2):      # warm-up: workspaces reach their steady state
        loss = tr.step(X, Tt)
    dev.sync()
    t0 = time.perf_counter()
    for _ in range(3):
        loss = tr.step(X, Tt)
    t_enqueue = time.perf_counter() - t0
    dev.sync()
    t_total = time.perf_counter() - t0
    assert t_enqueue < 0.5 * t_total, (t_enqueue, t_total)
    assert np.isfinite(loss.item())


def test_async_loss_read_matches_blocking_read(dev):
    """Tensor.read_async + Event.wait (agrs_download_async / agrs_event_wait): the loss of every step reaches the host, one step
    late, with the values a blocking item() gives -- and the host is not held while the device works."""
    import time
    ws, bs = W.mlp_init(3, 2048, seed=51)
    x, t = W.mlp_batch(4096, 2048, 2048, seed=52)
    X, Tt = T(dev, x), T(dev, t)
    Tt.requires_grad = False
    want = []
    tr = W.Trainer(W.build_mlp(dev, ws, bs, "relu"), 1e-4)
    for _ in range(4):
        want.append(tr.step(X, Tt).item())
    tr = W.Trainer(W.build_mlp(dev, ws, bs, "relu"), 1e-4)   # same start, same data: bit-identical losses
    slots = [dev.pinned_empty((1,)), dev.pinned_empty((1,))]
    evs = [F.Event(dev), F.Event(dev)]
    got = []
    for i in range(4):
        tr.step(X, Tt).read_async(slots[i % 2], evs[i % 2])
        if i > 0:
            evs[(i - 1) % 2].wait()
            got.append(float(slots[(i - 1) % 2][0]))
    dev.sync()
    t0 = time.perf_counter()
    tr.step(X, Tt).read_async(slots[0], evs[0])   # returns while the step is still running
    t_enqueue = time.perf_counter() - t0
    evs[0].wait()
    t_total = time.perf_counter() - t0
    evs[1].wait()
    got.append(float(slots[1][0]))
    assert got == want, (got, want)
    assert t_enqueue < 0.5 * t_total, (t_enqueue, t_total)
    fresh = F.Event(dev)
    fresh.wait()                                   # an event nothing was enqueued behind does not block
    for e in evs + [fresh]:
        e.close()


@pytest.mark.parametrize("unroll", [1, 5])
def test_captured_steps_track_oracle(unroll):
    """fit_sine with the training step recorded as a CUDA graph (agrs_capture_*): 2 ordinary warm-up steps + replays must land
    on the same loss and weights as the oracle after the same number of iterations."""
    d = F.Device(0)
    try:
        x, y = O.fit_sine_data(2000, np.float32)
        model = W.build_mlp(d, [np.zeros((3, 1), np.float32)], [np.zeros((1, 1), np.float32)], None)
        tr = W.Trainer(model, 1e-3)
        X, Tt = T(d, x), T(d, y)
        Tt.requires_grad = False
        tr.capture(X, Tt, unroll=unroll, warmup=2)
        assert tr._graph.kernel_nodes >= 8 * unroll
        replays = 60
        loss = tr.replay(replays)
        total = 2 + replays * unroll
        want, layers = O.mlp_train(x, y, [np.zeros((3, 1))], [np.zeros((1, 1))], None, 1e-3, total, np.float32)
        assert abs(loss.item() - want[-1]) / want[-1] < TOL
        scale = float(np.abs(layers[0].w.data).max())
        got_w = model.parameters()[0].data()
        assert np.abs(got_w - layers[0].w.data).max() < TOL * scale
        # the engine is still usable step by step afterwards, continuing from the replayed state
        l2 = tr.step(X, Tt).item()
        want2, _ = O.mlp_train(x, y, [np.zeros((3, 1))], [np.zeros((1, 1))], None, 1e-3, total + 1, np.float32)
        assert abs(l2 - want2[-1]) / want2[-1] < TOL
        with pytest.raises(F.Panic, match="captured"):
            d.capture_begin()
            try:
                X.data()  # a blocking read inside a capture is refused, not silently wrong
            finally:
                d.capture_end().close()
        tr._graph.close()
    finally:
        d.close()


def test_prefetch_copy_stream(dev):
    """copy-stream upload (input prefetch of the end-to-end loop): data is what was sent once the compute stream has joined"""
    rng = np.random.default_rng(61)
    host = dev.pinned_empty((512, 384))
    t = F.Tensor.zeros(dev, (512, 384))
    for _ in range(3):
        host[...] = rng.standard_normal((512, 384)).astype(np.float32)
        t.prefetch_from_numpy(host)
        dev.prefetch_join()
        doubled = t * 2.0  # a kernel enqueued after the join sees the new data
        assert np.array_equal(doubled.data(), host * 2.0)
    t.copy_from_numpy(host * 0 + 1, blocking=True)
    assert np.all(t.data() == 1.0)


def test_prefetched_tensor_waits_for_its_own_copy(dev):
    """no join: the first reader of a prefetched tensor waits for that tensor's copy (landing event kept with the buffer); a second
    prefetch into a buffer nobody has read yet, a host read and an in-place write right after a prefetch are all ordered behind it"""
    rng = np.random.default_rng(62)
    ha, hb = dev.pinned_empty((1024, 2048)), dev.pinned_empty((1024, 2048))
    a, b = F.Tensor.zeros(dev, (1024, 2048)), F.Tensor.zeros(dev, (1024, 2048))
    for _ in range(3):
        ha[...] = rng.standard_normal(ha.shape).astype(np.float32)
        hb[...] = rng.standard_normal(hb.shape).astype(np.float32)
        a.prefetch_from_numpy(ha)
        b.prefetch_from_numpy(hb)
        s = a + b                       # reads both: waits for both copies
        assert np.array_equal(s.data(), ha + hb)
    keep = ha.copy()
    hb[...] = 7.0
    a.prefetch_from_numpy(ha)
    a.prefetch_from_numpy(hb)           # overwrites the first copy, in order
    assert np.all(a.data() == 7.0)      # host read straight after a prefetch
    a.prefetch_from_numpy(ha)
    a *= 2.0                            # in-place write straight after a prefetch
    assert np.array_equal(a.data(), keep * 2.0)
    x, t = W.mlp_batch(1024, 2048, 2048, seed=63)
    ws, bs = W.mlp_init(2, 2048, seed=64)
    tr = W.Trainer(W.build_mlp(dev, ws, bs, "relu"), 1e-4)
    X, Tt = T(dev, x), T(dev, t)
    Tt.requires_grad = False
    want = tr.step(X, Tt).item()
    tr = W.Trainer(W.build_mlp(dev, ws, bs, "relu"), 1e-4)
    ha[...] = x
    hb[...] = t
    a.prefetch_from_numpy(ha)           # X lands first, the forward pass starts; T is needed by the loss only
    b.requires_grad = False
    b.prefetch_from_numpy(hb)
    assert tr.step(a, b).item() == want


def test_data_parallel_two_gpus_tracks_oracle():
    """N=2 on real GPUs (NCCL over NVLink): tools/dp_parity.py under torchrun; skipped on a single-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "dp_parity.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("width,rows", [(192, 320), (256, 512)])
def test_fused_optimizer_world1_matches_plain_sgd(dev, width, rows, overlap):
    """The fused peer-memory step with a single rank: parameters move into the arena (address changes, values and every
    clone follow), gradients reach the exchange slots (192 wide: scatter kernel after the GEMM; 256 wide: pushed by the
    epilogue of the tcgen05 pair GEMM itself), one kernel per step does the slot sum + SGD + broadcast, and the losses are those
    of the plain path bit for bit (mean over one rank is the identity; the update keeps the two roundings of optim.rs:31-33)."""
    ws, bs = W.mlp_init(3, width, seed=71)
    x, t = W.mlp_batch(rows, width, width, seed=72)
    want, params_plain, _ = run_device_mlp(dev, x, t, ws, bs, "sigmoid", 1e-4, 4)
    model = W.build_mlp(dev, ws, bs, "sigmoid")
    params = model.parameters()
    before = [p.data_ptr() for p in params]
    dp = F.DataParallel(dev, 1, 0, F.comm_unique_id(), params, overlap=True)
    dp.enable_fused(lambda h: [h])
    after = [p.data_ptr() for p in params]
    assert all(a != b for a, b in zip(after, before)) and all(a % 256 == 0 for a in after)
    assert after == sorted(after)  # one arena, registration order
    for p, w0 in zip(params, [a for pair in zip(ws, bs) for a in pair]):
        assert np.array_equal(p.data(), w0)
    tr = W.Trainer(model, 1e-4, dp=dp, fused_overlap=overlap)   # per bucket under backward / one kernel after backward
    X, Tt = T(dev, x), T(dev, t)
    Tt.requires_grad = False
    n0 = dev.launches()
    got = [tr.step(X, Tt).item() for _ in range(4)]
    # against the oracle, not only against the device's own plain path (north star tolerance)
    owant, olayers = O.mlp_train(x, t, ws, bs, "sigmoid", 1e-4, 4, np.float32)
    assert max(abs(a - c) / abs(c) for a, c in zip(got, owant)) < TOL
    for p, q in zip(model.parameters(), [q for l in olayers for q in l.parameters()]):
        assert O.rel_err(p.data(), q.data) < TOL
    # 192 wide: the same kernels in the same order as the plain path -> the same bits.  256 wide: the scattering GEMM never
    # splits K while the plain dW GEMM may (a different, equally valid summation order) -> equal to f32 rounding.
    same = (lambda a, b: np.array_equal(a, b)) if width == 192 else (lambda a, b: np.allclose(a, b, rtol=2e-5, atol=1e-7))
    assert same(np.array(got), np.array(want))
    for p, q in zip(model.parameters(), params_plain):
        assert same(p.data(), q)
    assert [p.data_ptr() for p in model.parameters()] == after  # the parameters stay in the arena
    final = [p.data() for p in model.parameters()]
    dp.close()
    # after close the parameters live in ordinary storage again, values intact
    for p, q in zip(model.parameters(), final):
        assert np.array_equal(p.data(), q)


def test_data_parallel_world1_is_identity(dev):
    """DataParallel with one rank: registration + sync are no-ops on the numbers (the N>1 path runs under torchrun)."""
    ws, bs = W.mlp_init(2, 128, seed=51)
    x, t = W.mlp_batch(256, 128, 128, seed=52)
    want, _, _ = run_device_mlp(dev, x, t, ws, bs, "relu", 1e-4, 3)
    model = W.build_mlp(dev, ws, bs, "relu")
    dp = F.DataParallel(dev, 1, 0, F.comm_unique_id(), model.parameters(), overlap=True)
    tr = W.Trainer(model, 1e-4, dp=dp)
    X, Tt = T(dev, x), T(dev, t)
    Tt.requires_grad = False
    got = [tr.step(X, Tt).item() for _ in range(3)]
    dp.close()
    assert got == want


@pytest.mark.parametrize("act", ["relu", "sigmoid"])
def test_backward_colsum_fusion_changes_no_bit(dev, act):
    """f2: with the fusion on (default) db comes from the backward kernel that produced the gradient; off, from a separate pass.
    Losses and every parameter after 3 steps are identical bit for bit, and both track the oracle."""
    ws, bs = W.mlp_init(3, 256, seed=81)
    x, t = W.mlp_batch(512, 256, 256, seed=82)
    runs = []
    for fused in (True, False):
        dev.set_fusions(backward_colsum=fused)
        try:
            n0 = dev.launches()
            losses, params, _ = run_device_mlp(dev, x, t, ws, bs, act, 1e-4, 3)
            runs.append((losses, params, dev.launches() - n0))
        finally:
            dev.set_fusions(backward_colsum=True)
    assert runs[0][0] == runs[1][0]
    for p, q in zip(runs[0][1], runs[1][1]):
        assert np.array_equal(p, q)
    assert runs[0][2] < runs[1][2]   # three column-sum launches per step fewer
    want, layers = O.mlp_train(x, t, ws, bs, act, 1e-4, 3, np.float32)
    assert max(abs(a - c) / abs(c) for a, c in zip(runs[0][0], want)) < TOL


# ====================================================================== examples and model serialisation
@pytest.mark.parametrize("script,args,needle", [("fit_sine.py", ["--epochs", "300"], "Result: y ="), ("fit_sine.py", ["--epochs", "300", "--graph"], "Result: y ="),
                                                ("example.py", [], "identical = True")])
def test_example_ports_run(script, args, needle):
    """the reference's two examples (examples/fit_sine.rs, examples/example.rs), ported to the front end, run to completion"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "autograd.rs_b200", "examples", script)] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and needle in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    if script == "fit_sine.py":
        losses = [float(l.split("loss [")[1].split("]")[0]) for l in r.stdout.splitlines() if "loss [" in l]
        assert len(losses) == 3 and losses[-1] < losses[0]


def test_example_rs_hand_written_step_tracks_oracle(dev):
    """examples/example.rs:43-74 with injected weights: forward, MSE, backward, `w -= grad * lr` by tensor arithmetic on the
    parameter handles (in place on the shared storage), zero_grad -- two epochs over two batches of 8, against the oracle."""
    ws, bs = W.mlp_init(2, 128, seed=91, in_features=784, out_features=10)
    rng = np.random.default_rng(92)
    xt, yt = rng.uniform(0, 1, (16, 784)).astype(np.float32), rng.uniform(0, 1, (16, 10)).astype(np.float32)
    model = W.build_mlp(dev, ws, bs, "sigmoid")
    ol = [O.OLinearLayer(w.copy(), b.copy()) for w, b in zip(ws, bs)]
    oparams = [q for l in ol for q in l.parameters()]
    lr = 0.01   # the example's 0.1 saturates the sigmoids on non-zero data (exp overflows): a parity check needs a conditioned problem
    for epoch in range(2):
        for i in (0, 8):
            x, y = T(dev, xt[i:i + 8]), T(dev, yt[i:i + 8])
            y.requires_grad = False
            loss = F.nn.MeanSquaredError().forward([model.forward([x]), y.clone()])
            loss.backward()
            for p in model.parameters():
                g = p.grad() * lr
                p -= g
            model.zero_grad()
            ox, oy = O.OTensor(xt[i:i + 8]), O.OTensor(yt[i:i + 8], requires_grad=False)
            oloss = O.OpMSE().forward([ol[1].forward(O.OpSigmoid().forward([ol[0].forward(ox)])), oy])
            oloss.backward()
            O.sgd(oparams, lr)
            O.model_zero_grad(oparams)
            assert abs(loss.item() - float(oloss.data[0, 0])) / abs(float(oloss.data[0, 0])) < TOL
    for p, q in zip(model.parameters(), oparams):
        assert O.rel_err(p.data(), q.data) < TOL


def test_save_and_load_parameters(dev, tmp_path):
    """AGRSWTS1 round trip: bit-exact values, written into the existing storages (clones follow), shape mismatches refused"""
    ws, bs = W.mlp_init(2, 96, seed=93)
    model = W.build_mlp(dev, ws, bs, "relu")
    path = str(tmp_path / "m.agrs")
    model.save(path)
    raw = open(path, "rb").read()
    assert raw[:8] == b"AGRSWTS1" and int.from_bytes(raw[8:12], "little") == 4
    assert len(raw) == 12 + sum(4 + 8 * a.ndim + a.nbytes for pair in zip(ws, bs) for a in pair)
    other = W.build_mlp(dev, [np.zeros_like(w) for w in ws], [np.zeros_like(b) for b in bs], "relu")
    clones = [p.clone() for p in other.parameters()]
    before = [p.data_ptr() for p in other.parameters()]
    other.load(path)
    for p, c, a in zip(other.parameters(), clones, [a for pair in zip(ws, bs) for a in pair]):
        assert np.array_equal(p.data(), a) and np.array_equal(c.data(), a)
    assert [p.data_ptr() for p in other.parameters()] == before
    wrong = W.build_mlp(dev, [np.zeros((96, 64), np.float32)], [np.zeros((1, 64), np.float32)], None)
    with pytest.raises(F.Panic):
        wrong.load(path)


@pytest.mark.parametrize("B,H,Wd,co,k,s,p", [(64, 28, 28, 6, 5, 1, 0), (3, 9, 11, 8, 3, 1, 1), (5, 12, 10, 4, 3, 2, 1), (1, 7, 7, 5, 2, 2, 0), (300, 28, 28, 6, 5, 1, 0),
                                                 (7, 16, 20, 8, 3, 1, 0), (5, 32, 32, 3, 5, 1, 0), (9, 13, 12, 1, 3, 1, 0), (2, 5, 8, 2, 5, 1, 0)])
def test_conv2d_implicit_gemm_matches_oracle_and_literal_path(dev, B, H, Wd, co, k, s, p):
    """f3: Conv2D with one input channel through the implicit-GEMM kernels (no col matrix): forward and all three gradients against
    the oracle's literal restatement, and against the device's own im2col + GEMM path; the node keeps three variables."""
    rng = np.random.default_rng(B + H + k + p + co)
    im = rng.standard_normal((B, 1, H, Wd)).astype(np.float32)
    filt = rng.standard_normal((k, k, co)).astype(np.float32)
    bias = rng.standard_normal((co,)).astype(np.float32)
    want_out, col2d = O.conv2d_fwd(im, filt, bias, 1, k, s, p)
    g = rng.standard_normal(want_out.shape).astype(np.float32)
    wdx, wdw, wdb = O.conv2d_bwd(im, filt, col2d, g, 1, k, s, p, literal=False)
    res = {}
    for implicit in (True, False):
        dev.set_conv_implicit(implicit)
        try:
            conv = F.Conv2D(1, co, k, s, p)
            X, Fl, Bi = T(dev, im), T(dev, filt), T(dev, bias)
            out = conv.forward([X.clone(), Fl.clone(), Bi.clone()])
            assert out.graph_info() == ("Conv2D", 3 if implicit else 4, 0)
            xs = [X, Fl, Bi] + ([] if implicit else [T(dev, col2d)])
            dx, dw, db = conv.backward(xs, T(dev, g))
            res[implicit] = (out.data(), dx.data(), dw.data(), db.data())
        finally:
            dev.set_conv_implicit(True)
    o, dx, dw, db = res[True]
    assert o.shape == want_out.shape and dx.shape == wdx.shape and dw.shape == (k, k, co) and db.shape == (co,)
    assert O.rel_err(o, want_out) < 1e-5 and O.rel_err(dx, wdx) < 1e-5 and O.rel_err(dw, wdw) < 2e-5 and O.rel_err(db, wdb) < 2e-5
    for a, b in zip(res[True], res[False]):
        assert O.rel_err(a, b) < 2e-5


def test_lenet_implicit_and_literal_conv_train_alike(dev):
    """C3 in miniature through both Conv2D paths: losses of 3 steps agree to 1e-5 and both track the oracle"""
    rng = np.random.default_rng(33)
    B = 48
    im = rng.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)
    t = rng.uniform(0, 1, (B, 10)).astype(np.float32)
    filt = rng.uniform(-0.3, 0.3, (5, 5, 6)).astype(np.float32)
    cb = rng.uniform(-0.1, 0.1, (6,)).astype(np.float32)
    w = rng.uniform(-0.05, 0.05, (3456, 10)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (1, 10)).astype(np.float32)
    runs = []
    for implicit in (True, False):
        dev.set_conv_implicit(implicit)
        try:
            tr = W.Trainer(W.build_lenet(dev, filt, cb, w, b), 1e-3)
            X, Tt = T(dev, im), T(dev, t)
            Tt.requires_grad = False
            n0 = dev.launches()
            runs.append(([tr.step(X, Tt).item() for _ in range(3)], dev.launches() - n0))
        finally:
            dev.set_conv_implicit(True)
    assert np.allclose(runs[0][0], runs[1][0], rtol=1e-5)
    assert runs[0][1] < runs[1][1]  # fewer kernels per step


# ====================================================================== f2: the forward activation out of the GEMM's epilogue
@pytest.mark.parametrize("act,width,rows,layers,path", [("relu", 256, 512, 3, "tc"), ("sigmoid", 384, 640, 3, "tc"),
                                                        ("relu", 1536, 384, 2, "tc"), ("sigmoid", 24, 9, 3, "auto")])
def test_forward_fusion_is_bit_identical_and_keeps_the_graph(dev, act, width, rows, layers, path):
    """Linear feeding ReLU / Sigmoid as ONE GEMM call (Operator.forward_then / agrs_eng_op_forward_pair; layers.rs:24-43): the same
    losses and the same weights bit for bit as the two separate forwards over three training iterations, the same two graph nodes
    (the activation's node keeps the pre-activation, operators.rs:120-129, :214-224), and one launch fewer per activation.
    Width 1536 runs K in two chunks: Z only exists once the second has been reduce-added, the activation warps wait for that."""
    ws, bs = W.mlp_init(layers, width, seed=41)
    x, t = W.mlp_batch(rows, width, width, seed=42)
    if path == "tc":
        dev.set_gemm_path(K.GEMM_TCGEN05)
        dev.set_tuning(K.TUNE_TC_SPLITK, 1)  # few tiles: the planner would split K over CTA pairs, which takes the second-pass form
        dev.set_tuning(K.TUNE_TC_ACT_WARPS, 1)  # the pair kernel's activation warps are opt-in
    try:
        runs = {}
        for fused in (True, False):
            dev.set_forward_fusion(fused)
            model = W.build_mlp(dev, ws, bs, act)
            X = T(dev, x)
            y = model.forward([X.clone()])
            assert y.graph_info() == ("Linear", 3, 1)  # the last layer has no activation
            h = F.nn.Linear(dev, width, width, w=T(dev, ws[0]), bias=T(dev, bs[0]))
            a = F.Linear().forward_then({"relu": F.ReLU, "sigmoid": F.Sigmoid}[act](), [X.clone()] + h.parameters())
            assert a.graph_info() == ({"relu": "ReLU", "sigmoid": "Sigmoid"}[act], 1, 1)
            n0 = dev.launches()
            model.forward([X.clone()])
            runs[fused] = (run_device_mlp(dev, x, t, ws, bs, act, 1e-4, 3), dev.launches() - n0, a.data())
            dev.set_forward_fusion(True)
        (lf, pf, _), nf, af = runs[True]
        (lu, pu, _), nu, au = runs[False]
        assert lf == lu
        assert all(np.array_equal(p, q) for p, q in zip(pf, pu))
        assert np.array_equal(af, au)
        assert nf < nu  # the activation launches are gone (run_device_mlp's launches are counted on both sides)
    finally:
        dev.set_forward_fusion(True)
        dev.set_tuning(K.TUNE_TC_SPLITK, 0)
        dev.set_tuning(K.TUNE_TC_ACT_WARPS, 0)
        dev.set_gemm_path(K.GEMM_AUTO)
    want, _ = O.mlp_train(x, t, ws, bs, act, 1e-4, 3, np.float32)
    assert max(abs(a - b) / abs(b) for a, b in zip(lf, want)) < TOL


def test_forward_pair_of_other_operators_is_two_forwards(dev):
    """agrs_eng_op_forward_pair on a pair it does not fuse: exactly second.forward([first.forward(xs)])"""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6, 5)).astype(np.float32)
    X = T(dev, x)
    y = F.ReLU().forward_then(F.Sigmoid(), [X.clone()])
    assert y.graph_info() == ("Sigmoid", 1, 1)
    want = O.sigmoid_fwd(O.relu_fwd(x))
    assert np.allclose(y.data(), want, rtol=4e-7)


@pytest.mark.parametrize("act", ["relu", "sigmoid"])
@pytest.mark.parametrize("B,H,Wd,co,k,s,p", [(64, 28, 28, 6, 5, 1, 0), (1026, 28, 28, 6, 5, 1, 0), (3, 9, 11, 8, 3, 1, 1), (5, 12, 10, 4, 3, 2, 1),
                                                 (7, 16, 20, 8, 3, 1, 0), (9, 13, 12, 1, 3, 1, 0)])
def test_conv2d_forward_fusion_is_bit_identical_and_keeps_the_graph(dev, act, B, H, Wd, co, k, s, p):
    """Conv2D feeding ReLU / Sigmoid as one launch (agrs_conv2d_fwd_act_f32 through Operator.forward_then): the pre-activation, the
    activation and the gradients that flow back through both nodes carry the bits of the two separate forwards -- register-tiled and
    generic kernels, ragged last group of samples, stride / padding -- and both track the oracle (operators.rs:350-390 into :112-116)."""
    rng = np.random.default_rng(B + H + k + p + co)
    im = rng.standard_normal((B, 1, H, Wd)).astype(np.float32)
    filt = rng.standard_normal((k, k, co)).astype(np.float32)
    bias = rng.standard_normal((co,)).astype(np.float32)
    act_op = {"relu": F.ReLU, "sigmoid": F.Sigmoid}[act]
    res = {}
    for fused in (True, False):
        dev.set_forward_fusion(fused)
        try:
            X, Fl, Bi = T(dev, im), T(dev, filt), T(dev, bias)
            n0 = dev.launches()
            y = F.Conv2D(1, co, k, s, p).forward_then(act_op(), [X.clone(), Fl.clone(), Bi.clone()])
            launches = dev.launches() - n0
            assert y.graph_info() == ({"relu": "ReLU", "sigmoid": "Sigmoid"}[act], 1, 1)
            F.Mean().forward([y]).backward()
            res[fused] = (y.data(), X.grad().data(), Fl.grad().data(), Bi.grad().data(), launches)
        finally:
            dev.set_forward_fusion(True)
    for a, b in zip(res[True][:4], res[False][:4]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert res[True][4] < res[False][4]
    z, _ = O.conv2d_fwd(im, filt, bias, 1, k, s, p)
    want = O.relu_fwd(z) if act == "relu" else O.sigmoid_fwd(z)
    assert O.rel_err(res[True][0], want) < 1e-5


def test_lenet_forward_fusion_trains_alike(dev):
    """C3 in miniature: Conv2D + ReLU as one launch against two, three steps, identical losses and parameters"""
    rng = np.random.default_rng(34)
    B = 48
    im = rng.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)
    t = rng.uniform(0, 1, (B, 10)).astype(np.float32)
    filt = rng.uniform(-0.3, 0.3, (5, 5, 6)).astype(np.float32)
    cb = rng.uniform(-0.1, 0.1, (6,)).astype(np.float32)
    w = rng.uniform(-0.05, 0.05, (3456, 10)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, (1, 10)).astype(np.float32)
    runs = []
    for fused in (True, False):
        dev.set_forward_fusion(fused)
        try:
            net = W.build_lenet(dev, filt, cb, w, b)
            tr = W.Trainer(net, 1e-3)
            X, Tt = T(dev, im), T(dev, t)
            Tt.requires_grad = False
            n0 = dev.launches()
            losses = [tr.step(X, Tt).item() for _ in range(3)]
            runs.append((losses, [q.data() for q in net.parameters()], dev.launches() - n0))
        finally:
            dev.set_forward_fusion(True)
    assert runs[0][0] == runs[1][0]
    assert all(np.array_equal(a, b) for a, b in zip(runs[0][1], runs[1][1]))
    assert runs[0][2] < runs[1][2]
