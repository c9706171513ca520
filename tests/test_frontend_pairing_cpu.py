"""CPU test of the front end's host logic for `Layer::forward_default` (src/nn/layers.rs:24-43) and `Sequential`: which consecutive
operators go to the engine as ONE call (agrs_eng_op_forward_pair: Linear / Conv2D followed by ReLU / Sigmoid) and which as two.
The engine library is replaced by a recorder -- no device, no arithmetic: what is checked is the sequence of C-ABI calls and their
operator kinds / operand counts.  (What the pair call computes is the GPU suite's business: tests/test_gemm_act_gpu.py,
tests/test_engine_gpu.py.)"""
import ctypes as C
import gc

import pytest

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import frontend as F


class Recorder:
    """stands in for libagrs_engine.so: every function succeeds, out-parameters receive a fresh fake handle"""

    def __init__(self):
        self.calls = []
        self._next = 1000

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            for a in args:
                obj = getattr(a, "_obj", None)  # ctypes.byref(c_void_p): an out handle
                if isinstance(obj, C.c_void_p):
                    self._next += 1
                    obj.value = self._next
            return 0
        return fn

    def forwards(self):
        out = []
        for name, args in self.calls:
            if name == "agrs_eng_op_forward":
                out.append(("forward", args[0], args[3]))            # kind, number of inputs
            elif name == "agrs_eng_op_forward_pair":
                out.append(("pair", args[0], args[2], args[5]))      # first kind, second kind, number of inputs
        return out


@pytest.fixture
def rec():
    before, r = F._lib, Recorder()
    F._lib = r
    try:
        yield r
    finally:
        gc.collect()      # the fake handles die while the recorder is still the library their finaliser calls
        F._lib = before


def fake_tensor(rec):
    rec._next += 1
    return F.Tensor(C.c_void_p(rec._next), dev=None)


def linear(rec):
    return F.nn.Linear(None, 4, 4, w=fake_tensor(rec), bias=fake_tensor(rec))


def test_sequential_pairs_linear_with_the_activation_behind_it(rec):
    model = F.nn.Sequential([linear(rec), F.nn.ReLU(), linear(rec), F.nn.Sigmoid(), linear(rec)])
    rec.calls.clear()
    model.forward([fake_tensor(rec)])
    assert rec.forwards() == [("pair", F.OP_LINEAR, F.OP_RELU, 3), ("pair", F.OP_LINEAR, F.OP_SIGMOID, 3), ("forward", F.OP_LINEAR, 3)]


def test_sequential_pairs_conv2d_and_leaves_other_neighbours_alone(rec):
    conv = F.nn.Conv2D(None, 1, 6, 5, kernels=fake_tensor(rec), bias=fake_tensor(rec))
    model = F.nn.Sequential([conv, F.nn.ReLU(), F.nn.Flatten(), linear(rec), F.nn.Identity(), F.nn.ReLU(), F.nn.Sigmoid()])
    rec.calls.clear()
    model.forward([fake_tensor(rec)])
    assert rec.forwards() == [("pair", F.OP_CONV2D, F.OP_RELU, 3), ("forward", F.OP_FLATTEN, 1), ("forward", F.OP_LINEAR, 3),
                              ("forward", F.OP_IDENTITY, 1), ("forward", F.OP_RELU, 1), ("forward", F.OP_SIGMOID, 1)]
    # the Conv2D operator's parameters travel with the pair call (in_channels, out_channels, k, stride, pad)
    name, args = next(c for c in rec.calls if c[0] == "agrs_eng_op_forward_pair")
    assert list(args[1]) == [1, 6, 5, 1, 0] and args[3] is None


def test_a_layer_subclass_with_its_own_forward_is_not_paired(rec):
    class MyLinear(F.nn.Linear):
        def forward(self, xs):
            return super().forward(xs)

    model = F.nn.Sequential([MyLinear(None, 4, 4, w=fake_tensor(rec), bias=fake_tensor(rec)), F.nn.ReLU()])
    rec.calls.clear()
    model.forward([fake_tensor(rec)])
    assert rec.forwards() == [("forward", F.OP_LINEAR, 3), ("forward", F.OP_RELU, 1)]


def test_forward_default_pairs_adjacent_operators_of_one_subgraph(rec):
    """a layer whose subgraph holds several operators (layers.rs:1-6: "layers can have multiple operators"): the loop of
    forward_default feeds each output to the next operator and pairs Linear / Conv2D with an activation that follows directly"""
    class Block(F.nn.Layer):
        def __init__(self, w, b):
            self.w, self.b = w, b

        def get_op_subgraph(self):
            return [F.Linear(), F.ReLU(), F.Sigmoid()]

        def parameters(self):
            return [self.w, self.b]

        def forward(self, xs):
            return self.forward_default(list(xs) + self.parameters())

    blk = Block(fake_tensor(rec), fake_tensor(rec))
    rec.calls.clear()
    blk.forward([fake_tensor(rec)])
    assert rec.forwards() == [("pair", F.OP_LINEAR, F.OP_RELU, 3), ("forward", F.OP_SIGMOID, 1)]


def test_layers_called_one_by_one_stay_two_calls(rec):
    """the reference's own examples call layer after layer (model.rs:47-51): nothing is deferred, so nothing is paired"""
    l1, act = linear(rec), F.nn.ReLU()
    rec.calls.clear()
    x = l1.forward([fake_tensor(rec)])
    act.forward([x])
    assert rec.forwards() == [("forward", F.OP_LINEAR, 3), ("forward", F.OP_RELU, 1)]
