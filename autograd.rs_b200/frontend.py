"""Python handles over the C++ host engine (include/agrs_engine.h, libagrs_engine.so).

The names and semantics are the reference's (NickLucche/autograd.rs) -- this is how its unit tests and
examples read when the device arm exists:

  Tensor                         src/tensor/tensor.rs:27-34 and src/tensor/tensor_arithmetic.rs
  ReLU, Sigmoid, Linear, MatMul, MeanSquaredError, Mean, Identity, Conv2D
        .forward(xs) / .backward(xs, grad)        trait Operator, src/operators/operators.rs:18-50
  nn.Linear / nn.Conv2D / nn.ReLU / nn.Sigmoid / nn.MeanSquaredError / nn.Identity, nn.NN, nn.SGD
                                 src/nn/layers.rs, src/nn/model.rs, src/nn/optim.rs

Every method is one call into the engine, which turns it into asynchronous CUDA launches through the
kernel-level C ABI (include/agrs.h).  There is no CPU implementation behind any of it: without the
built libraries or without an sm_100 device, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libagrs_engine.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "agrs_engine.h")

OP_RELU, OP_SIGMOID, OP_LINEAR, OP_MATMUL, OP_MSE, OP_MEAN, OP_IDENTITY, OP_CONV2D, OP_FLATTEN = range(9)

_lib = None


class Panic(RuntimeError):
    """The reference panics; the engine reports the same message with a status code."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def declared_symbols():
    return _capi.declared_symbols(HEADER)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    _capi.lib()  # libagrs_b200.so first (the engine links against it)
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, i64, f32, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int
    pvp, pi64, pi32 = C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)
    sig = {
        "agrs_eng_device_create": [i32, pvp],
        "agrs_eng_device_create_on_stream": [i32, vp, pvp],
        "agrs_eng_device_destroy": [vp],
        "agrs_eng_device_sync": [vp],
        "agrs_eng_device_seed": [vp, C.c_uint64],
        "agrs_eng_device_set_matmul_backward_literal": [vp, i32],
        "agrs_eng_device_set_fusions": [vp, i32],
        "agrs_eng_device_set_conv_implicit": [vp, i32],
        "agrs_eng_device_set_forward_fusion": [vp, i32],
        "agrs_eng_save_parameters": [pvp, i32, C.c_char_p],
        "agrs_eng_load_parameters": [pvp, i32, C.c_char_p],
        "agrs_eng_tensor_from_host": [vp, vp, pi64, i32, pvp],
        "agrs_eng_tensor_full": [vp, pi64, i32, f32, pvp],
        "agrs_eng_tensor_uniform": [vp, pi64, i32, f32, f32, pvp],
        "agrs_eng_tensor_kaiming_uniform": [vp, pi64, i32, pvp],
        "agrs_eng_tensor_copy_from_host": [vp, vp, i64, i32],
        "agrs_eng_tensor_prefetch_from_host": [vp, vp, i64],
        "agrs_eng_device_prefetch_join": [vp],
        "agrs_eng_event_create": [vp, C.POINTER(vp)],
        "agrs_eng_event_destroy": [vp, vp],
        "agrs_eng_event_wait": [vp, vp],
        "agrs_eng_tensor_to_host_async": [vp, vp, i64, vp],
        "agrs_eng_tensor_clone": [vp, pvp],
        "agrs_eng_tensor_to_owned": [vp, pvp],
        "agrs_eng_tensor_free": [vp],
        "agrs_eng_tensor_ndim": [vp, pi32],
        "agrs_eng_tensor_shape": [vp, pi64, i32],
        "agrs_eng_tensor_len": [vp, pi64],
        "agrs_eng_tensor_is_contiguous": [vp, pi32],
        "agrs_eng_tensor_to_host": [vp, vp, i64],
        "agrs_eng_tensor_data_ptr": [vp, pvp],
        "agrs_eng_tensor_requires_grad": [vp, pi32],
        "agrs_eng_tensor_set_requires_grad": [vp, i32],
        "agrs_eng_tensor_has_grad": [vp, pi32],
        "agrs_eng_tensor_grad": [vp, pvp],
        "agrs_eng_tensor_graph_info": [vp, C.c_char_p, i32, pi32, pi32],
        "agrs_eng_tensor_reshape": [vp, pi64, i32],
        "agrs_eng_tensor_t": [vp],
        "agrs_eng_tensor_t_clone": [vp, pvp],
        "agrs_eng_tensor_unsqueeze": [vp, i32, pvp],
        "agrs_eng_tensor_squeeze": [vp, pvp],
        "agrs_eng_tensor_binary": [i32, i32, vp, vp, pvp],
        "agrs_eng_tensor_scalar": [i32, i32, vp, f32, pvp],
        "agrs_eng_tensor_neg": [vp, pvp],
        "agrs_eng_tensor_equal": [vp, vp, pi32],
        "agrs_eng_tensor_fill": [vp, f32],
        "agrs_eng_tensor_powi": [vp, i32, i32, pvp],
        "agrs_eng_tensor_sum": [vp, C.POINTER(f32)],
        "agrs_eng_tensor_sum_axis": [vp, i32, pvp],
        "agrs_eng_tensor_mean": [vp, pvp],
        "agrs_eng_tensor_dot": [vp, vp, pvp],
        "agrs_eng_tensor_zero_grad": [vp],
        "agrs_eng_tensor_accumulate_grad": [vp, vp],
        "agrs_eng_tensor_backward": [vp],
        "agrs_eng_op_forward": [i32, pi64, pvp, i32, pvp],
        "agrs_eng_op_forward_pair": [i32, pi64, i32, pi64, pvp, i32, pvp],
        "agrs_eng_op_backward": [i32, pi64, pvp, i32, vp, pvp, i32, pi32],
        "agrs_eng_im2col": [vp, i64, i64, i64, pvp],
        "agrs_eng_im2col_backward": [vp, i64, i64, i64, i64, pvp],
        "agrs_eng_sgd_step": [pvp, i32, f32],
        "agrs_eng_zero_grad": [pvp, i32],
        "agrs_eng_dp_create": [vp, i32, i32, vp, pvp],
        "agrs_eng_dp_destroy": [vp],
        "agrs_eng_dp_register": [vp, pvp, i32, i32],
        "agrs_eng_dp_sync_grads": [vp],
        "agrs_eng_dp_mean_scalar": [vp, vp],
        "agrs_eng_dp_fused_prepare": [vp, vp],
        "agrs_eng_dp_fused_connect": [vp, vp],
        "agrs_eng_dp_fused_sgd_step": [vp, f32],
        "agrs_eng_dp_fused_begin": [vp, f32],
        "agrs_eng_dp_fused_exposed_ms": [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)],
        "agrs_eng_dp_fused_finish": [vp],
        "agrs_eng_dp_fused_disable": [vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.agrs_eng_last_error.argtypes = []
    L.agrs_eng_last_error.restype = C.c_char_p
    L.agrs_eng_device_ctx.argtypes = [vp]
    L.agrs_eng_device_ctx.restype = vp
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise Panic(rc, lib().agrs_eng_last_error().decode(errors="replace"))


def _shape_arg(shape):
    shape = tuple(int(s) for s in shape)
    return (C.c_int64 * max(len(shape), 1))(*shape), len(shape)


def _handles(tensors):
    arr = (C.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t._h
    return arr


class Device:
    """One GPU: an agrs_ctx (stream + memory pool) and the host RNG of the initialisers."""

    def __init__(self, ordinal=0, stream=None):
        h = C.c_void_p()
        L = lib()
        rc = L.agrs_eng_device_create(ordinal, C.byref(h)) if stream is None else L.agrs_eng_device_create_on_stream(ordinal, stream, C.byref(h))
        _check(rc)
        self._h = h
        self.ordinal = ordinal
        self.ctx_handle = C.c_void_p(L.agrs_eng_device_ctx(h))

    def close(self):
        if self._h:
            lib().agrs_eng_device_destroy(self._h)
            self._h = None

    def sync(self):
        _check(lib().agrs_eng_device_sync(self._h))

    def seed(self, s):
        _check(lib().agrs_eng_device_seed(self._h, int(s)))

    def prefetch_join(self):
        """the compute stream waits for every Tensor.prefetch_from_numpy issued so far"""
        _check(lib().agrs_eng_device_prefetch_join(self._h))

    def set_fusions(self, backward_colsum=True):
        """operator-boundary fusions (on by default; bit-identical either way): db left by the backward kernel that produced the gradient"""
        _check(lib().agrs_eng_device_set_fusions(self._h, int(bool(backward_colsum))))

    def set_forward_fusion(self, on=True):
        """Linear feeding ReLU / Sigmoid as one GEMM call whose epilogue writes the activation next to the pre-activation
        (default; bit-identical to the two separate forwards)"""
        _check(lib().agrs_eng_device_set_forward_fusion(self._h, int(bool(on))))

    def set_conv_implicit(self, on=True):
        """Conv2D with one input channel: implicit-GEMM kernels without a col matrix (default) or the literal im2col lowering"""
        _check(lib().agrs_eng_device_set_conv_implicit(self._h, int(bool(on))))

    def set_matmul_backward_literal(self, literal):
        _check(lib().agrs_eng_device_set_matmul_backward_literal(self._h, int(bool(literal))))

    def launches(self):
        """kernels launched so far on this device's context (bench.py's gpu_launches)"""
        return int(_capi.lib().agrs_launch_count(self.ctx_handle))

    def stream(self):
        return _capi.lib().agrs_stream(self.ctx_handle)

    def set_tuning(self, key, value):
        rc = _capi.lib().agrs_set_tuning(self.ctx_handle, int(key), int(value))
        if rc:
            raise Panic(rc, _capi.lib().agrs_last_error(self.ctx_handle).decode())

    def get_tuning(self, key):
        v = C.c_int64()
        rc = _capi.lib().agrs_get_tuning(self.ctx_handle, int(key), C.byref(v))
        if rc:
            raise Panic(rc, _capi.lib().agrs_last_error(self.ctx_handle).decode())
        return int(v.value)

    def set_gemm_path(self, path):
        rc = _capi.lib().agrs_set_gemm_path(self.ctx_handle, int(path))
        if rc:
            raise Panic(rc, _capi.lib().agrs_last_error(self.ctx_handle).decode())

    def _k(self, name, *args):
        rc = getattr(_capi.lib(), name)(self.ctx_handle, *args)
        if rc:
            raise Panic(rc, _capi.lib().agrs_last_error(self.ctx_handle).decode())

    def timer_begin(self):
        self._k("agrs_timer_begin")

    def timer_end(self):
        """milliseconds of device time since timer_begin (CUDA events on the launching stream); synchronises"""
        ms = C.c_float()
        self._k("agrs_timer_end", C.byref(ms))
        return float(ms.value)

    def profile_enable(self, on):
        self._k("agrs_profile_enable", int(bool(on)))

    def profile_gemm(self, path=_capi.GEMM_TCGEN05):
        """(launches, total_ms, total_algorithmic_flops) of the GEMM launches on `path` since the last call"""
        n, ms, fl = C.c_uint64(), C.c_double(), C.c_double()
        self._k("agrs_profile_gemm", int(path), C.byref(n), C.byref(ms), C.byref(fl))
        return int(n.value), float(ms.value), float(fl.value)

    # ---- memory pool: back it with device memory up front so that no training step pays for pool growth (a driver call)
    def pool_reserve(self, nbytes):
        self._k("agrs_pool_reserve", int(nbytes))

    def pool_trim(self):
        """give the pool's unused blocks back to the driver and reset its high-water marks (blocks)"""
        self._k("agrs_pool_trim")

    def pool_stats(self):
        """(reserved now, used now, reserved high-water mark, used high-water mark) in bytes"""
        a = [C.c_size_t() for _ in range(4)]
        self._k("agrs_pool_stats", *[C.byref(x) for x in a])
        return tuple(int(x.value) for x in a)

    # ---- whole-step CUDA graphs (agrs_capture_*): record one or more training steps, replay them with one driver call each
    def capture_begin(self):
        self._k("agrs_capture_begin")

    def capture_end(self):
        g = C.c_void_p()
        self._k("agrs_capture_end", C.byref(g))
        return Graph(self, g)

    # pinned host staging for end-to-end runs (agrs_host_alloc): returns a float32 numpy view
    def pinned_empty(self, shape):
        n = int(np.prod(shape))
        p = C.c_void_p()
        rc = _capi.lib().agrs_host_alloc(self.ctx_handle, max(n, 1) * 4, C.byref(p))
        if rc:
            raise Panic(rc, _capi.lib().agrs_last_error(self.ctx_handle).decode())
        buf = (C.c_float * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float32, count=n).reshape(shape)
        return arr


class Event:
    """Completion ticket of an asynchronous read-back (agrs_event)."""

    def __init__(self, dev):
        self.dev = dev
        h = C.c_void_p()
        _check(lib().agrs_eng_event_create(dev._h, C.byref(h)))
        self._h = h

    def wait(self):
        _check(lib().agrs_eng_event_wait(self.dev._h, self._h))

    def close(self):
        if self._h:
            lib().agrs_eng_event_destroy(self.dev._h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self.dev._h:
                self.close()
        except Exception:
            pass


class Graph:
    """A captured sequence of training steps (agrs_graph): every kernel, memset, allocation and collective of the step,
    replayed by the driver without the host walking the model again."""

    def __init__(self, dev, handle):
        self.dev, self._h = dev, handle
        k, n = C.c_uint64(), C.c_uint64()
        _capi.lib().agrs_graph_info(handle, C.byref(k), C.byref(n))
        self.kernel_nodes, self.total_nodes = int(k.value), int(n.value)

    def launch(self, times=1):
        self.dev._k("agrs_graph_launch", self._h, int(times))

    def close(self):
        if self._h:
            _capi.lib().agrs_graph_destroy(self.dev.ctx_handle, self._h)
            self._h = None


class Tensor:
    """`Tensor<f32>` handle: clones share data, grad and graph (tensor.rs:24-34)."""

    __slots__ = ("_h", "dev")

    def __init__(self, handle, dev):
        self._h = handle
        self.dev = dev

    def __del__(self):
        try:
            if self._h and _lib is not None:
                _lib.agrs_eng_tensor_free(self._h)
        except Exception:
            pass
        self._h = None

    @staticmethod
    def _new(dev, fn, *args):
        h = C.c_void_p()
        _check(fn(*args, C.byref(h)))
        return Tensor(h, dev) if h.value else None

    # ---- constructors ----
    @staticmethod
    def from_numpy(dev, arr):
        """Tensor::from(ArrayD) (tensor.rs:37-49): uploads; requires_grad = true, graph = None, grad = None."""
        a = np.ascontiguousarray(arr, dtype=np.float32)
        sh, nd = _shape_arg(a.shape)
        return Tensor._new(dev, lib().agrs_eng_tensor_from_host, dev._h, a.ctypes.data, sh, nd)

    @staticmethod
    def full(dev, shape, value):
        sh, nd = _shape_arg(shape)
        return Tensor._new(dev, lib().agrs_eng_tensor_full, dev._h, sh, nd, float(value))

    @staticmethod
    def zeros(dev, shape):
        return Tensor.full(dev, shape, 0.0)

    @staticmethod
    def ones(dev, shape):
        return Tensor.full(dev, shape, 1.0)

    @staticmethod
    def uniform(dev, shape, low, high):
        sh, nd = _shape_arg(shape)
        return Tensor._new(dev, lib().agrs_eng_tensor_uniform, dev._h, sh, nd, float(low), float(high))

    @staticmethod
    def kaiming_uniform(dev, shape):
        sh, nd = _shape_arg(shape)
        return Tensor._new(dev, lib().agrs_eng_tensor_kaiming_uniform, dev._h, sh, nd)

    def copy_from_numpy(self, arr, blocking=True):
        """Overwrite the data in place from a host array (blocking=False: `arr` must be pinned and stay alive until a sync)."""
        assert arr.dtype == np.float32 and arr.flags["C_CONTIGUOUS"]
        _check(lib().agrs_eng_tensor_copy_from_host(self._h, arr.ctypes.data, int(arr.size), int(blocking)))

    def prefetch_from_numpy(self, arr):
        """Same, on the device's copy stream: overlaps the step already enqueued.  `arr` must be pinned (Device.pinned_empty).
        The tensor carries the copy's completion event: the first kernel that reads (or overwrites) it waits, on the device, for
        this copy and no other.  (Device.prefetch_join() still makes the compute stream wait for every prefetch issued so far.)"""
        assert arr.dtype == np.float32 and arr.flags["C_CONTIGUOUS"]
        _check(lib().agrs_eng_tensor_prefetch_from_host(self._h, arr.ctypes.data, int(arr.size)))

    def clone(self):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_clone, self._h)

    def to_owned(self):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_to_owned, self._h)

    # ---- metadata / read back ----
    def shape(self):
        nd = C.c_int()
        _check(lib().agrs_eng_tensor_ndim(self._h, C.byref(nd)))
        buf = (C.c_int64 * max(nd.value, 1))()
        _check(lib().agrs_eng_tensor_shape(self._h, buf, nd.value))
        return [int(buf[i]) for i in range(nd.value)]

    def ndim(self):
        return len(self.shape())

    def len(self):
        n = C.c_int64()
        _check(lib().agrs_eng_tensor_len(self._h, C.byref(n)))
        return int(n.value)

    size = len

    def is_empty(self):
        return self.len() == 0

    def is_contiguous(self):
        v = C.c_int()
        _check(lib().agrs_eng_tensor_is_contiguous(self._h, C.byref(v)))
        return bool(v.value)

    def data(self, out=None):
        """Read back (D2H, blocks): the device analogue of `data()` / `into_raw_vec` (storage.rs:139-177)."""
        sh = self.shape()
        if out is None:
            out = np.empty(sh, dtype=np.float32)
        _check(lib().agrs_eng_tensor_to_host(self._h, out.ctypes.data, int(out.size)))
        return out

    numpy = data

    def read_async(self, out, event):
        """Enqueue the read-back of this tensor into the pinned array `out` (Device.pinned_empty) and return at once;
        `event.wait()` blocks until it has landed.  The loss of step i can be read while step i+1 already runs."""
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        _check(lib().agrs_eng_tensor_to_host_async(self._h, out.ctypes.data, int(out.size), event._h))

    def item(self):
        return float(self.data().reshape(-1)[0])

    def data_ptr(self):
        p = C.c_void_p()
        _check(lib().agrs_eng_tensor_data_ptr(self._h, C.byref(p)))
        return p.value

    @property
    def requires_grad(self):
        v = C.c_int()
        _check(lib().agrs_eng_tensor_requires_grad(self._h, C.byref(v)))
        return bool(v.value)

    @requires_grad.setter
    def requires_grad(self, yes):
        _check(lib().agrs_eng_tensor_set_requires_grad(self._h, int(bool(yes))))

    def has_grad(self):
        v = C.c_int()
        _check(lib().agrs_eng_tensor_has_grad(self._h, C.byref(v)))
        return bool(v.value)

    def grad(self):
        """`grad()`: None, or a handle on the gradient storage."""
        return Tensor._new(self.dev, lib().agrs_eng_tensor_grad, self._h)

    def graph_info(self):
        name = C.create_string_buffer(64)
        nv, npar = C.c_int(), C.c_int()
        _check(lib().agrs_eng_tensor_graph_info(self._h, name, 64, C.byref(nv), C.byref(npar)))
        return name.value.decode(), nv.value, npar.value

    # ---- shape ops ----
    def reshape(self, shape):
        sh, nd = _shape_arg(shape)
        _check(lib().agrs_eng_tensor_reshape(self._h, sh, nd))
        return self

    def t(self):
        _check(lib().agrs_eng_tensor_t(self._h))
        return self

    def t_clone(self):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_t_clone, self._h)

    def unsqueeze(self, axis):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_unsqueeze, self._h, int(axis))

    def squeeze(self):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_squeeze, self._h)

    # ---- arithmetic (tensor_arithmetic.rs): Python's `a + b` is the reference's `&a + &b` (new tensor);
    # the consuming / assigning forms `a + &b`, `a += &b` are add_(), sub_(), mul_() and the augmented operators.
    def _binary(self, op, inplace, other):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_binary, op, int(inplace), self._h, other._h)

    def _scalar(self, op, inplace, s):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_scalar, op, int(inplace), self._h, float(s))

    def __add__(self, o):
        return self._binary(_capi.ADD, False, o)

    def __sub__(self, o):
        return self._binary(_capi.SUB, False, o)

    def __mul__(self, o):
        return self._binary(_capi.MUL, False, o) if isinstance(o, Tensor) else self._scalar(_capi.SMUL, False, o)

    def __truediv__(self, s):
        return self._scalar(_capi.SDIV, False, s)

    def __rsub__(self, s):
        return self._scalar(_capi.SRSUB, False, s)

    def __neg__(self):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_neg, self._h)

    def add_(self, o):
        return self._binary(_capi.ADD, True, o)

    def sub_(self, o):
        return self._binary(_capi.SUB, True, o)

    def mul_(self, o):
        return self._binary(_capi.MUL, True, o) if isinstance(o, Tensor) else self._scalar(_capi.SMUL, True, o)

    def div_(self, s):
        return self._scalar(_capi.SDIV, True, s)

    def __iadd__(self, o):
        self.add_(o)
        return self

    def __isub__(self, o):
        self.sub_(o)
        return self

    def __imul__(self, o):
        self.mul_(o)
        return self

    def __itruediv__(self, s):
        self.div_(s)
        return self

    def equals(self, o):
        v = C.c_int()
        _check(lib().agrs_eng_tensor_equal(self._h, o._h, C.byref(v)))
        return bool(v.value)

    def fill(self, v):
        _check(lib().agrs_eng_tensor_fill(self._h, float(v)))

    def powi(self, e):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_powi, self._h, int(e), 0)

    def powi_inplace(self, e):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_powi, self._h, int(e), 1)

    def sum(self):
        v = C.c_float()
        _check(lib().agrs_eng_tensor_sum(self._h, C.byref(v)))
        return float(v.value)

    def sum_axis(self, axis):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_sum_axis, self._h, int(axis))

    def mean(self):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_mean, self._h)

    def dot(self, o):
        return Tensor._new(self.dev, lib().agrs_eng_tensor_dot, self._h, o._h)

    # ---- autograd ----
    def zero_grad(self):
        _check(lib().agrs_eng_tensor_zero_grad(self._h))

    def accumulate_grad_from_grad_tensor(self, g):
        _check(lib().agrs_eng_tensor_accumulate_grad(self._h, g._h))

    def backward(self):
        _check(lib().agrs_eng_tensor_backward(self._h))


# ---------------------------------------------------------------------------------------------- operators
class Operator:
    """trait Operator (operators.rs:18-50)."""

    kind = None

    def _params(self):
        return None

    def forward(self, xs):
        xs = list(xs)
        return Tensor._new(xs[0].dev, lib().agrs_eng_op_forward, self.kind, self._params(), _handles(xs), len(xs))

    def forward_then(self, second, xs):
        """second.forward([self.forward(xs)]) -- two consecutive iterations of Layer::forward_default's loop (layers.rs:24-43):
        the same two graph nodes and the same bits, as ONE engine call.  Linear followed by ReLU / Sigmoid goes out as a single
        GEMM whose epilogue writes the activation next to the pre-activation."""
        xs = list(xs)
        return Tensor._new(xs[0].dev, lib().agrs_eng_op_forward_pair, self.kind, self._params(), second.kind, second._params(),
                           _handles(xs), len(xs))

    def backward(self, xs, grad):
        xs = list(xs)
        outs = (C.c_void_p * 8)()
        n = C.c_int()
        _check(lib().agrs_eng_op_backward(self.kind, self._params(), _handles(xs), len(xs), grad._h, outs, 8, C.byref(n)))
        return [Tensor(C.c_void_p(outs[i]), grad.dev) for i in range(n.value)]


class ReLU(Operator):
    kind = OP_RELU


class Sigmoid(Operator):
    kind = OP_SIGMOID


class Linear(Operator):
    kind = OP_LINEAR


class MatMul(Operator):
    kind = OP_MATMUL


class MeanSquaredError(Operator):
    kind = OP_MSE


class Mean(Operator):
    kind = OP_MEAN


class Identity(Operator):
    kind = OP_IDENTITY


class Flatten(Operator):
    """Not in the reference (SURVEY gap G1): a [B, ...] -> [B, prod(...)] view, so Conv2D can feed Linear."""
    kind = OP_FLATTEN


class Conv2D(Operator):
    """operators.rs:66-73, :288-450."""
    kind = OP_CONV2D

    def __init__(self, in_channels, out_channels, k_size, stride=1, pad=0):
        self.in_channels, self.out_channels, self.k_size, self.stride, self.pad = in_channels, out_channels, k_size, stride, pad

    def _params(self):
        return (C.c_int64 * 5)(self.in_channels, self.out_channels, self.k_size, self.stride, self.pad)

    def get_output_shape(self, h, w):
        return ((h - self.k_size + 2 * self.pad) // self.stride + 1, (w - self.k_size + 2 * self.pad) // self.stride + 1)

    @staticmethod
    def im2col(im, k_size, stride, pad):
        return Tensor._new(im.dev, lib().agrs_eng_im2col, im._h, k_size, stride, pad)

    @staticmethod
    def im2col_backward(col, h_out, w_out, stride, k_size):
        return Tensor._new(col.dev, lib().agrs_eng_im2col_backward, col._h, h_out, w_out, stride, k_size)


# ---------------------------------------------------------------------------------------------- nn
class nn:
    """src/nn: Layer / NN / SGD."""

    class Layer:  # layers.rs:16-44
        def get_op_subgraph(self):
            raise NotImplementedError

        def parameters(self):
            return []

        def forward(self, xs):
            return self.forward_default(xs)

        def forward_default(self, xs):
            xs = list(xs)
            res = xs[0].clone()
            ops = self.get_op_subgraph()
            i = 0
            while i < len(ops):
                if i + 1 < len(ops) and type(ops[i]) in (Linear, Conv2D) and type(ops[i + 1]) in (ReLU, Sigmoid):
                    res = ops[i].forward_then(ops[i + 1], xs)  # one GEMM call, two graph nodes
                    i += 2
                else:
                    res = ops[i].forward(xs)
                    i += 1
                xs = [res.clone()]
            return res

    class Identity(Layer):
        def get_op_subgraph(self):
            return [Identity()]

    class ReLU(Layer):
        def get_op_subgraph(self):
            return [ReLU()]

    class Sigmoid(Layer):
        def get_op_subgraph(self):
            return [Sigmoid()]

    class MeanSquaredError(Layer):
        def get_op_subgraph(self):
            return [MeanSquaredError()]

    class Flatten(Layer):
        def get_op_subgraph(self):
            return [Flatten()]

    class Linear(Layer):
        """layers.rs:66-105: w [in, out] kaiming_uniform, bias [1, out] ~ U(+-1/sqrt(fan_in(w)))."""

        def __init__(self, dev, in_features, out_features, w=None, bias=None):
            if w is None:
                w = Tensor.kaiming_uniform(dev, (in_features, out_features))
                bound = 1.0 / np.sqrt(float(out_features))  # fan_in = shape[1] (init.rs:30-45)
                bias = Tensor.uniform(dev, (1, out_features), -bound, bound)
            self.w, self.bias = w, bias

        def get_op_subgraph(self):
            return [Linear()]

        def parameters(self):
            return [self.w.clone(), self.bias.clone()]

        def forward(self, xs):
            assert len(xs) == 1
            return self.forward_default(list(xs) + self.parameters())

    class Conv2D(Layer):
        """layers.rs:143-186: kernels [k, k, Co], bias [Co]."""

        def __init__(self, dev, in_channels, out_channels, k_size, stride=1, pad=0, kernels=None, bias=None):
            if kernels is None:
                kernels = Tensor.kaiming_uniform(dev, (k_size, k_size, out_channels))
                bound = 1.0 / np.sqrt(float(k_size * out_channels))
                bias = Tensor.uniform(dev, (out_channels,), -bound, bound)
            self.kernels, self.bias = kernels, bias
            self.op = Conv2D(in_channels, out_channels, k_size, stride, pad)

        def get_op_subgraph(self):
            return [self.op]

        def parameters(self):
            return [self.kernels.clone(), self.bias.clone()]

        def forward(self, xs):
            assert len(xs) == 1
            return self.forward_default(list(xs) + self.parameters())

    class NN:  # model.rs:7-23
        def layers(self):
            raise NotImplementedError

        def parameters(self):
            return [p for l in self.layers() for p in l.parameters()]

        def zero_grad(self):
            ps = self.parameters()
            _check(lib().agrs_eng_zero_grad(_handles(ps), len(ps)))

        def save(self, path):
            """model serialisation (README.md:71 lists it as a TODO of the reference): every parameter, in parameters() order"""
            ps = self.parameters()
            _check(lib().agrs_eng_save_parameters(_handles(ps), len(ps), os.fsencode(path)))

        def load(self, path):
            """write a saved file into this model's parameter storages (shapes must match)"""
            ps = self.parameters()
            _check(lib().agrs_eng_load_parameters(_handles(ps), len(ps), os.fsencode(path)))

    class Sequential(NN):
        """Convenience for the benchmark configurations: layers applied in order."""

        def __init__(self, layers):
            self._layers = list(layers)
            self._params = [p for l in self._layers for p in l.parameters()]

        def layers(self):
            return self._layers

        def parameters(self):
            return self._params

        def forward(self, xs):
            # a Linear / Conv2D layer followed by a ReLU / Sigmoid layer: both operators in one engine call (same graph, same bits; the
            # activation comes out of the GEMM's epilogue -- Operator.forward_then)
            t = xs[0]
            ls = self._layers
            i = 0
            while i < len(ls):
                if i + 1 < len(ls) and type(ls[i]) in (nn.Linear, nn.Conv2D) and type(ls[i + 1]) in (nn.ReLU, nn.Sigmoid):
                    t = ls[i].get_op_subgraph()[0].forward_then(ls[i + 1].get_op_subgraph()[0], [t] + ls[i].parameters())
                    i += 2
                else:
                    t = ls[i].forward([t])
                    i += 1
            return t

    class SGD:  # optim.rs:13-35
        def __init__(self, parameters, lr):
            self.parameters, self.lr = list(parameters), float(lr)
            self._arr = _handles(self.parameters)

        def step(self):
            _check(lib().agrs_eng_sgd_step(self._arr, len(self.parameters), self.lr))

        def zero_grad(self, model):
            model.zero_grad()


# ---------------------------------------------------------------------------------------------- data parallel
def comm_unique_id():
    """128-byte NCCL id made on rank 0; the launcher (torch.distributed / a file / MPI) broadcasts it."""
    buf = C.create_string_buffer(128)
    rc = _capi.lib().agrs_comm_unique_id(buf)
    if rc:
        raise Panic(rc, "agrs_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw


def shard_rows(rows, world, rank):
    """rows [begin, end) of a global batch owned by `rank` (equal shards, remainder to the first ranks) -- DataParallel::shard_rows"""
    base, rem = divmod(rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class DataParallel:
    """Replicated parameters, sharded minibatch rows, parameter gradients averaged with NCCL over NVLink."""

    def __init__(self, dev, world, rank, unique_id, parameters, overlap=True):
        h = C.c_void_p()
        _check(lib().agrs_eng_dp_create(dev._h, world, rank, unique_id, C.byref(h)))
        self._h, self.world, self.rank = h, world, rank
        self._params = list(parameters)
        _check(lib().agrs_eng_dp_register(h, _handles(self._params), len(self._params), int(overlap)))

    def sync_grads(self):
        _check(lib().agrs_eng_dp_sync_grads(self._h))

    def mean_scalar(self, t):
        _check(lib().agrs_eng_dp_mean_scalar(self._h, t._h))

    # ---- fused optimiser step over NVLink peer memory (csrc/fused_dp.cu) ----
    fused = False

    def enable_fused(self, all_gather):
        """Move the parameters into an IPC-exported arena and map every peer's arena.  `all_gather(bytes) -> list[bytes]` is the
        launcher's exchange (rank order), e.g. torch.distributed.all_gather_object; for world == 1 pass `lambda h: [h]`."""
        # Both phases end in a collective vote, so a rank whose set-up fails (no IPC / no peer access on this machine) can never
        # leave the others waiting: either every rank switches to the fused step or none does (the NCCL exchange stays in use;
        # parameters that already moved into the arena simply keep living there).  Returns True when the fused step is active.
        buf = C.create_string_buffer(64)
        err = None
        try:
            _check(lib().agrs_eng_dp_fused_prepare(self._h, buf))
            mine = buf.raw
        except Panic as e:
            err, mine = e, b""
        handles = all_gather(mine)
        assert len(handles) == self.world
        if any(len(h) != 64 for h in handles):
            self.fused_error = err or "a peer could not export its arena"
            lib().agrs_eng_dp_fused_disable(self._h)
            return False
        try:
            _check(lib().agrs_eng_dp_fused_connect(self._h, b"".join(handles)))
            ok = b"\x01"
        except Panic as e:
            err, ok = e, b"\x00"
        votes = all_gather(ok)
        if any(v != b"\x01" for v in votes):
            # this rank may have connected: undo it, or its dW GEMMs would keep pushing tiles to peers that never sum them
            self.fused_error = err or "a peer could not map the arenas"
            lib().agrs_eng_dp_fused_disable(self._h)
            return False
        self.fused = True
        return True

    def fused_begin(self, lr):
        """before backward: every layer's reduce-scatter + SGD + all-gather then starts the moment backward has produced that
        layer's gradients and overlaps the backward of the earlier layers (call fused_finish after backward)"""
        _check(lib().agrs_eng_dp_fused_begin(self._h, float(lr)))

    def fused_finish(self):
        _check(lib().agrs_eng_dp_fused_finish(self._h))

    def fused_exposed_ms(self):
        """(total ms, steps): time the compute stream spent waiting for the exchange while Device.profile_enable was on"""
        ms, n = C.c_double(), C.c_uint64()
        _check(lib().agrs_eng_dp_fused_exposed_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def fused_sgd_step(self, lr):
        """reduce-scatter(grad) + w -= lr * mean(grad) + all-gather(w) in one kernel; replaces sync_grads() and SGD.step()"""
        _check(lib().agrs_eng_dp_fused_sgd_step(self._h, float(lr)))

    def close(self):
        if self._h:
            lib().agrs_eng_dp_destroy(self._h)
            self._h = None
