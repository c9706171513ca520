"""ctypes binding of include/agrs.h (libagrs_b200.so) -- the kernel-level C ABI.

There is no fallback: if the library is missing or the machine has no sm_100 device,
loading / context creation raises.
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libagrs_b200.so")
HEADER = os.path.join(ROOT, "include", "agrs.h")

OK = 0
ADD, SUB, MUL = 0, 1, 2
SMUL, SDIV, SRSUB, SADD = 0, 1, 2, 3
BCAST_TRAILING, BCAST_LEADING = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05, GEMM_SKINNY = 0, 1, 2, 3
TUNE_TC_BN, TUNE_TC_EXPLICIT_HI, TUNE_TC_MIN_MACS, TUNE_TC_KCHUNK, TUNE_TC_2CTA, TUNE_TC_SPLITK, TUNE_TC_CROSS, TUNE_TC_EPI_TMA, TUNE_TC_REUSE_SCALES, TUNE_TC_WAVE_SYNC, TUNE_TC_ACT_WARPS = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10

_lib = None


class AgrsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"agrs error {code}: {msg}")
        self.code = code


def declared_symbols(header=HEADER):
    """Every function include/agrs.h declares (used by the CPU-side export test)."""
    with open(header) as f:
        src = f.read()
    return sorted(set(re.findall(r"^AGRS_API [\w\s\*]+?\b(agrs_\w+)\(", src, flags=re.M)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, sz, i64, f32, i32 = C.c_void_p, C.c_size_t, C.c_int64, C.c_float, C.c_int
        sig = {
            "agrs_ctx_create": [i32, C.POINTER(vp)],
            "agrs_ctx_create_on_stream": [i32, vp, C.POINTER(vp)],
            "agrs_ctx_destroy": [vp],
            "agrs_sync": [vp],
            "agrs_device_info": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(sz), C.POINTER(sz)],
            "agrs_alloc": [vp, sz, C.POINTER(vp)],
            "agrs_free": [vp, vp],
            "agrs_pool_reserve": [vp, sz],
            "agrs_pool_trim": [vp],
            "agrs_pool_stats": [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)],
            "agrs_host_alloc": [vp, sz, C.POINTER(vp)],
            "agrs_host_free": [vp, vp],
            "agrs_upload": [vp, vp, vp, sz],
            "agrs_download": [vp, vp, vp, sz],
            "agrs_event_create": [vp, C.POINTER(vp)],
            "agrs_event_destroy": [vp, vp],
            "agrs_download_async": [vp, vp, vp, sz, vp],
            "agrs_event_wait": [vp, vp],
            "agrs_upload_prefetch_event": [vp, vp, vp, sz, vp],
            "agrs_wait_event": [vp, vp],
            "agrs_copy": [vp, vp, vp, sz],
            "agrs_upload_prefetch": [vp, vp, vp, sz],
            "agrs_prefetch_join": [vp],
            "agrs_fill_f32": [vp, vp, f32, sz],
            "agrs_gemm_f32": [vp, i32, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, vp],
            "agrs_gemm_bias_act_f32": [vp, i32, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, vp, i32, vp, i64],
            "agrs_last_gemm_act_fused": [vp],
            "agrs_set_gemm_path": [vp, i32],
            "agrs_gemm_operand_absmax": [vp, vp, vp],
            "agrs_gemm_wants_absmax": [vp, i64, i64, i64],
            "agrs_absmax_f32": [vp, vp, sz, vp],
            "agrs_next_output_absmax": [vp, vp],
            "agrs_last_gemm_path": [vp],
            "agrs_set_tuning": [vp, i32, i64],
            "agrs_get_tuning": [vp, i32, C.POINTER(i64)],
            "agrs_peer_alloc": [vp, sz, C.POINTER(vp)],
            "agrs_peer_free": [vp, vp],
            "agrs_peer_export": [vp, vp, vp],
            "agrs_peer_open": [vp, vp, C.POINTER(vp)],
            "agrs_peer_close": [vp, vp],
            "agrs_fused_create": [vp, i32, i32, C.POINTER(vp), sz, sz, sz, sz, i32, C.POINTER(vp)],
            "agrs_fused_weights_offset": [vp, C.POINTER(sz)],
            "agrs_fused_exposed_ms": [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)],
            "agrs_fused_info": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
            "agrs_fused_begin_step": [vp, f32],
            "agrs_fused_bucket_ready": [vp, i32, sz, sz],
            "agrs_fused_finish_step": [vp],
            "agrs_fused_destroy": [vp],
            "agrs_fused_sgd_step": [vp, f32],
            "agrs_fused_scatter_f32": [vp, vp, sz, sz],
            "agrs_gemm_f32_scatter": [vp, i32, i32, i64, i64, i64, vp, i64, vp, i64, vp, vp, vp, sz],
            "agrs_capture_begin": [vp],
            "agrs_capture_end": [vp, C.POINTER(vp)],
            "agrs_graph_launch": [vp, vp, i32],
            "agrs_graph_destroy": [vp, vp],
            "agrs_timer_begin": [vp],
            "agrs_timer_end": [vp, C.POINTER(f32)],
            "agrs_profile_enable": [vp, i32],
            "agrs_profile_gemm": [vp, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)],
            "agrs_comm_unique_id": [vp],
            "agrs_comm_create": [vp, i32, i32, vp, C.POINTER(vp)],
            "agrs_comm_destroy": [vp],
            "agrs_comm_info": [vp, C.POINTER(i32), C.POINTER(i32)],
            "agrs_comm_allreduce_f32": [vp, vp, sz, i32],
            "agrs_comm_join": [vp],
            "agrs_relu_fwd_f32": [vp, vp, vp, sz],
            "agrs_relu_bwd_f32": [vp, vp, vp, vp, sz],
            "agrs_sigmoid_fwd_f32": [vp, vp, vp, sz],
            "agrs_sigmoid_bwd_f32": [vp, vp, vp, vp, sz],
            "agrs_mse_fwd_f32": [vp, vp, vp, sz, vp],
            "agrs_mse_bwd_f32": [vp, vp, vp, vp, f32, sz, vp],
            "agrs_relu_bwd_colsum_f32": [vp, vp, vp, vp, i64, i64, vp],
            "agrs_sigmoid_bwd_colsum_f32": [vp, vp, vp, vp, i64, i64, vp],
            "agrs_mse_bwd_colsum_f32": [vp, vp, vp, vp, f32, i64, i64, vp, vp],
            "agrs_binary_f32": [vp, i32, vp, vp, vp, sz, sz, i32],
            "agrs_scalar_f32": [vp, i32, vp, vp, f32, sz],
            "agrs_neg_f32": [vp, vp, vp, sz],
            "agrs_powi_f32": [vp, vp, vp, i32, sz],
            "agrs_equal_f32": [vp, vp, vp, sz, C.POINTER(i32)],
            "agrs_transpose_f32": [vp, vp, i64, i64, vp],
            "agrs_sum_f32": [vp, vp, sz, vp],
            "agrs_mean_f32": [vp, vp, sz, vp],
            "agrs_sum_axis_f32": [vp, vp, i64, i64, i64, vp],
            "agrs_sgd_step_f32": [vp, vp, vp, f32, sz, sz],
            "agrs_sgd_step_multi_f32": [vp, i32, vp, vp, vp, vp, f32],
            "agrs_im2col_f32": [vp, vp, i64, i64, i64, i64, i64, i64, i64, vp],
            "agrs_col2im_f32": [vp, vp, i64, i64, i64, i64, i64, i64, vp],
            "agrs_broadcast_filter_f32": [vp, vp, i64, i64, i64, vp],
            "agrs_conv2d_fwd_f32": [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, vp],
            "agrs_conv2d_fwd_act_f32": [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, vp, i32, vp],
            "agrs_conv2d_bwd_f32": [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, vp, vp, vp],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = i32
        L.agrs_last_error.argtypes = [vp]
        L.agrs_last_error.restype = C.c_char_p
        L.agrs_stream.argtypes = [vp]
        L.agrs_stream.restype = vp
        L.agrs_launch_count.argtypes = [vp]
        L.agrs_launch_count.restype = C.c_uint64
        L.agrs_graph_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.agrs_graph_info.restype = i32
        L.agrs_fused_flag_floats.argtypes = []
        L.agrs_fused_flag_floats.restype = C.c_size_t
        L.agrs_fused_max_buckets.argtypes = []
        L.agrs_fused_max_buckets.restype = i32
        L.agrs_conv2d_implicit_ok.argtypes = [i64] * 8
        L.agrs_conv2d_implicit_ok.restype = i32
        L.agrs_version.argtypes = []
        L.agrs_version.restype = C.c_char_p
        _lib = L
    return _lib


class Buf:
    """A device allocation owned by a Ctx (freed stream-ordered on __del__)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, nbytes
        p = C.c_void_p()
        ctx.check(lib().agrs_alloc(ctx.h, nbytes, C.byref(p)))
        self.ptr = p.value

    def __del__(self):
        try:
            if self.ptr and self.ctx.h:
                lib().agrs_free(self.ctx.h, self.ptr)
        except Exception:
            pass
        self.ptr = None


class Ctx:
    """Thin convenience wrapper used by tests and bench: numpy in, numpy out, every op a C-ABI call."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        L = lib()
        rc = L.agrs_ctx_create(device, C.byref(h)) if stream is None else L.agrs_ctx_create_on_stream(device, stream, C.byref(h))
        if rc != OK:
            raise AgrsError(rc, "agrs_ctx_create failed (no sm_100 device? there is no CPU fallback)")
        self.h = h
        self.L = L

    def close(self):
        if self.h:
            self.L.agrs_ctx_destroy(self.h)
            self.h = None

    def check(self, rc):
        if rc != OK:
            raise AgrsError(rc, self.L.agrs_last_error(self.h).decode())

    def call(self, name, *args):
        self.check(getattr(self.L, name)(self.h, *args))

    # memory helpers
    def empty(self, n_floats):
        return Buf(self, int(n_floats) * 4)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        b = Buf(self, max(arr.nbytes, 4))
        if arr.nbytes:
            self.call("agrs_upload", b.ptr, arr.ctypes.data, arr.nbytes)
            self.sync()  # pageable source: keep `arr` alive semantics simple
        return b

    def to_host(self, buf, shape, offset_floats=0):
        out = np.empty(shape, dtype=np.float32)
        if out.nbytes:
            self.call("agrs_download", out.ctypes.data, buf.ptr + 4 * offset_floats, out.nbytes)
        return out

    def sync(self):
        self.call("agrs_sync")

    def launches(self):
        return int(self.L.agrs_launch_count(self.h))

    def gemm(self, A, B, transA=False, transB=False, bias=None, path=GEMM_AUTO):
        """numpy convenience: returns op(A) @ op(B) (+bias) computed on the device."""
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        M, K = (A.shape[1], A.shape[0]) if transA else A.shape
        N = B.shape[0] if transB else B.shape[1]
        dA, dB = self.to_device(A), self.to_device(B)
        dbias = self.to_device(bias) if bias is not None else None
        dC = self.empty(max(M * N, 1))
        self.call("agrs_set_gemm_path", path)
        try:
            self.call("agrs_gemm_f32", int(transA), int(transB), M, N, K, dA.ptr, A.shape[1], dB.ptr, B.shape[1], dC.ptr, N,
                      dbias.ptr if dbias else None)
        finally:
            self.L.agrs_set_gemm_path(self.h, GEMM_AUTO)
        return self.to_host(dC, (M, N))
