"""Do peer copies progress while the tcgen05 GEMM owns every SM?  GPU 0 runs back-to-back 8192^3 GEMMs (2.6 ms each) on the library's
stream; meanwhile a 32 MB copy to GPU 1 is issued on another stream, as a contiguous cudaMemcpyAsync and as the strided
cudaMemcpy2DAsync the fused exchange used first.  Reported: time from issue to completion of the copy (idle GPU: ~55 us)."""
import json
import os
import sys
import time

import numpy as np
from cuda.bindings import runtime as rt

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import _capi  # noqa: E402


def ck(r):
    err, *rest = r if isinstance(r, tuple) else (r,)
    if err != rt.cudaError_t.cudaSuccess:
        raise RuntimeError(str(err))
    return rest[0] if len(rest) == 1 else rest


def main():
    if ck(rt.cudaGetDeviceCount()) < 2:
        print(json.dumps({"error": "needs 2 GPUs"}))
        return
    nbytes = 32 << 20
    ck(rt.cudaSetDevice(1))
    dst = ck(rt.cudaMalloc(4 * nbytes))
    ck(rt.cudaSetDevice(0))
    ck(rt.cudaDeviceEnablePeerAccess(1, 0))
    src = ck(rt.cudaMalloc(4 * nbytes))
    loc = ck(rt.cudaMalloc(4 * nbytes))
    s = ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
    D2D = rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice
    ctx = _capi.Ctx(0)
    n = 8192
    rng = np.random.default_rng(0)
    dA, dB, dC = ctx.to_device(rng.uniform(-1, 1, (n, n)).astype(np.float32)), ctx.to_device(rng.uniform(-1, 1, (n, n)).astype(np.float32)), ctx.empty(n * n)
    ctx.call("agrs_set_gemm_path", _capi.GEMM_TCGEN05)

    def gemms(k):
        for _ in range(k):
            ctx.call("agrs_gemm_f32", 0, 0, n, n, n, dA.ptr, n, dB.ptr, n, dC.ptr, n, None)

    G = 64 << 10
    cases = {
        "1-D peer": lambda: ck(rt.cudaMemcpyAsync(dst, src, nbytes, D2D, s)),
        "2-D peer (64 KB rows, pitch x2)": lambda: ck(rt.cudaMemcpy2DAsync(dst, 2 * G, src, 2 * G, G, nbytes // G, D2D, s)),
        "1-D local": lambda: ck(rt.cudaMemcpyAsync(loc, src, nbytes, D2D, s)),
        "2-D local (64 KB rows, pitch x2)": lambda: ck(rt.cudaMemcpy2DAsync(loc, 2 * G, src, 2 * G, G, nbytes // G, D2D, s)),
    }
    gemms(2)
    ctx.sync()
    for name, fn in cases.items():
        for busy in (False, True):
            fn()
            ck(rt.cudaStreamSynchronize(s))
            ctx.sync()
            if busy:
                gemms(6)          # ~16 ms of back-to-back GEMMs in flight
                time.sleep(0.002)  # the first one is running now
            t0 = time.perf_counter()
            fn()
            ck(rt.cudaStreamSynchronize(s))
            dt = time.perf_counter() - t0
            ctx.sync()
            print(json.dumps({"copy": name, "MB": nbytes >> 20, "gemm_running": busy, "issue_to_done_us": round(dt * 1e6, 1)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
