"""Peer copy bandwidth between GPU 0 and GPU 1 of one box through the copy engines, as the fused exchange issues them:
contiguous (cudaMemcpyAsync) against strided (cudaMemcpy2DAsync, rows of `G` bytes, pitch G * world), one stream and several.
One JSON line per case.  python tools/p2p_copy_bench.py"""
import json
import time

from cuda.bindings import runtime as rt


def ck(r):
    if isinstance(r, tuple):
        err, *rest = r
    else:
        err, rest = r, []
    if err != rt.cudaError_t.cudaSuccess:
        raise RuntimeError(str(err))
    return rest[0] if len(rest) == 1 else rest


def main():
    n = ck(rt.cudaGetDeviceCount())
    if n < 2:
        print(json.dumps({"error": "needs 2 GPUs"}))
        return
    total = 256 << 20
    ck(rt.cudaSetDevice(1))
    dst = ck(rt.cudaMalloc(total))
    ck(rt.cudaSetDevice(0))
    ck(rt.cudaDeviceEnablePeerAccess(1, 0))
    src = ck(rt.cudaMalloc(total))
    loc = ck(rt.cudaMalloc(total))
    streams = [ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking)) for _ in range(8)]
    e0, e1 = ck(rt.cudaEventCreate()), ck(rt.cudaEventCreate())
    D2D = rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice

    def timed(fn, reps=5):
        fn()
        ck(rt.cudaDeviceSynchronize())
        best = 1e9
        for _ in range(reps):
            ck(rt.cudaDeviceSynchronize())
            t0 = time.perf_counter()
            fn()
            ck(rt.cudaDeviceSynchronize())
            best = min(best, time.perf_counter() - t0)
        return best

    for mb in (8, 32, 128):
        nbytes = mb << 20
        for target, name in ((dst, "peer"), (loc, "local")):
            t = timed(lambda: ck(rt.cudaMemcpyAsync(target, src, nbytes, D2D, streams[0])))
            print(json.dumps({"case": f"1-D {name}", "MB": mb, "GBps": nbytes / t / 1e9}), flush=True)
            for k in (2, 4):
                part = nbytes // k
                t = timed(lambda: [ck(rt.cudaMemcpyAsync(target + i * part, src + i * part, part, D2D, streams[i])) for i in range(k)])
                print(json.dumps({"case": f"1-D {name}, {k} streams", "MB": mb, "GBps": nbytes / t / 1e9}), flush=True)
        for world in (2, 8):
            for G in (64 << 10, 1 << 20):
                rows = nbytes // G
                if rows * G * world > total:
                    rows = total // (G * world)
                moved = rows * G
                t = timed(lambda: ck(rt.cudaMemcpy2DAsync(dst, G * world, src, G * world, G, rows, D2D, streams[0])))
                print(json.dumps({"case": f"2-D peer, rows of {G >> 10} KB, pitch x{world}", "MB": moved >> 20, "GBps": moved / t / 1e9}), flush=True)
                t = timed(lambda: ck(rt.cudaMemcpy2DAsync(dst, G, src, G * world, G, rows, D2D, streams[0])))
                print(json.dumps({"case": f"2-D peer gather-to-contiguous, rows of {G >> 10} KB, src pitch x{world}", "MB": moved >> 20, "GBps": moved / t / 1e9}), flush=True)


if __name__ == "__main__":
    main()
