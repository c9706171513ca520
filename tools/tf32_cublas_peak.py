"""Reference point for the GEMM roofline: what cuBLAS reaches on this GPU for bf16, plain TF32 (1 MMA per product, NOT f32-faithful)
and true f32 (SIMT) at 8192^3, burst (best of 10) and sustained (back to back for ~3 s).  torch is used here only as a way to call cuBLAS."""
import json
import time

import torch


def bench(fn, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = max(best, flops / (s.elapsed_time(e) * 1e-3) / 1e12)
    n = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    return best, n * flops / (s.elapsed_time(e) * 1e-3) / 1e12


def main():
    n = 8192
    fl = 2.0 * n ** 3
    a32 = torch.randn(n, n, device="cuda")
    b32 = torch.randn(n, n, device="cuda")
    c32 = torch.empty(n, n, device="cuda")
    a16, b16, c16 = a32.bfloat16(), b32.bfloat16(), torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
    out = {}
    out["bf16"] = bench(lambda: torch.matmul(a16, b16, out=c16), fl)
    torch.backends.cuda.matmul.allow_tf32 = True
    out["tf32"] = bench(lambda: torch.matmul(a32, b32, out=c32), fl)
    torch.backends.cuda.matmul.allow_tf32 = False
    out["f32_simt"] = bench(lambda: torch.matmul(a32, b32, out=c32), fl)
    print(json.dumps({k: {"burst_tflops": v[0], "sustained_tflops": v[1]} for k, v in out.items()}))


if __name__ == "__main__":
    main()
