#!/bin/bash
# bf16x3 mode of the pair GEMM (AGRS_TUNE_TC_CROSS 3) against the default on one GPU: parity + timing, exact known-answer tests
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
big = [r for r in rows if "tflops" in r and r["M"] * r["N"] * r["K"] >= 4096 ** 3]
print(" ".join(f"{r['tflops']:.1f}" for r in big), "| maxerr", max([r["rel_err_vs_f64"] for r in rows if "M" in r] or [-1]), "| cases", len([r for r in rows if "M" in r]))
PY
}
AGRS_TC_CROSS=3 timeout 200 python tools/gemm_sweep.py --pair 1 --sizes 4096,8192 --linear-c2 --odd --iters 10 > gpurun_out/cross3.jsonl 2> gpurun_out/cross3.err
echo "cross=3 rc=$? $(summ gpurun_out/cross3.jsonl)"; tail -c 300 gpurun_out/cross3.err
AGRS_TC_CROSS=1 timeout 200 python tools/gemm_sweep.py --pair 1 --sizes 4096,8192 --linear-c2 --iters 10 > gpurun_out/cross1_ref.jsonl 2> gpurun_out/cross1_ref.err
echo "cross=1 rc=$? $(summ gpurun_out/cross1_ref.jsonl)"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "cross" > gpurun_out/pytest_cross3.log 2>&1
echo "pytest cross rc=$?"; tail -4 gpurun_out/pytest_cross3.log
AGRS_TC_CROSS=3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cross3.json 2> gpurun_out/bench_cross3.err
echo "bench cross=3 rc=$?"; cut -c1-170 gpurun_out/bench_cross3.json
