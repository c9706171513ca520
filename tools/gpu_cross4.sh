#!/bin/bash
# fp16x3-with-scaling mode of the pair GEMM (AGRS_TUNE_TC_CROSS 4) on one GPU: parity + timing, exact known-answer tests, bench
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
big = [r for r in rows if "tflops" in r and r["M"] * r["N"] * r["K"] >= 4096 ** 3]
print(" ".join(f"{r['tflops']:.1f}" for r in big), "| maxerr", max([r["rel_err_vs_f64"] for r in rows if "M" in r] or [-1]), "| cases", len([r for r in rows if "M" in r]))
PY
}
AGRS_TC_CROSS=4 timeout 120 python tools/gemm_sweep.py --pair 1 --sizes 4096 --linear-c2 --odd --iters 10 > gpurun_out/cross4.jsonl 2> gpurun_out/cross4.err
echo "cross=4 rc=$? $(summ gpurun_out/cross4.jsonl)"; tail -c 400 gpurun_out/cross4.err
grep -h '"M"' gpurun_out/cross4.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['M'], r['N'], r['K'], r['ta'], r['tb'], 'err %.2e' % r['rel_err_vs_f64'], '%.1f' % r.get('tflops', 0))"
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "cross" > gpurun_out/pytest_cross4.log 2>&1
echo "pytest cross rc=$?"; tail -6 gpurun_out/pytest_cross4.log
AGRS_TC_CROSS=4 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cross4.json 2> gpurun_out/bench_cross4.err
echo "bench cross=4 rc=$?"; cut -c1-170 gpurun_out/bench_cross4.json; tail -2 gpurun_out/bench_cross4.err
