#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gemm or split or linear or mlp or fullsize" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2y_pytest.log
python tools/gemm_sweep.py --pair 1 --sizes 4096 --majors 10 --linear-c2 --odd --announce --auto --iters 10 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith(chr(123)):
        r=json.loads(l)
        if 'tflops' in r: print(r['M'],r['N'],r['K'],r['ta'],r['tb'], round(r['tflops'],1), r['rel_err_vs_f64'])
        elif 'cases' in r: print(r)
"
timeout 400 python bench.py --no-also > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r2y_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["gpu_launches"], d["parity"]["ok"], d["clocks"])
PY
