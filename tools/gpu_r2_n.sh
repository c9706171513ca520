#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "async or no_host_sync or example or mlp_training" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 200 python tools/e2e_diag.py 12 > gpurun_out/r2_e2e_diag.jsonl 2> gpurun_out/r2_e2e_diag.err; echo "diag rc=$?"; cat gpurun_out/r2_e2e_diag.jsonl; tail -3 gpurun_out/r2_e2e_diag.err
timeout 400 python bench.py --no-also > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r2n_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["clocks"])
PY
tail -3 gpurun_out/r2n_bench.err
