#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "prefetch or async or example or captured" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
timeout 200 python tools/e2e_diag.py 10 > gpurun_out/r2_e2e_diag.jsonl 2> gpurun_out/r2_e2e_diag.err; echo "diag rc=$?"; cat gpurun_out/r2_e2e_diag.jsonl; tail -3 gpurun_out/r2_e2e_diag.err
for i in 1 2; do
timeout 400 python bench.py --no-also > gpurun_out/r2p_bench$i.json 2> gpurun_out/r2p_bench$i.err; echo "bench rc=$?"; python - $i <<'PY'
import json, sys
for l in open(f"gpurun_out/r2p_bench{sys.argv[1]}.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["gpu_launches"], d["clocks"])
PY
done
