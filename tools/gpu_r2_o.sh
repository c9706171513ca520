#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
timeout 200 python tools/e2e_timeline.py > gpurun_out/r2_e2e_timeline_copies_after.txt 2>&1; grep -c "<<<" gpurun_out/r2_e2e_timeline_copies_after.txt
timeout 400 python bench.py --no-also > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r2o_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["gpu_launches"], d["clocks"])
PY
tail -3 gpurun_out/r2o_bench.err
