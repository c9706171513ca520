#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2aa_pytest.log
timeout 300 python tools/bench_configs.py --only c1 > gpurun_out/r2aa_configs_c1.jsonl 2> gpurun_out/r2aa_configs_c1.err; echo "c1 rc=$?"; cut -c1-200 gpurun_out/r2aa_configs_c1.jsonl
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r2aa_configs_c3.jsonl 2> gpurun_out/r2aa_configs_c3.err; echo "c3 rc=$?"; cut -c1-200 gpurun_out/r2aa_configs_c3.jsonl
timeout 400 python bench.py --no-also > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r2aa_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["gpu_launches"], d["parity"]["ok"], d["clocks"])
PY
