#!/bin/bash
# Round 2, GPU call T (8 GPUs): aggregate host-to-device rate with all ranks copying at once, then the bench line at N = 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 $TR --nproc-per-node 8 --master-port 29511 tools/h2d_bench.py --iters 4 > gpurun_out/r2t_h2d_n8.jsonl 2> gpurun_out/r2t_h2d_n8.err; echo "h2d n8 rc=$?"; cat gpurun_out/r2t_h2d_n8.jsonl | cut -c1-220; tail -2 gpurun_out/r2t_h2d_n8.err
timeout 120 $TR --nproc-per-node 2 --master-port 29512 tools/h2d_bench.py --iters 4 > gpurun_out/r2t_h2d_n2.jsonl 2> gpurun_out/r2t_h2d_n2.err; echo "h2d n2 rc=$?"; head -2 gpurun_out/r2t_h2d_n2.jsonl | cut -c1-220
timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2t_bench_n8.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["parity"]["ok"], d["clocks"])
        a = d["also"]["mlp8192"]; print("mlp8192", a["value"], a["ms_per_step"], a["roofline"]["achieved"], a.get("exchange_exposed_ms_per_step"))
PY
tail -3 gpurun_out/r2t_bench_n8.err
