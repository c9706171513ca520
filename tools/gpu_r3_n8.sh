#!/bin/bash
# Round 2, third session (8 GPUs): the final tree at N = 8, fused copy-engine exchange (C2 headline + C5), parity gates on
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --no-cpu-baseline --no-e2e > gpurun_out/r3_bench_n8.json 2> gpurun_out/r3_bench_n8.err; echo "bench N=8 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r3_bench_n8.json"):
    if l.startswith("{"):
        d = json.loads(l); a = d["also"]["mlp8192"]
        print(f"C2 {d['ms_per_step']:.3f} ms {d['value']/1e6:.3f} M/s gemm {d['roofline']['achieved']:.0f} exposed {d.get('exchange_exposed_ms_per_step')} | C5 {a['ms_per_step']:.2f} ms {a['value']/1e3:.0f} k/s gemm {a['roofline']['achieved']:.0f} exposed {a.get('exchange_exposed_ms_per_step')} | parity {d.get('parity', {}).get('ok')} dp {d.get('parity', {}).get('data_parallel', {}).get('ok')}")
PY
tail -2 gpurun_out/r3_bench_n8.err | cut -c1-300
