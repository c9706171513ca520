#!/bin/bash
# Round 2, third session (2 GPUs): the tree with agrs_gemm_bias_act_f32 at N = 2 -- data-parallel parity (NCCL and fused exchange), then the bench line the driver asks for
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_parity.py > gpurun_out/r3_dp_parity_n2.log 2>&1
echo "dp_parity rc=$?"; grep -E "world=|PARITY|Panic|timeout|Error" gpurun_out/r3_dp_parity_n2.log | cut -c1-60,230-330
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r3_bench_n2.json 2> gpurun_out/r3_bench_n2.err
echo "bench N=2 rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r3_bench_n2.json"):
    if l.startswith("{"):
        d = json.loads(l); a = d["also"]["mlp8192"]
        print(f"C2 {d['ms_per_step']:.3f} ms {d['value']/1e6:.3f} M/s gemm {d['roofline']['achieved']:.0f} exposed {d.get('exchange_exposed_ms_per_step')} e2e {(d.get('e2e') or {}).get('ms_per_step')} | C5 {a['ms_per_step']:.2f} ms gemm {a['roofline']['achieved']:.0f} | parity {d.get('parity', {}).get('ok')} dp {d.get('parity', {}).get('data_parallel', {}).get('ok')}")
PY
tail -2 gpurun_out/r3_bench_n2.err | cut -c1-300
