#!/bin/bash
# Round 2, GPU call Z (8 GPUs): the final tree at N = 8, fused copy-engine exchange against NCCL on the same box (C2 headline + C5)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); a = d["also"]["mlp8192"]
        print(f"C2 {d['ms_per_step']:.3f} ms {d['value']/1e6:.3f} M/s gemm {d['roofline']['achieved']:.0f} exposed {d.get('exchange_exposed_ms_per_step')} e2e {(d.get('e2e') or {}).get('ms_per_step')} | C5 {a['ms_per_step']:.2f} ms {a['value']/1e3:.0f} k/s gemm {a['roofline']['achieved']:.0f} exposed {a.get('exchange_exposed_ms_per_step')} | parity {d.get('parity', {}).get('ok')}")
PY
}
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2z_n8_fused.json 2> gpurun_out/r2z_n8_fused.err; echo "fused rc=$? $(show gpurun_out/r2z_n8_fused.json)"
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --nccl --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2z_n8_nccl.json 2> gpurun_out/r2z_n8_nccl.err; echo "nccl rc=$? $(show gpurun_out/r2z_n8_nccl.json)"
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2z_n8_fused2.json 2> gpurun_out/r2z_n8_fused2.err; echo "fused again rc=$? $(show gpurun_out/r2z_n8_fused2.json)"
