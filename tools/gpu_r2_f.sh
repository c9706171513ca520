#!/bin/bash
# Round 2, GPU call F (N GPUs, default 2): data-parallel parity gate + bench through every exchange variant
N=${1:-2}
mkdir -p gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" \
    > gpurun_out/r2_n${N}_$name.json 2> gpurun_out/r2_n${N}_$name.err
  echo "$name rc=$? $(python - gpurun_out/r2_n${N}_$name.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    a = d.get("also", {}).get("mlp8192")
    print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 3) if d.get("e2e") else None,
          "| mlp8192 ms", round(a["ms_per_step"], 2) if a else None, "| parity", json.dumps(d.get("parity"))[:700])
except Exception as e:
    print("no line:", e)
PY
)"
  tail -3 gpurun_out/r2_n${N}_$name.err | cut -c1-300
}
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
run default X=1 -- --steps 10 --warmup 3
run trace AGRS_FUSED_TRACE=9 -- --steps 10 --warmup 3 --no-parity --no-e2e --no-also
grep "fused trace rank 0" gpurun_out/r2_n${N}_trace.err | sort -k8 -n | tail -14
run nccl X=1 -- --steps 10 --warmup 3 --nccl --no-parity --no-e2e
python - <<PY
import json
for n in ("default", "nccl"):
    d = json.loads([l for l in open("gpurun_out/r2_n${N}_%s.json" % n) if l.startswith("{")][-1])
    print(n, "exposed ms/step c2", d.get("exchange_exposed_ms_per_step"), "c5", d.get("also", {}).get("mlp8192", {}).get("exchange_exposed_ms_per_step"))
PY
