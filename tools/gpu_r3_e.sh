#!/bin/bash
# Round 2, third session, GPU call E (1 GPU): activation warps without per-vector index arithmetic -- tests, then act warps on / off alternating
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_act_gpu.py -q -x > gpurun_out/r3e_pytest_act.log 2>&1; echo "pytest act rc=$?"; tail -2 gpurun_out/r3e_pytest_act.log
timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -k "forward_fusion or forward_pair" > gpurun_out/r3e_pytest_eng.log 2>&1; echo "pytest engine rc=$?"; tail -2 gpurun_out/r3e_pytest_eng.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); a = (d.get("also") or {}).get("mlp8192")
        print(f"C2 {d['ms_per_step']:.3f} ms {d['value']/1e3:.1f} k/s gemm {d['roofline']['achieved']:.0f} launches {d.get('gpu_launches')}" + (f" | C5 {a['ms_per_step']:.2f} ms gemm {a['roofline']['achieved']:.0f}" if a else ""))
PY
}
for i in 1 2; do
  timeout 600 python bench.py --no-cpu-baseline --no-e2e --act-warps > gpurun_out/r3e_bench_warps_$i.json 2> gpurun_out/r3e_bench_warps_$i.err; echo "act warps $i rc=$? $(show gpurun_out/r3e_bench_warps_$i.json)"
  timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r3e_bench_default_$i.json 2> gpurun_out/r3e_bench_default_$i.err; echo "default   $i rc=$? $(show gpurun_out/r3e_bench_default_$i.json)"
done
for i in 3 4 5; do
  timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 20 --act-warps > gpurun_out/r3e_bench_warps_$i.json 2> gpurun_out/r3e_bench_warps_$i.err; echo "act warps $i rc=$? $(show gpurun_out/r3e_bench_warps_$i.json)"
  timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 20 > gpurun_out/r3e_bench_default_$i.json 2> gpurun_out/r3e_bench_default_$i.err; echo "default   $i rc=$? $(show gpurun_out/r3e_bench_default_$i.json)"
done
