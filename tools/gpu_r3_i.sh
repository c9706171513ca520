#!/bin/bash
# Round 2, third session, GPU call I (1 GPU): fused backward + column-sum kernels with all loads of four rows issued before any row is finished
# (reduce.cu ld4 / fin4) against the build before that change (build_variants/libagrs_b200_before_twophase.so): tests, HBM fractions, C2 / C5 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -x -k "colsum or sum_axis or mse or sigmoid or relu or mlp_training or fusion" > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r3i_pytest.log
timeout 300 python tools/eltwise_bench.py > gpurun_out/r3i_eltwise_new.jsonl 2>&1; echo "eltwise new rc=$?"; grep -E "colsum|\"mse_bwd\"|\"sigmoid_bwd\"" gpurun_out/r3i_eltwise_new.jsonl | cut -c1-160
cp autograd.rs_b200/libagrs_b200.so /tmp/new.so; cp build_variants/libagrs_b200_before_twophase.so autograd.rs_b200/libagrs_b200.so
timeout 300 python tools/eltwise_bench.py > gpurun_out/r3i_eltwise_before.jsonl 2>&1; echo "eltwise before rc=$?"; grep -E "colsum|\"mse_bwd\"|\"sigmoid_bwd\"" gpurun_out/r3i_eltwise_before.jsonl | cut -c1-160
cp /tmp/new.so autograd.rs_b200/libagrs_b200.so
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); a = (d.get("also") or {}).get("mlp8192")
        print(f"C2 {d['ms_per_step']:.3f} ms {d['value']/1e3:.1f} k/s gemm {d['roofline']['achieved']:.0f}" + (f" | C5 {a['ms_per_step']:.2f} ms gemm {a['roofline']['achieved']:.0f}" if a else ""))
PY
}
for i in 1 2; do
  timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r3i_bench_new_$i.json 2> gpurun_out/r3i_bench_new_$i.err; echo "new    $i rc=$? $(show gpurun_out/r3i_bench_new_$i.json)"
  cp build_variants/libagrs_b200_before_twophase.so autograd.rs_b200/libagrs_b200.so
  timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r3i_bench_before_$i.json 2> gpurun_out/r3i_bench_before_$i.err; echo "before $i rc=$? $(show gpurun_out/r3i_bench_before_$i.json)"
  cp /tmp/new.so autograd.rs_b200/libagrs_b200.so
done
