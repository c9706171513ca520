#!/bin/bash
# Round 2, GPU call A: pipeline-shape variants of the pair GEMM (one vs two splitter groups, early convert) in cross modes 1/3/4,
# then the default library's bench line.   gpurun --timeout 900 -- 'bash tools/gpu_r2_a.sh'
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
big = [r for r in rows if "tflops" in r and r["M"] * r["N"] * r["K"] >= 4096 ** 3]
print(" ".join(f"{r['tflops']:.1f}" for r in big), "| maxerr", max([r["rel_err_vs_f64"] for r in rows if "M" in r] or [-1]), "| cases", len([r for r in rows if "M" in r]))
PY
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
cp autograd.rs_b200/libagrs_b200.so /tmp/libagrs_b200_default.so
for v in R4L2G2 R4L2G1 R4L2G1E R4L2G2E; do
  [ -f build_variants/libagrs_b200_$v.so ] || { echo "missing build_variants/libagrs_b200_$v.so"; continue; }
  cp build_variants/libagrs_b200_$v.so autograd.rs_b200/libagrs_b200.so
  for X in 1 3 4; do
    AGRS_TC_CROSS=$X timeout 120 python tools/gemm_sweep.py --pair 1 --sizes 4096,8192 --linear-c2 --odd --iters 10 > gpurun_out/r2_${v}_cross$X.jsonl 2> gpurun_out/r2_${v}_cross$X.err
    echo "$v cross=$X rc=$? $(summ gpurun_out/r2_${v}_cross$X.jsonl)"
  done
done
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
timeout 300 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench (default) rc=$?"; cut -c1-400 gpurun_out/r2_bench_default.json
