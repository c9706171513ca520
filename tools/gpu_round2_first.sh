#!/bin/bash
# First GPU call of the next round: one splitter group against two, on the final kernel.
#   here:        tools/build_pair_variants.sh "4 2 2" "4 2 1" "4 2 1 1" "4 2 2 1"   (writes build_variants/, shipped with the snapshot;
#                a 4th field of 1 = convert rows before waiting for the destination stage, suffix E)
#   on the box:  gpurun --timeout 420 -- 'bash tools/gpu_round2_first.sh'
# Per variant: parity + timing sweep in every cross mode, then -- with the one-group library swapped in -- the GPU suite and bench.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
big = [r for r in rows if "tflops" in r and r["M"] * r["N"] * r["K"] >= 4096 ** 3]
print(" ".join(f"{r['tflops']:.1f}" for r in big), "| maxerr", max([r["rel_err_vs_f64"] for r in rows if "M" in r] or [-1]), "| cases", len([r for r in rows if "M" in r]))
PY
}
cp autograd.rs_b200/libagrs_b200.so /tmp/libagrs_b200_default.so
for v in R4L2G2 R4L2G1 R4L2G1E R4L2G2E; do
  [ -f build_variants/libagrs_b200_$v.so ] || { echo "missing build_variants/libagrs_b200_$v.so"; continue; }
  cp build_variants/libagrs_b200_$v.so autograd.rs_b200/libagrs_b200.so
  for X in 1 3 4 0; do
    AGRS_TC_CROSS=$X timeout 150 python tools/gemm_sweep.py --pair 1 --sizes 4096 --linear-c2 --odd --iters 10 > gpurun_out/r2_${v}_cross$X.jsonl 2> gpurun_out/r2_${v}_cross$X.err
    echo "$v cross=$X rc=$? $(summ gpurun_out/r2_${v}_cross$X.jsonl)"
  done
done
for v in R4L2G1 R4L2G1E; do  # the whole GPU suite and the bench line with each one-group library swapped in
  [ -f build_variants/libagrs_b200_$v.so ] || continue
  cp build_variants/libagrs_b200_$v.so autograd.rs_b200/libagrs_b200.so
  timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_$v.log 2>&1; echo "pytest ($v) rc=$?"; tail -3 gpurun_out/r2_pytest_$v.log
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; echo "bench ($v) rc=$?"; cut -c1-170 gpurun_out/r2_bench_$v.json
done
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
timeout 300 python bench.py > gpurun_out/r2_bench_G2.json 2> gpurun_out/r2_bench_G2.err; echo "bench (default, two groups) rc=$?"; cut -c1-170 gpurun_out/r2_bench_G2.json
