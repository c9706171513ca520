#!/bin/bash
# Round 2, GPU call K: pair-GEMM variants (packed FMUL2/FADD2 split x0/x1, raw stage released by the splitters r0/r1, ring shapes R4L2 / R3L3)
# in the default mode (fp16 x3, magnitudes announced) and bf16 x3; then the GPU tests of the default build, conv bench, bench line.
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
for v in R4L2G2Ex0r0 R4L2G2Ex1r0 R4L2G2Ex0r1 R4L2G2Ex1r1 R3L3G2Ex0r0 R3L3G2Ex0r1 R3L3G2Ex1r1; do
  [ -f build_variants/libagrs_b200_$v.so ] || { echo "missing build_variants/libagrs_b200_$v.so"; continue; }
  cp build_variants/libagrs_b200_$v.so autograd.rs_b200/libagrs_b200.so
  XS="5"; case $v in R4L2G2Ex0r0|R4L2G2Ex1r1|R3L3G2Ex1r1) XS="5 3";; esac
  for X in $XS; do
    AGRS_TC_CROSS=$X timeout 120 python tools/gemm_sweep.py --pair 1 --sizes 8192 --linear-c2 --odd --announce --iters 10 > gpurun_out/r2k_${v}_cross$X.jsonl 2> gpurun_out/r2k_${v}_cross$X.err
    echo "$v cross=$X rc=$? $(summ gpurun_out/r2k_${v}_cross$X.jsonl)"
  done
done
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
timeout 200 python tools/conv_bench.py > gpurun_out/r2k_conv_bench.jsonl 2>&1; echo "conv bench rc=$?"; cat gpurun_out/r2k_conv_bench.jsonl
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r2k_configs_c3.jsonl 2> gpurun_out/r2k_configs_c3.err; echo "c3 rc=$?"; cut -c1-330 gpurun_out/r2k_configs_c3.jsonl
timeout 400 python bench.py --no-also > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2k_bench.json
