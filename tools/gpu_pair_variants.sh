#!/bin/bash
# One-GPU pass: (1) the pair GEMM's pipeline-shape variants (tools/build_pair_variants.sh) on the three Linear shapes of C2 with
# bf16 correction terms, (2) one `ncu --set full` capture of the default build on the same command.
mkdir -p gpurun_out
cp autograd.rs_b200/libagrs_b200.so /tmp/libagrs_b200_default.so
for f in build_variants/libagrs_b200_*.so; do
  v=$(basename $f .so); v=${v#libagrs_b200_}
  cp $f autograd.rs_b200/libagrs_b200.so
  for X in 1 0; do
    AGRS_TC_CROSS=$X timeout 120 python tools/gemm_sweep.py --pair 1 --sizes , --linear-c2 --iters 10 > gpurun_out/variant_${v}_cross$X.jsonl 2> gpurun_out/variant_${v}_cross$X.err
    echo "variant $v cross=$X rc=$? $(python - gpurun_out/variant_${v}_cross$X.jsonl <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print(" ".join(f"{r['tflops']:.1f}" for r in rows if "tflops" in r), "maxerr", max([r["rel_err_vs_f64"] for r in rows if "M" in r] or [-1]))
PY
)"
  done
done
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
CMD="python tools/gemm_sweep.py --pair 1 --sizes , --linear-c2 --iters 3"
AGRS_TC_CROSS=1 timeout 120 $CMD > gpurun_out/plain_cross1.log 2>&1 &&
AGRS_TC_CROSS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gemm_3xtf32_2cta -s 3 -c 2 -f -o gpurun_out/prof_gemm_cross1 $CMD > gpurun_out/ncu_cross1.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_cross1.log
