#!/bin/bash
# Round 2, GPU call L: L2 prefetch distance (k-blocks ahead of the TMA loads) for the pair GEMM, default mode; same box comparison with p0.
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
for v in default R4L2G2Ex1r1p4 R4L2G2Ex1r1p8 R4L2G2Ex1r1p16 default; do
  if [ $v = default ]; then cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so; else cp build_variants/libagrs_b200_$v.so autograd.rs_b200/libagrs_b200.so; fi
  AGRS_TC_CROSS=5 timeout 120 python tools/gemm_sweep.py --pair 1 --sizes 8192 --linear --odd --announce --iters 10 > gpurun_out/r2l_${v}.jsonl 2> gpurun_out/r2l_${v}.err
  echo "$v rc=$? $(summ gpurun_out/r2l_${v}.jsonl)"
done
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
