#!/bin/bash
# One-GPU pass over the staged TMA epilogue of the pair GEMM: parity on ragged shapes, then timing with and without it, on the
# default ring split and on the alternative ones from tools/build_pair_variants.sh.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
big = [r for r in rows if "tflops" in r and r["M"] * r["N"] * r["K"] >= 4096 ** 3]
print(" ".join(f"{r['tflops']:.1f}" for r in big), "| maxerr", max([r["rel_err_vs_f64"] for r in rows if "M" in r] or [-1]), "| cases", len([r for r in rows if "M" in r]))
PY
}
AGRS_TC_CROSS=1 timeout 200 python tools/gemm_sweep.py --pair 1 --sizes 4096 --linear-c2 --odd --splitk 0,2 --iters 10 > gpurun_out/epi1_cross1.jsonl 2> gpurun_out/epi1_cross1.err
echo "epi=1 cross=1 rc=$? $(summ gpurun_out/epi1_cross1.jsonl)"; tail -c 300 gpurun_out/epi1_cross1.err
AGRS_TC_CROSS=0 timeout 200 python tools/gemm_sweep.py --pair 1 --sizes 4096 --linear-c2 --odd --iters 10 > gpurun_out/epi1_cross0.jsonl 2> gpurun_out/epi1_cross0.err
echo "epi=1 cross=0 rc=$? $(summ gpurun_out/epi1_cross0.jsonl)"
AGRS_TC_EPI_TMA=0 AGRS_TC_CROSS=1 timeout 200 python tools/gemm_sweep.py --pair 1 --sizes 4096 --linear-c2 --iters 10 > gpurun_out/epi0_cross1.jsonl 2> gpurun_out/epi0_cross1.err
echo "epi=0 cross=1 rc=$? $(summ gpurun_out/epi0_cross1.jsonl)"
cp autograd.rs_b200/libagrs_b200.so /tmp/libagrs_b200_default.so
for f in build_variants/libagrs_b200_*.so; do
  v=$(basename $f .so); v=${v#libagrs_b200_}
  cp $f autograd.rs_b200/libagrs_b200.so
  AGRS_TC_CROSS=1 timeout 120 python tools/gemm_sweep.py --pair 1 --sizes 4096 --linear-c2 --iters 10 > gpurun_out/epi1_${v}.jsonl 2> gpurun_out/epi1_${v}.err
  echo "variant $v epi=1 cross=1 rc=$? $(summ gpurun_out/epi1_${v}.jsonl)"
done
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
AGRS_TC_CROSS=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm" > gpurun_out/pytest_gemm_epi.log 2>&1
echo "pytest gemm rc=$?"; tail -4 gpurun_out/pytest_gemm_epi.log
