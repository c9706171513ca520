#!/bin/bash
# Round 2, GPU call X: wave alignment of the pair GEMM's TMA producers (AGRS_TC_WAVE_SYNC 0 off / 1 always / 2 large problems)
mkdir -p gpurun_out
for v in 0 1 0 1; do
  AGRS_TC_WAVE_SYNC=$v timeout 300 python tools/gemm_sweep.py --pair 1 --sizes 4096,8192,16384 --majors 00,10 --linear --announce --auto --iters 10 > gpurun_out/r2x_sync$v.jsonl 2> gpurun_out/r2x_sync$v.err
  echo "wave_sync=$v rc=$? $(python - gpurun_out/r2x_sync$v.jsonl <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print(" ".join(f"{r['M']}x{r['N']}x{r['K']}:{r['tflops']:.0f}" for r in rows if "tflops" in r), "| maxerr", max(r["rel_err_vs_f64"] for r in rows if "M" in r))
PY
)"
done
AGRS_TC_WAVE_SYNC=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_gemm_3xtf32_2cta --launch-skip 2 --launch-count 1 python tools/gemm_sweep.py --pair 1 --sizes 16384 --majors 00 --announce --auto --iters 2 2>&1 | grep -E "dram__|duration|hit_rate"
