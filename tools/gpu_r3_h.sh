#!/bin/bash
# Round 2, third session, GPU call H (1 GPU): AGRS_TC_L2_HINT=2 (op(B) evict_last) WITH a persisting L2 carve-out (AGRS_L2_PERSIST_MB): DRAM bytes, then step time
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(f"C2 {d['ms_per_step']:.3f} ms {d['value']/1e3:.1f} k/s gemm {d['roofline']['achieved']:.0f} clocks {d['clocks']['sm_mhz']} MHz {d['clocks'].get('power_w_max')} W")
PY
}
for mb in 48 72; do
  AGRS_L2_PERSIST_MB=$mb AGRS_TC_L2_HINT=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_gemm_3xtf32_2cta --launch-skip 36 --launch-count 6 --csv --log-file gpurun_out/r3h_dram_persist$mb.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/r3h_ncu_persist$mb.log 2>&1; echo "ncu persist $mb rc=$?"; grep "persisting" gpurun_out/r3h_ncu_persist$mb.log | head -1
done
i=0
for cfg in "0 0" "72 1" "48 1" "0 0" "72 1" "48 1"; do
  set -- $cfg; i=$((i+1))
  AGRS_L2_PERSIST_MB=$1 AGRS_TC_L2_HINT=$2 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 20 > gpurun_out/r3h_bench_${i}_p$1_h$2.json 2> gpurun_out/r3h_bench_${i}_p$1_h$2.err; echo "persist $1 hint $2 rc=$? $(show gpurun_out/r3h_bench_${i}_p$1_h$2.json)"
done
