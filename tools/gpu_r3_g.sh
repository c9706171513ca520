#!/bin/bash
# Round 2, third session, GPU call G (1 GPU): L2 eviction policies on the pair GEMM's operand loads (AGRS_TC_L2_HINT: 0 off, 1 op(B) evict_last
# when it is re-read per group of tile rows and fits, 3 = that + A evict_first): C2 step time alternating, then DRAM bytes of three GEMMs per setting
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(f"C2 {d['ms_per_step']:.3f} ms {d['value']/1e3:.1f} k/s gemm {d['roofline']['achieved']:.0f} clocks {d['clocks']['sm_mhz']} MHz {d['clocks'].get('power_w_max')} W")
PY
}
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > gpurun_out/r3g_pytest.log 2>&1; echo "pytest gemm (hint 0) rc=$?"; tail -1 gpurun_out/r3g_pytest.log
AGRS_TC_L2_HINT=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > gpurun_out/r3g_pytest_h3.log 2>&1; echo "pytest gemm (hint 3) rc=$?"; tail -1 gpurun_out/r3g_pytest_h3.log
i=0
for h in 0 1 3 0 1 3 0 1; do
  i=$((i+1))
  AGRS_TC_L2_HINT=$h timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-also --steps 20 > gpurun_out/r3g_bench_${i}_hint$h.json 2> gpurun_out/r3g_bench_${i}_hint$h.err; echo "hint $h rc=$? $(show gpurun_out/r3g_bench_${i}_hint$h.json)"
done
for h in 0 1 3; do
  AGRS_TC_L2_HINT=$h timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_gemm_3xtf32_2cta --launch-skip 36 --launch-count 6 --csv --log-file gpurun_out/r3g_dram_hint$h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/r3g_ncu_hint$h.log 2>&1; echo "ncu hint $h rc=$?"
done
