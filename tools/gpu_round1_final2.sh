#!/bin/bash
# round-1 closing evidence pass on one GPU after the GEMM changes (bf16 correction terms, staged TMA epilogue, 4+2 rings):
# tests, smoke, bench, GEMM sweep, then the two ncu passes of the bench command (each only after the plain command exited 0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python tools/gemm_sweep.py --pair 1 --sizes 1024,2048,4096,8192 --linear --iters 10 > gpurun_out/sweep_final2.jsonl 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/sweep_final2.jsonl
AGRS_TC_CROSS=0 timeout 200 python tools/gemm_sweep.py --pair 1 --sizes 4096 --linear-c2 --iters 10 > gpurun_out/sweep_final2_cross0.jsonl 2>&1; echo "sweep cross0 rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
if timeout 300 $BENCH > gpurun_out/plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_3xtf32_2cta -s 24 -c 3 -f -o gpurun_out/prof_gemm_step2 $BENCH > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
else
  echo "plain bench failed: no ncu pass"; tail -5 gpurun_out/plain.log
fi
