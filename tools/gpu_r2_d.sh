#!/bin/bash
# Round 2, GPU call D: GPU suite on the new defaults, the bench line, the ncu launch list of the same command, the C4 sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_d.log
timeout 500 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_d.err; cut -c1-300 gpurun_out/r2_bench_d.json
timeout 600 python tools/gemm_sweep.py --pair 1 --sizes 256,512,1024,2048,4096,8192,16384 --linear --announce --iters 10 > gpurun_out/r2_sweep_c4_announced.jsonl 2> gpurun_out/r2_sweep_c4_announced.err; echo "sweep announced rc=$?"; tail -1 gpurun_out/r2_sweep_c4_announced.jsonl
timeout 600 python tools/gemm_sweep.py --pair 1 --sizes 256,512,1024,2048,4096,8192,16384 --iters 10 > gpurun_out/r2_sweep_c4_plain.jsonl 2> gpurun_out/r2_sweep_c4_plain.err; echo "sweep plain rc=$?"; tail -1 gpurun_out/r2_sweep_c4_plain.jsonl
timeout 300 python tools/gemm_sweep.py --pair 1 --auto --sizes 256,512,1024 --announce --iters 50 > gpurun_out/r2_sweep_small_auto.jsonl 2>&1; echo "sweep small auto rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/r2_ncu_d.log 2>&1; echo "ncu rc=$?"
