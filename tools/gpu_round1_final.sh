#!/bin/bash
# round-1 evidence pass on one GPU: tests, smoke, bench, sweep, elementwise, configs, then the two ncu passes of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 900 python tools/gemm_sweep.py --pair 1 --sizes 256,512,1024,2048,4096,8192,16384 --linear --iters 5 > gpurun_out/sweep_pair.jsonl 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/sweep_pair.jsonl
timeout 600 python tools/eltwise_bench.py > gpurun_out/eltwise.jsonl 2>&1; echo "eltwise rc=$?"; tail -1 gpurun_out/eltwise.jsonl | cut -c1-300
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2>&1; echo "configs rc=$?"; grep -c config gpurun_out/configs.jsonl
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_3xtf32_2cta -s 24 -c 3 -o gpurun_out/prof_gemm_step $BENCH > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
