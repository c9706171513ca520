#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "conv or lenet or example" > gpurun_out/r2_pytest_i.log 2>&1; echo "pytest conv rc=$?"; tail -4 gpurun_out/r2_pytest_i.log
timeout 200 python tools/conv_bench.py > gpurun_out/r2_conv_bench.jsonl 2>&1; echo "conv bench rc=$?"; cat gpurun_out/r2_conv_bench.jsonl
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r2_configs_c3.jsonl 2> gpurun_out/r2_configs_c3.err; echo "c3 rc=$?"; cut -c1-330 gpurun_out/r2_configs_c3.jsonl
