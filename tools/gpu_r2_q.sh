#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "conv or lenet" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest conv rc=$?"; tail -5 gpurun_out/r2q_pytest.log
timeout 200 python tools/conv_bench.py > gpurun_out/r2q_conv_bench.jsonl 2>&1; echo "conv bench rc=$?"; head -3 gpurun_out/r2q_conv_bench.jsonl; tail -1 gpurun_out/r2q_conv_bench.jsonl
AGRS_CONV_TILED=0 timeout 200 python tools/conv_bench.py 2>&1 | head -2
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r2q_configs_c3.jsonl 2> gpurun_out/r2q_configs_c3.err; echo "c3 rc=$?"; cut -c1-330 gpurun_out/r2q_configs_c3.jsonl
