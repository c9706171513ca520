#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/eltwise_bench.py > gpurun_out/eltwise.jsonl 2>&1; echo "eltwise rc=$?"; grep -E "colsum|im2col|col2im|below" gpurun_out/eltwise.jsonl | cut -c1-200
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2>&1; echo "configs rc=$?"; cut -c1-400 gpurun_out/configs.jsonl
