#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py --cpu 1 > gpurun_out/configs.jsonl 2>&1; echo "configs rc=$?"; grep config gpurun_out/configs.jsonl | cut -c1-420
