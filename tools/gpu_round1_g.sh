#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline'])"; tail -3 gpurun_out/bench.err
timeout 600 python tools/eltwise_bench.py > gpurun_out/eltwise.jsonl 2>&1; echo "eltwise rc=$?"; grep -E "colsum|im2col|col2im|below" gpurun_out/eltwise.jsonl | cut -c1-200
