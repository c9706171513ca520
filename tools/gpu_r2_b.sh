#!/bin/bash
# Round 2, GPU call B: new full-size tests + the restructured bench (parity block, also.mlp8192)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_b.log
timeout 400 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_b.err; cat gpurun_out/r2_bench_b.json
