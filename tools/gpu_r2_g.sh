#!/bin/bash
# fused exchange tests on one GPU, then (same box) nothing else
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "fused or data_parallel or colsum" > gpurun_out/r2_pytest_g.log 2>&1; echo "pytest fused rc=$?"; tail -5 gpurun_out/r2_pytest_g.log
