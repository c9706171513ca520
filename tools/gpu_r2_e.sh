#!/bin/bash
# Round 2, GPU call E (1 GPU): the bucket protocol of the fused exchange with world = 1, then the whole suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -s -k "fused or data_parallel" > gpurun_out/r2_pytest_e0.log 2>&1; echo "pytest fused rc=$?"; tail -15 gpurun_out/r2_pytest_e0.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_e.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_e.log
