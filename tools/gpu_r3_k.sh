#!/bin/bash
# Round 2, third session, GPU call K (1 GPU): mse_bwd + colsum -- loads first, power-of-two batch by multiplication: tests + HBM fractions
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "colsum or mse or power_of_two" > gpurun_out/r3k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3k_pytest.log
timeout 300 python tools/eltwise_bench.py > gpurun_out/r3k_eltwise.jsonl 2>&1; echo "eltwise rc=$?"; grep -E "colsum|\"mse_bwd\"" gpurun_out/r3k_eltwise.jsonl | cut -c1-160
