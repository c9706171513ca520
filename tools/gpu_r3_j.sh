#!/bin/bash
# Round 2, third session, GPU call J (1 GPU): mse_bwd + colsum with the upstream scalar read once per thread (reduce.cu prep): tests + HBM fractions
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py -q -x > gpurun_out/r3j_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r3j_pytest.log
timeout 300 python tools/eltwise_bench.py > gpurun_out/r3j_eltwise.jsonl 2>&1; echo "eltwise rc=$?"; grep -E "colsum|\"mse_bwd\"" gpurun_out/r3j_eltwise.jsonl | cut -c1-160
