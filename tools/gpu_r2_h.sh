#!/bin/bash
# Round 2, GPU call H (1 GPU): whole suite, C1 / C3 configs, ncu full capture of the fp16 x3 GEMM inside the training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_h.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_h.log
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r2_configs_c3.jsonl 2> gpurun_out/r2_configs_c3.err; echo "c3 rc=$?"; cut -c1-330 gpurun_out/r2_configs_c3.jsonl
timeout 300 python tools/bench_configs.py --only c1 > gpurun_out/r2_configs_c1.jsonl 2> gpurun_out/r2_configs_c1.err; echo "c1 rc=$?"; cut -c1-330 gpurun_out/r2_configs_c1.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_3xtf32_2cta --launch-skip 40 --launch-count 1 -f -o gpurun_out/r2_gemm_fp16x3_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/r2_gemm_fp16x3_step.ncu-rep
