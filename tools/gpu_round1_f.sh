#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-200; tail -3 gpurun_out/bench.err
timeout 600 python tools/eltwise_bench.py > gpurun_out/eltwise.jsonl 2>&1; echo "eltwise rc=$?"; cut -c1-160 gpurun_out/eltwise.jsonl
timeout 300 python tools/tf32_cublas_peak.py > gpurun_out/cublas_peaks.json 2>&1; cat gpurun_out/cublas_peaks.json
CMD="python tools/gemm_sweep.py --pair 1 --sizes 8192 --majors 00 --iters 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:2cta -c 3 -o gpurun_out/prof_gemm_2cta_b $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
