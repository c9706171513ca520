#!/bin/bash
# Round 2, GPU call M: what bounds the end-to-end number -- host-to-device copy rates on this box (idle and next to GEMMs)
mkdir -p gpurun_out
nvidia-smi -q | grep -i -A8 "GPU Link Info" | head -12
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv
lscpu | grep -i "model name\|numa\|^CPU(s)\|socket" | head -8
timeout 200 python tools/h2d_bench.py > gpurun_out/r2m_h2d_idle.jsonl 2> gpurun_out/r2m_h2d_idle.err; echo "h2d idle rc=$?"; cat gpurun_out/r2m_h2d_idle.jsonl; tail -3 gpurun_out/r2m_h2d_idle.err
timeout 200 python tools/h2d_bench.py --under-gemm > gpurun_out/r2m_h2d_gemm.jsonl 2> gpurun_out/r2m_h2d_gemm.err; echo "h2d under gemm rc=$?"; cat gpurun_out/r2m_h2d_gemm.jsonl; tail -3 gpurun_out/r2m_h2d_gemm.err
