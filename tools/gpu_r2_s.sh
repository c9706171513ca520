#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "skinny or conv or lenet or gemm_small or linear or kat" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
python tools/c3_timeline.py > gpurun_out/r2_c3_timeline_graph.txt 2>&1; tail -22 gpurun_out/r2_c3_timeline_graph.txt | cut -c1-130
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r2s_configs_c3.jsonl 2> gpurun_out/r2s_configs_c3.err; echo "c3 rc=$?"; cut -c1-200 gpurun_out/r2s_configs_c3.jsonl
