#!/bin/bash
# Round 2, third session, GPU call O (1 GPU): kernel-level time of the Conv2D forward with and without the activation in its output stream
mkdir -p gpurun_out
timeout 60 python tools/conv_bench.py > gpurun_out/r3o_conv_bench.jsonl 2>&1; echo "conv rc=$?"; head -4 gpurun_out/r3o_conv_bench.jsonl | cut -c1-200
