#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on -k regex:k_conv1_bwd_tiled --launch-skip 6 --launch-count 1 -f -o gpurun_out/r2_conv1_tiled python tools/conv_bench.py > gpurun_out/r2_ncu_conv_tiled.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_conv_tiled.log
