#!/bin/bash
# Round 2, third session, GPU call L (1 GPU): Conv2D + activation as one launch (agrs_conv2d_fwd_act_f32): tests, then C3 with / without, alternating
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -k "conv or lenet or fusion" > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3l_pytest.log
for i in 1 2; do
  timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r3l_c3_fused_$i.jsonl 2> gpurun_out/r3l_c3_fused_$i.err; echo "fused $i rc=$?"; python -c "
import json,sys
for l in open('gpurun_out/r3l_c3_fused_$i.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['config'], round(d['ms_per_step'],4), 'ms', d.get('kernels_per_step', d.get('graph_kernel_nodes')))"
  AGRS_NO_FWD_FUSION=1 timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r3l_c3_unfused_$i.jsonl 2> gpurun_out/r3l_c3_unfused_$i.err; echo "unfused $i rc=$?"; python -c "
import json,sys
for l in open('gpurun_out/r3l_c3_unfused_$i.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['config'], round(d['ms_per_step'],4), 'ms', d.get('kernels_per_step', d.get('graph_kernel_nodes')))"
done
