#!/bin/bash
# two-GPU sanity pass after a kernel change: data-parallel parity (NCCL and fused exchange), then the bench line the driver asks for
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_parity.py > gpurun_out/dp_parity_n2_v3.log 2>&1
echo "dp_parity rc=$?"; grep -E "world=|PARITY|Panic|timeout|Error" gpurun_out/dp_parity_n2_v3.log | cut -c1-60,230-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
  > gpurun_out/bench_n2_v3.json 2> gpurun_out/bench_n2_v3.err
echo "bench N=2 rc=$?"; tail -1 gpurun_out/bench_n2_v3.json | cut -c1-200; tail -2 gpurun_out/bench_n2_v3.err | cut -c1-300
