#!/bin/bash
# fused peer-memory optimiser at N ranks: parity (NCCL and fused), then bench fused vs NCCL
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dp_parity.py > gpurun_out/dp_parity_fused_n$N.log 2>&1
echo "dp_parity rc=$?"; grep -E "world=|PARITY|Panic|timeout" gpurun_out/dp_parity_fused_n$N.log | cut -c1-60,230-330
for extra in "--fused" "--nccl"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $extra \
    > "gpurun_out/bench_n${N}_$extra.json" 2> "gpurun_out/bench_n${N}_$extra.err"
  echo "bench N=$N $extra rc=$?"; tail -1 "gpurun_out/bench_n${N}_$extra.json" | cut -c1-170; tail -2 "gpurun_out/bench_n${N}_$extra.err" | cut -c1-300
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --config mlp8192 --fused --no-e2e \
    > gpurun_out/bench_n${N}_mlp8192_fused.json 2> gpurun_out/bench_n${N}_mlp8192_fused.err
echo "bench mlp8192 fused N=$N rc=$?"; tail -1 gpurun_out/bench_n${N}_mlp8192_fused.json | cut -c1-170
