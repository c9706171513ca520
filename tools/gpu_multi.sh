#!/bin/bash
# multi-GPU confirmation of the data-parallel path: bench at N ranks (as the driver launches it), overlap on and off
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dp_parity.py > gpurun_out/dp_parity_n$N.log 2>&1
echo "dp_parity rc=$?"; tail -5 gpurun_out/dp_parity_n$N.log
for extra in "" "--no-overlap" "--fused"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $extra \
    > gpurun_out/bench_n$N$extra.json 2> gpurun_out/bench_n$N$extra.err
  echo "bench N=$N $extra rc=$?"; cut -c1-400 gpurun_out/bench_n$N$extra.json; tail -5 gpurun_out/bench_n$N$extra.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --config mlp8192 \
    > gpurun_out/bench_n${N}_mlp8192.json 2> gpurun_out/bench_n${N}_mlp8192.err
echo "bench mlp8192 N=$N rc=$?"; cut -c1-400 gpurun_out/bench_n${N}_mlp8192.json; tail -5 gpurun_out/bench_n${N}_mlp8192.err
