#!/bin/bash
# first GPU contact: kernel-level parity tests, then the tcgen05 GEMM diagnostics (one process per case)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/kernels.log
: > gpurun_out/gemm_check.log
run() { echo "## $*" >> gpurun_out/gemm_check.log; timeout 120 python tools/gemm_check.py "$@" >> gpurun_out/gemm_check.log 2>&1; echo "rc=$?" >> gpurun_out/gemm_check.log; }
# smallest possible case first: one tile, one k-block
run --path tc --ta 0 --tb 1 --m 128 --n 256 --k 32
run --path tc --ta 0 --tb 1 --m 128 --n 256 --k 32 --explicit-hi
run --path tc --ta 0 --tb 0 --m 128 --n 256 --k 32
run --path tc --ta 1 --tb 1 --m 128 --n 256 --k 32
run --path tc --ta 1 --tb 0 --m 128 --n 256 --k 32
run --path tc --ta 0 --tb 1 --m 128 --n 128 --k 32 --bn 128
for ta in 0 1; do for tb in 0 1; do
  run --path tc --ta $ta --tb $tb --m 512 --n 768 --k 256 --bias
  run --path tc --ta $ta --tb $tb --m 200 --n 300 --k 100
  run --path tc --ta $ta --tb $tb --m 2048 --n 2048 --k 2048 --iters 5
  run --path tc --ta $ta --tb $tb --m 2048 --n 2048 --k 2048 --iters 5 --bn 128
done; done
run --path tc --ta 0 --tb 0 --m 2048 --n 2048 --k 2048 --explicit-hi --iters 5
run --path tc --ta 0 --tb 0 --m 8192 --n 4096 --k 4096 --iters 5
run --path tc --ta 0 --tb 1 --m 8192 --n 4096 --k 4096 --iters 5
run --path tc --ta 1 --tb 0 --m 4096 --n 4096 --k 8192 --iters 5
run --path simt --ta 0 --tb 0 --m 2048 --n 2048 --k 2048 --iters 3
tail -5 gpurun_out/kernels.log; cat gpurun_out/gemm_check.log | tail -60
