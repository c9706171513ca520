#!/bin/bash
# tcgen05 GEMM diagnostics round 2: MN-major layouts + K-chunked accumulation
mkdir -p gpurun_out
: > gpurun_out/gemm_check2.log
run() { echo "## $*" >> gpurun_out/gemm_check2.log; timeout 120 python tools/gemm_check.py "$@" >> gpurun_out/gemm_check2.log 2>&1; echo "rc=$?" >> gpurun_out/gemm_check2.log; }
for ta in 0 1; do for tb in 0 1; do
  run --path tc --ta $ta --tb $tb --m 128 --n 256 --k 32
  run --path tc --ta $ta --tb $tb --m 512 --n 768 --k 256 --bias
  run --path tc --ta $ta --tb $tb --m 200 --n 300 --k 100
  run --path tc --ta $ta --tb $tb --m 2048 --n 2048 --k 2048 --iters 5
done; done
run --path tc --ta 0 --tb 0 --m 136 --n 72 --k 40 --bias
run --path tc --ta 1 --tb 0 --m 1000 --n 520 --k 3000 --bias
for kc in 0 256 512 1024 2048; do
  run --path tc --ta 0 --tb 1 --m 8192 --n 4096 --k 4096 --iters 5 --kchunk $kc
done
run --path tc --ta 0 --tb 0 --m 8192 --n 4096 --k 4096 --iters 5
run --path tc --ta 1 --tb 0 --m 4096 --n 4096 --k 8192 --iters 5
run --path tc --ta 0 --tb 0 --m 8192 --n 8192 --k 8192 --iters 3
run --path tc --ta 0 --tb 1 --m 8192 --n 8192 --k 8192 --iters 3 --kchunk 512
run --path tc --ta 0 --tb 1 --m 4096 --n 4096 --k 16384 --iters 3
run --path tc --ta 0 --tb 1 --m 256 --n 256 --k 256 --iters 20
run --path tc --ta 0 --tb 1 --m 1024 --n 1024 --k 1024 --iters 20
run --path tc --ta 0 --tb 1 --m 4096 --n 4096 --k 4096 --iters 5
grep -c '"rel_err' gpurun_out/gemm_check2.log
