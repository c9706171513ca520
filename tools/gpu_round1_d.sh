#!/bin/bash
# round 1, call D: first contact of the cta_group::2 pair kernel -- smallest cases first, each under its own timeout
mkdir -p gpurun_out
L=gpurun_out/pair_check.log; : > $L
run() { echo "## $*" >> $L; timeout 100 python tools/gemm_check.py "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run --path tc --pair 1 --ta 0 --tb 1 --m 256 --n 256 --k 32
run --path tc --pair 1 --ta 0 --tb 1 --m 256 --n 256 --k 256
run --path tc --pair 1 --ta 0 --tb 0 --m 256 --n 256 --k 64
run --path tc --pair 1 --ta 1 --tb 1 --m 256 --n 256 --k 64
run --path tc --pair 1 --ta 1 --tb 0 --m 256 --n 256 --k 64
run --path tc --pair 1 --ta 0 --tb 1 --m 1024 --n 1024 --k 2048 --bias
tail -30 $L
if grep -q '"rel_err_vs_f64": [0-9.]*e-0[5-9]' $L; then
  timeout 900 python tools/gemm_sweep.py --odd --sizes 1024,4096 --linear --iters 5 > gpurun_out/pair_sweep.log 2>&1
  echo "sweep rc=$?"; tail -40 gpurun_out/pair_sweep.log
fi
nvidia-smi --help-query-gpu | grep -i -A1 "reasons" | head -40 > gpurun_out/smi_fields.txt 2>&1
