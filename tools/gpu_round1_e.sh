#!/bin/bash
# round 1, call E: whole GPU suite on the pair kernel, bench, sweep, ncu full capture of the pair kernel (NN and TN)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 python tools/gemm_sweep.py --odd --sizes 256,512,1024,2048,4096,8192,16384 --linear --iters 5 > gpurun_out/sweep.log 2>&1
echo "sweep rc=$?"; tail -3 gpurun_out/sweep.log
CMD="python tools/gemm_sweep.py --pair 1 --sizes 4096 --majors 00,10 --iters 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:2cta -c 10 -o gpurun_out/prof_gemm_2cta $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
