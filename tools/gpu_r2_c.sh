#!/bin/bash
# Round 2, GPU call C: fp16 split mode with magnitudes kept by the engine.  Whole GPU suite in mode 4, bench in modes 1 and 4
# (default pipeline and the early-convert build).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "absmax or magnitudes or skinny" > gpurun_out/r2_pytest_c0.log 2>&1; echo "pytest new rc=$?"; tail -5 gpurun_out/r2_pytest_c0.log
AGRS_TC_CROSS=4 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_c4.log 2>&1; echo "pytest cross4 rc=$?"; tail -8 gpurun_out/r2_pytest_c4.log
AGRS_TC_CROSS=4 timeout 400 python bench.py --no-also > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "bench cross4 rc=$?"; tail -2 gpurun_out/r2_bench_c4.err; cut -c1-260 gpurun_out/r2_bench_c4.json
cp autograd.rs_b200/libagrs_b200.so /tmp/default.so
cp build_variants/libagrs_b200_R4L2G2E.so autograd.rs_b200/libagrs_b200.so
AGRS_TC_CROSS=4 timeout 400 python bench.py --no-also --no-cpu-baseline > gpurun_out/r2_bench_c4E.json 2> gpurun_out/r2_bench_c4E.err; echo "bench cross4 E rc=$?"; cut -c1-260 gpurun_out/r2_bench_c4E.json
cp /tmp/default.so autograd.rs_b200/libagrs_b200.so
timeout 400 python bench.py --no-also > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; echo "bench cross1 rc=$?"; tail -2 gpurun_out/r2_bench_c1.err; cut -c1-260 gpurun_out/r2_bench_c1.json
