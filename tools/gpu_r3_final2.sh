#!/bin/bash
# Round 2, third session (1 GPU): the tree with agrs_gemm_bias_act_f32 -- whole GPU suite, smoke(), the bench line, the reference arm,
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3z_smoke.log
timeout 600 python bench.py > gpurun_out/r3z_bench_n1.json 2> gpurun_out/r3z_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r3z_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3z_bench_reference.json 2> gpurun_out/r3z_bench_reference.err; echo "reference arm rc=$?"; cut -c1-160 gpurun_out/r3z_bench_reference.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3z_launches.csv $CMD > gpurun_out/r3z_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
