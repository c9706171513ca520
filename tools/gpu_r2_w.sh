#!/bin/bash
# Round 2, GPU call W (1 GPU): evidence run on the final tree -- whole GPU suite, smoke(), the bench line, ncu launch list and full capture of the GEMM
# inside the step, the other configs, the C4 sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2w_smoke.log
timeout 600 python bench.py > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2w_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/r2w_bench_reference.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2w_launches.csv $CMD > gpurun_out/r2w_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_3xtf32_2cta --launch-skip 40 --launch-count 3 -f -o gpurun_out/r2w_gemm_final $CMD > gpurun_out/r2w_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2w_ncu_full.log
timeout 300 python tools/bench_configs.py --only c1 > gpurun_out/r2w_configs_c1.jsonl 2> gpurun_out/r2w_configs_c1.err; echo "c1 rc=$?"; cut -c1-260 gpurun_out/r2w_configs_c1.jsonl
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/r2w_configs_c3.jsonl 2> gpurun_out/r2w_configs_c3.err; echo "c3 rc=$?"; cut -c1-260 gpurun_out/r2w_configs_c3.jsonl
timeout 600 python tools/gemm_sweep.py --pair 1 --sizes 256,512,1024,2048,4096,8192,16384 --linear --announce --auto --graph --iters 10 > gpurun_out/r2w_gemm_sweep_c4.jsonl 2> gpurun_out/r2w_gemm_sweep_c4.err; echo "sweep rc=$?"; tail -1 gpurun_out/r2w_gemm_sweep_c4.jsonl
timeout 200 python tools/eltwise_bench.py > gpurun_out/r2w_eltwise.jsonl 2> gpurun_out/r2w_eltwise.err; echo "eltwise rc=$?"; tail -3 gpurun_out/r2w_eltwise.jsonl | cut -c1-200
