#!/bin/bash
# Round 2, GPU call AB: suspend-time hint on EVERY mbarrier wait of the pair GEMM (splitters, MMA issuer as well), in-step effect
mkdir -p gpurun_out
cp autograd.rs_b200/libagrs_b200.so /tmp/libagrs_b200_default.so
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); a = d["also"]["mlp8192"]
        print(f"C2 {d['ms_per_step']:.3f} ms gemm {d['roofline']['achieved']:.1f} | C5 {a['ms_per_step']:.2f} ms gemm {a['roofline']['achieved']:.1f} | clocks {d['clocks']['sm_mhz']}")
PY
}
for v in default R4L2G2Ex1r1e4000t500a500 R4L2G2Ex1r1e4000t500a2000 default R4L2G2Ex1r1e4000t500a500; do
  if [ $v = default ]; then cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so; else cp build_variants/libagrs_b200_$v.so autograd.rs_b200/libagrs_b200.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2ab_$v.json 2> gpurun_out/r2ab_$v.err; echo "$v rc=$? $(show gpurun_out/r2ab_$v.json)"
done
cp build_variants/libagrs_b200_R4L2G2Ex1r1e4000t500a500.so autograd.rs_b200/libagrs_b200.so
python tools/gemm_sweep.py --pair 1 --sizes 2048 --majors 00,01,10,11 --odd --announce --auto --iters 3 | tail -1
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
