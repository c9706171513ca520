#!/bin/bash
# Round 2, GPU call V: suspend-time hints on the long mbarrier waits of the pair GEMM -- in-step effect on C2 and C5 (power-capped)
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
for v in default R4L2G2Ex1r1e1000t0 R4L2G2Ex1r1e1000t200 R4L2G2Ex1r1e4000t500 default; do
  if [ $v = default ]; then cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so; else cp build_variants/libagrs_b200_$v.so autograd.rs_b200/libagrs_b200.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2v_$v.json 2> gpurun_out/r2v_$v.err; echo "$v rc=$? $(show gpurun_out/r2v_$v.json)"
done
cp /tmp/libagrs_b200_default.so autograd.rs_b200/libagrs_b200.so
