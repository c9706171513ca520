#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gemm_sweep.py --pair 1 --sizes 1024,2048,4096 --linear --odd --splitk 0,1,2,3 --iters 5 > gpurun_out/sweep_splitk.jsonl 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/sweep_splitk.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/sweep_splitk.jsonl'):
    if l.startswith('{"M"'):
        r=json.loads(l)
        if r.get('ms'): print(r['M'],r['N'],r['K'],r['ta'],r['tb'],'splitk',r['splitk'],round(r['ms']*1e3,1),'us',round(r['tflops'],1),'TF/s',f"{r['rel_err_vs_f64']:.1e}")
PY
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'])"
