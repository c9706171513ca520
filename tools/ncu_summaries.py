"""Turns the two ncu passes of a GPU evidence run into the text summaries kept under profiles/.
  python tools/ncu_summaries.py launches gpurun_out/launches.csv "<command line>"   > profiles/..._launches_summary.txt
  python tools/ncu_summaries.py full gpurun_out/prof.ncu-rep "<command line>" [--traffic-json out.json] > profiles/..._ncu_full_summary.txt
(the second form shells out to `ncu -i ... --page raw --csv`)"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "launch__block_size", "launch__cluster_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_elapsed.avg.per_second", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed"]


def launches(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    u = hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        t = float(r[v].replace(",", ""))
        t = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}.get(r[u], t)
        tot[r[k]] += t
        cnt[r[k]] += 1
    total = sum(tot.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {cmd} (cold-cache, serialised launches: compare SHARES)")
    for name in sorted(tot, key=lambda n: -tot[n]):
        print(f"{tot[name]:12.1f} us {cnt[name]:5d} x {100 * tot[name] / total:6.1f}%  {name[:150]}")
    print(f"total us {total:.1f}")


def full(path, cmd, traffic_json=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(cmd)
    print("Kernel Name", [r[hdr.index("Kernel Name")][:70] for r in data])
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            print(m, f"[{units[i]}]", [r[i] for r in data])
    if traffic_json:
        def col(m):
            i = hdr.index(m)
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
            return [float(r[i].replace(",", "")) * scale for r in data]
        rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
        per = (sum(rd) + sum(wr)) / len(rd)
        alg = 4 * (8192 * 4096 + 4096 * 4096 + 8192 * 4096)
        json.dump({"kernel": "tc2::k_gemm_3xtf32_2cta, the C2 Linear GEMMs (8192x4096x4096) inside bench.py's step", "dram_bytes_per_launch": per,
                   "dram_read_bytes": rd, "dram_write_bytes": wr, "algorithmic_bytes_per_launch": alg, "ratio": per / alg, "source": cmd},
                  open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
        full(sys.argv[2], sys.argv[3], tj)
