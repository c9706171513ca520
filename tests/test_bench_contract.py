"""bench.py's contract on a machine without a GPU: the reference arm prints exactly one JSON line with the keys the driver reads,
and the product arm refuses to run (there is no CPU fallback to time).  CPU only; a few seconds."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, timeout=240):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")  # also on a GPU box: this file tests the no-device behaviour
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--config", "mlp_small")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]  # anything else a library prints goes to stderr
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "train_samples_per_sec" and line["unit"] == "samples/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None and "workload" in line["config"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_a_gpu():
    r = run("--steps", "1", "--warmup", "3", "--config", "mlp_small", timeout=120)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]  # no bench line is invented
    assert "sm_100" in r.stderr or "CUDA" in r.stderr or "device" in r.stderr
