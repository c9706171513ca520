"""The operand-splitting schemes of the tcgen05 GEMM, modelled on the CPU (tools/split_numerics.py): representation error of each
scheme with exact products and sums.  Pins the ordering the design relies on and that every offered scheme stays
two orders of magnitude inside the 1e-4 contract before the tensor core's accumulator adds its own error.  CPU only."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    spec = importlib.util.spec_from_file_location("split_numerics", os.path.join(ROOT, "tools", "split_numerics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def errors(mod, a, b):
    res, truth = mod.schemes(a.astype(np.float32), b.astype(np.float32))
    scale = np.abs(truth).max()
    return {k: float(np.abs(c - truth).max() / scale) for k, (c, _) in res.items()}


def test_scheme_errors_are_ordered_and_small():
    mod = load()
    rng = np.random.default_rng(0)
    e = errors(mod, rng.uniform(-1, 1, (96, 2048)), rng.uniform(-1, 1, (2048, 80)))
    assert e["tf32 x1"] > 1e-4                                   # a single tf32 product is NOT f32-faithful
    assert e["fp16 x3, row/column scaled"] < e["tf32 x3 (cross 0)"] < e["tf32 + 2 x bf16 (cross 1, default)"] < e["bf16 x3"] < 1e-5
    assert e["tf32 + 2 x bf16 (cross 1, default)"] < 3e-6


def test_row_scaling_keeps_badly_scaled_rows_faithful():
    """gradients of 1e-7 whose rows differ by up to 2^20: fp16 needs the per-row scale, and with it every ROW keeps its accuracy"""
    mod = load()
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((64, 1024)) * 1e-7 * np.exp2(rng.integers(-20, 1, (64, 1)))).astype(np.float32)
    b = rng.uniform(-1, 1, (1024, 48)).astype(np.float32)
    res, truth = mod.schemes(a, b)
    rowscale = np.abs(truth).max(axis=1, keepdims=True)
    for name in ("fp16 x3, row/column scaled", "tf32 + 2 x bf16 (cross 1, default)", "bf16 x3"):
        assert float((np.abs(res[name][0] - truth) / rowscale).max()) < 2e-5, name
    assert float((np.abs(res["fp16 x3, row/column scaled"][0] - truth) / rowscale).max()) < 5e-7
