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


def scale_of_max(max_bits):
    """host copy of scale_of_max() in csrc/gemm_tcgen05_2cta.cu (cross mode 4): (mult, inv) as float32"""
    e = min(max(int(max_bits) >> 23, 14), 254)
    mult = np.array([(267 - e) << 23], np.uint32).view(np.float32)[0]
    inv = np.array([(e - 13) << 23], np.uint32).view(np.float32)[0]
    return mult, inv


def test_fp16_scale_bit_formulas():
    """the power-of-two scale taken from the bit pattern of a row's largest magnitude: scaled maximum in [2^13, 2^14) (inside fp16's
    65504), mult * inv == 1 exactly, both factors normal floats for every possible maximum (zero, subnormal, huge, inf, NaN)"""
    rng = np.random.default_rng(2)
    mags = np.concatenate([np.exp2(rng.uniform(-110, 126, 2000)), [1.0, 1.9999999, 2.0, 3.0e38, 1e-30]]).astype(np.float32)
    for m in mags:
        mult, inv = scale_of_max(m.view(np.uint32))
        s = np.float32(m) * mult
        assert np.float32(2.0 ** 13) <= s < np.float32(2.0 ** 14), (m, s)
        assert mult * inv == np.float32(1.0)
    for bits in (0x00000000, 0x00000001, 0x007FFFFF, 0x06800000, 0x7F7FFFFF, 0x7F800000, 0x7FC00000):  # 0, subnormals, 2^-114, max, inf, NaN
        mult, inv = scale_of_max(bits)
        assert np.isfinite(mult) and np.isfinite(inv) and mult > 0 and inv > 0
        assert mult * inv == np.float32(1.0)
        assert (int(np.float32(mult).view(np.uint32)) >> 23) not in (0, 255) and (int(np.float32(inv).view(np.uint32)) >> 23) not in (0, 255)
        m = np.array([bits], np.uint32).view(np.float32)[0]
        if np.isfinite(m):
            assert abs(float(m) * float(mult)) < 2.0 ** 14  # never beyond fp16's range
