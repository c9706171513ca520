"""GPU test of padded leading dimensions through agrs_gemm_f32 on every GEMM path (tcgen05 pair and one-CTA kernels with both
epilogues, SIMT, skinny): lda / ldb / ldc larger than the extents."""
import numpy as np
import pytest

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import _capi as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = K.Ctx(0)
    yield c
    c.close()


def padded(rng, rows, cols, pad):
    buf = rng.uniform(-2, 2, (rows, cols + pad)).astype(np.float32)
    return buf, buf[:, :cols]


@pytest.mark.parametrize("path,M,N,Kd", [(K.GEMM_TCGEN05, 520, 392, 1100), (K.GEMM_TCGEN05, 96, 264, 260), (K.GEMM_SIMT, 65, 63, 17),
                                         (K.GEMM_AUTO, 1024, 12, 300)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("pads", [(4, 8, 12), (0, 0, 16)])
@pytest.mark.parametrize("epi_tma", [1, 0])
def test_gemm_with_padded_leading_dimensions(ctx, path, M, N, Kd, ta, tb, pads, epi_tma):
    """lda / ldb / ldc larger than the extents (multiples of 4 so the tcgen05 path stays eligible): same values as the contiguous
    product, and the padding columns of C are left untouched"""
    pa, pb, pc = pads
    rng = np.random.default_rng(M + N + Kd + 7 * ta + 3 * tb + pa)
    Abuf, A = padded(rng, *((Kd, M) if ta else (M, Kd)), pa)
    Bbuf, B = padded(rng, *((N, Kd) if tb else (Kd, N)), pb)
    bias = rng.uniform(-1, 1, N).astype(np.float32)
    Cbuf = np.full((M, N + pc), 123.0, np.float32)
    dA, dB, dC, dbias = ctx.to_device(Abuf), ctx.to_device(Bbuf), ctx.to_device(Cbuf), ctx.to_device(bias)
    ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, epi_tma)
    ctx.call("agrs_set_gemm_path", path)
    try:
        ctx.call("agrs_gemm_f32", ta, tb, M, N, Kd, dA.ptr, Abuf.shape[1], dB.ptr, Bbuf.shape[1], dC.ptr, N + pc, dbias.ptr)
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)
        ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, 1)
    got = ctx.to_host(dC, Cbuf.shape)
    opA, opB = (A.T if ta else A).astype(np.float64), (B.T if tb else B).astype(np.float64)
    truth = opA @ opB + bias
    assert np.abs(got[:, :N] - truth).max() <= 2e-5 * np.abs(truth).max()
    assert np.all(got[:, N:] == 123.0)
