"""GPU tests of agrs_gemm_bias_act_f32 (the forward activation written by the GEMM's epilogue next to the
pre-activation).  The contract is bit-identity with agrs_gemm_f32 followed by agrs_relu_fwd_f32 / agrs_sigmoid_fwd_f32 on every
path, fused or not -- the pre-activation Z, the activation Y and Y's magnitude word -- plus the oracle's Linear -> activation
(operators.rs:140-144 into :112-116 / :195-200) within the GEMM tolerance (1e-5 of max|Z|, inside the 1e-4 contract)."""
import ctypes as C

import numpy as np
import pytest

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import _capi as K
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

RELU, SIGMOID = 1, 2


@pytest.fixture(scope="module")
def ctx():
    c = K.Ctx(0)
    yield c
    c.close()


@pytest.fixture
def act_warps(ctx):
    """the pair kernel's activation warps are opt-in (AGRS_TUNE_TC_ACT_WARPS; DESIGN 4.1 has the measurement that made them so)"""
    ctx.call("agrs_set_tuning", K.TUNE_TC_ACT_WARPS, 1)
    yield
    ctx.call("agrs_set_tuning", K.TUNE_TC_ACT_WARPS, 0)


def run_pair(ctx, A, B, bias, act, ta=0, tb=0, announce=False, ldz=None, ldy=None, fused_call=True):
    """Z, Y, magnitude word of Y, fused flag -- through the one-call entry (fused_call) or the two separate entries."""
    M, Kd = (A.shape[1], A.shape[0]) if ta else A.shape
    N = B.shape[0] if tb else B.shape[1]
    ldz, ldy = ldz or N, ldy or N
    dA, dB, db = ctx.to_device(A), ctx.to_device(B), ctx.to_device(bias)
    # padding stays recognisable: the kernels must not touch it
    dZ, dY = ctx.to_device(np.full((M, ldz), 7.0, np.float32)), ctx.to_device(np.full((M, ldy), -7.0, np.float32))
    sa, sb, sy = ctx.empty(1), ctx.empty(1), ctx.to_device(np.zeros(1, np.float32))
    if announce:
        ctx.call("agrs_absmax_f32", dA.ptr, A.size, sa.ptr)
        ctx.call("agrs_absmax_f32", dB.ptr, B.size, sb.ptr)
        ctx.call("agrs_gemm_operand_absmax", sa.ptr, sb.ptr)
    fused = None
    if fused_call:
        ctx.call("agrs_next_output_absmax", sy.ptr)
        ctx.call("agrs_gemm_bias_act_f32", ta, tb, M, N, Kd, dA.ptr, A.shape[1], dB.ptr, B.shape[1], dZ.ptr, ldz, db.ptr, act, dY.ptr, ldy)
        fused = ctx.L.agrs_last_gemm_act_fused(ctx.h)
    else:
        assert ldz == N and ldy == N
        ctx.call("agrs_gemm_f32", ta, tb, M, N, Kd, dA.ptr, A.shape[1], dB.ptr, B.shape[1], dZ.ptr, ldz, db.ptr)
        ctx.call("agrs_next_output_absmax", sy.ptr)
        ctx.call("agrs_relu_fwd_f32" if act == RELU else "agrs_sigmoid_fwd_f32", dZ.ptr, dY.ptr, M * N)
    Z, Y = ctx.to_host(dZ, (M, ldz)), ctx.to_host(dY, (M, ldy))
    mag = ctx.to_host(sy, (1,)).view(np.uint32)[0]
    assert np.all(Z[:, N:] == 7.0) and np.all(Y[:, N:] == -7.0), "padding columns were written"
    return Z[:, :N].copy(), Y[:, :N].copy(), int(mag), fused


def operands(M, N, Kd, ta, tb, seed):
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (Kd, M) if ta else (M, Kd)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, Kd) if tb else (Kd, N)).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, N).astype(np.float32)
    return A, B, bias


def check_against_unfused(ctx, M, N, Kd, act, ta=0, tb=0, announce=False, expect_fused=None, seed=0):
    A, B, bias = operands(M, N, Kd, ta, tb, seed + M + 3 * N + 7 * Kd)
    Z1, Y1, m1, fused = run_pair(ctx, A, B, bias, act, ta, tb, announce)
    Z0, Y0, m0, _ = run_pair(ctx, A, B, bias, act, ta, tb, announce, fused_call=False)
    if expect_fused is not None:
        assert fused == int(expect_fused), (fused, expect_fused)
    assert np.array_equal(Z1.view(np.uint32), Z0.view(np.uint32)), "pre-activation differs from agrs_gemm_f32"
    assert np.array_equal(Y1.view(np.uint32), Y0.view(np.uint32)), "activation differs from the elementwise kernel"
    assert m1 == m0 == int(np.abs(Y0).max().view(np.uint32)), (m1, m0)
    # and the oracle: Linear::forward then the activation, f64 flavour as the arbiter of the GEMM
    opA, opB = (A.T if ta else A), (B.T if tb else B)
    truth = opA.astype(np.float64) @ opB.astype(np.float64) + bias.astype(np.float64)
    assert np.abs(Z1 - truth).max() / np.abs(truth).max() < 1e-5
    want = O.OpReLU().forward([O.OTensor(Z1)]).data if act == RELU else O.OpSigmoid().forward([O.OTensor(Z1)]).data
    if act == RELU:
        assert np.array_equal(Y1, want)
    else:
        assert np.allclose(Y1, want, rtol=4e-7, atol=1e-37)  # CUDA expf <= 2 ulp against numpy's exp (test_sigmoid's bound)
    return fused


# ------------------------------------------------------------------ SIMT kernel (tiny / unaligned shapes): always fused
@pytest.mark.parametrize("act", [RELU, SIGMOID])
@pytest.mark.parametrize("M,N,Kd", [(8, 128, 784), (1, 1, 1), (5, 7, 3), (200, 3, 17), (65, 130, 40), (64, 10, 2048)])
def test_simt_epilogue_writes_the_activation(ctx, M, N, Kd, act):
    """example.rs's first layer (batch 8, 784 -> 128, Sigmoid) and ragged shapes, including the split-K fold kernel (few tiles, long K)"""
    ctx.call("agrs_set_gemm_path", K.GEMM_SIMT)
    try:
        check_against_unfused(ctx, M, N, Kd, act, expect_fused=True)
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


# ------------------------------------------------------------------ pair kernel
@pytest.mark.parametrize("act", [RELU, SIGMOID])
@pytest.mark.parametrize("announce", [False, True])  # bf16 correction terms / fp16 x3 with announced magnitudes: the two default schemes
@pytest.mark.parametrize("M,N,Kd,splitk", [
    (512, 256, 256, 1),      # one K chunk: the epilogue holds the final value
    (512, 256, 1024, 1),     # exactly one chunk
    (512, 512, 4096, 1),     # four chunks: the last one fetches the folded partial sum
    (300, 200, 1100, 1),     # ragged rows, columns and K (two chunks, the second of three k-blocks)
    (2048, 2304, 2048, 0),   # 72 tiles: the planner's own choice is no split
    (1000, 36, 2080, 1),     # a single, clipped column block
])
def test_pair_epilogue_writes_the_activation(ctx, act_warps, M, N, Kd, splitk, announce, act):
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, splitk)
    try:
        check_against_unfused(ctx, M, N, Kd, act, announce=announce, expect_fused=True)
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 0)
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


@pytest.mark.parametrize("ta,tb", [(0, 1), (1, 0), (1, 1)])
def test_pair_epilogue_every_operand_major(ctx, act_warps, ta, tb):
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 1)
    try:
        check_against_unfused(ctx, 384, 320, 2560, RELU, ta=ta, tb=tb, announce=True, expect_fused=True)
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 0)
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


def test_pair_padded_outputs(ctx, act_warps):
    """row pitches beyond N on both outputs: the two tensor maps carry them, the padding stays untouched"""
    M, N, Kd = 512, 256, 2048
    A, B, bias = operands(M, N, Kd, 0, 0, 11)
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 1)
    try:
        Z1, Y1, m1, fused = run_pair(ctx, A, B, bias, SIGMOID, ldz=N + 8, ldy=N + 32)
        Z0, Y0, m0, _ = run_pair(ctx, A, B, bias, SIGMOID, fused_call=False)
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 0)
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)
    assert fused == 1
    assert np.array_equal(Z1, Z0) and np.array_equal(Y1, Y0) and m1 == m0


# ------------------------------------------------------------------ the second-pass fallback inside the same call
@pytest.mark.parametrize("act", [RELU, SIGMOID])
def test_split_k_falls_back_to_a_second_pass(ctx, act_warps, act):
    """K split over CTA pairs: no work item holds a final value, so the call runs the activation itself afterwards"""
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 2)
    try:
        check_against_unfused(ctx, 512, 256, 2048, act, announce=True, expect_fused=False)
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_SPLITK, 0)
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


def test_other_back_ends_fall_back(ctx, act_warps):
    """one-CTA tcgen05 kernel (M <= 128), skinny kernels (N <= 32), all-tf32 split mode, direct epilogue: second pass, same bits"""
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    try:
        check_against_unfused(ctx, 128, 256, 512, RELU, expect_fused=False)
        ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, 0)
        check_against_unfused(ctx, 512, 256, 512, RELU, expect_fused=False)
        ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, 5)
        ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, 0)
        check_against_unfused(ctx, 512, 256, 512, SIGMOID, expect_fused=False)
    finally:
        ctx.call("agrs_set_tuning", K.TUNE_TC_CROSS, 5)
        ctx.call("agrs_set_tuning", K.TUNE_TC_EPI_TMA, 1)
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)
    ctx.call("agrs_set_gemm_path", K.GEMM_SKINNY)
    try:
        check_against_unfused(ctx, 1024, 10, 3456, RELU, expect_fused=False)
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


def test_pair_kernel_default_is_the_second_pass(ctx):
    """without the opt-in the pair kernel leaves the activation to the second pass of the same call: same bits, fused flag 0"""
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    try:
        check_against_unfused(ctx, 512, 256, 2048, RELU, announce=True, expect_fused=False)
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)


def test_padded_fallback_and_argument_checks(ctx):
    M, N, Kd = 96, 40, 64
    A, B, bias = operands(M, N, Kd, 0, 0, 3)
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)  # M <= 128: the one-CTA kernel, which does not fuse
    try:
        Z1, Y1, m1, fused = run_pair(ctx, A, B, bias, RELU, ldz=N + 4, ldy=N + 12)
        Z0, Y0, m0, _ = run_pair(ctx, A, B, bias, RELU, fused_call=False)
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)
    assert fused == 0 and np.array_equal(Z1, Z0) and np.array_equal(Y1, Y0) and m1 == m0
    d = ctx.empty(64)
    L = ctx.L
    assert L.agrs_gemm_bias_act_f32(ctx.h, 0, 0, 4, 4, 4, d.ptr, 4, d.ptr, 4, d.ptr, 4, None, 3, d.ptr + 64, 4) == 1  # AGRS_ERR_INVALID: unknown act
    assert L.agrs_gemm_bias_act_f32(ctx.h, 0, 0, 4, 4, 4, d.ptr, 4, d.ptr, 4, d.ptr, 4, None, RELU, None, 4) == 1    # null Y
    assert L.agrs_gemm_bias_act_f32(ctx.h, 0, 0, 4, 4, 4, d.ptr, 4, d.ptr, 4, d.ptr, 4, None, RELU, d.ptr, 4) == 1   # Y aliases Z
    assert L.agrs_gemm_bias_act_f32(ctx.h, 0, 0, 0, 4, 4, d.ptr, 4, d.ptr, 4, d.ptr, 4, None, RELU, None, 4) == K.OK              # empty


def test_relu_special_values_through_the_epilogue(ctx, act_warps):
    """-0.0 and NaN pre-activations become +0.0 (functional_ndarray.rs:14-22), +inf stays: rows of A that produce them exactly"""
    M, N, Kd = 256, 64, 64
    A = np.zeros((M, Kd), np.float32)
    B = np.zeros((Kd, N), np.float32)
    B[0, :] = 1.0
    A[0, 0] = 3.0
    A[1, 0] = -2.0
    A[2, 0] = np.float32("nan")
    bias = np.zeros(N, np.float32)
    bias[5] = -0.0
    ctx.call("agrs_set_gemm_path", K.GEMM_TCGEN05)
    try:
        Z, Y, mag, fused = run_pair(ctx, A, B, bias, RELU)
    finally:
        ctx.call("agrs_set_gemm_path", K.GEMM_AUTO)
    assert fused == 1
    assert np.all(Y[0] == 3.0) and np.all(Y[1] == 0.0) and not np.signbit(Y[1]).any()
    assert np.isnan(Z[2]).all() and np.all(Y[2] == 0.0)
    assert np.all(Y[3:] == 0.0) and not np.signbit(Y[3:]).any()
    assert mag == int(np.float32(3.0).view(np.uint32))
