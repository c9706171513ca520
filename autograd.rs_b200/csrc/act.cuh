// The forward activations as device functions, shared by the elementwise kernels (elementwise.cu) and by the GEMM epilogues that
// emit the activation next to the pre-activation (agrs_gemm_bias_act_f32: SURVEY 8 f2) -- one definition, so the fused and the
// unfused forward give the same bits.
#pragma once

// functional_ndarray.rs:14-22: x > 0 ? x : 0  (-0.0, NaN -> +0.0)
__device__ __forceinline__ float agrs_act_relu(float x) { return x > 0.f ? x : 0.f; }
// operators.rs:190-199: 1/(1+exp(-x)) in f32 (full-precision expf, IEEE divide)
__device__ __forceinline__ float agrs_act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
