// Optional by-product of the kernels that write a whole tensor: the largest finite |out[i]| as its bit pattern
// (agrs_next_output_absmax).  The split GEMM's fp16 mode scales each operand by a power of two taken from its largest magnitude; a
// producer that already touches every element reports it for free instead of a separate pass over the tensor.  Non-finite values are
// left out (the scale must serve the finite ones; inf / NaN stay non-finite whatever the scale).  Unsigned compare == float compare
// for magnitudes, and a maximum does not depend on the order: deterministic.
#pragma once
#include "agrs_internal.h"

__device__ __forceinline__ uint32_t mag_bits(float v) {
  const uint32_t b = __float_as_uint(v) & 0x7FFFFFFFu;
  return b < 0x7F800000u ? b : 0u;
}
__device__ __forceinline__ uint32_t mag4(uint32_t m, float4 r) { return max(max(m, mag_bits(r.x)), max(max(mag_bits(r.y), mag_bits(r.z)), mag_bits(r.w))); }
__device__ __forceinline__ void mag_commit(uint32_t m, uint32_t* slot) {  // all 32 lanes of the warp must call
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(slot, m);
}
// Zeroing one magnitude word: a one-thread kernel, NOT cudaMemsetAsync.  A 4-byte memset is a copy-engine operation, and while a
// host-to-device input copy is in flight on another stream each one costs ~40 us of idle compute stream instead of ~5
// (profiles/r2_e2e_timeline.md: 13 of them per C2 step were the whole difference between the end-to-end and the device-resident loop).
static __global__ void k_zero_word(uint32_t* p) { *p = 0u; }
static inline int agrs_zero_word(agrs_ctx* ctx, uint32_t* slot) {
  k_zero_word<<<1, 1, 0, ctx->stream>>>(slot);
  AGRS_LAUNCHED(ctx);
  return AGRS_OK;
}
// the slot of a pending agrs_next_output_absmax request, zeroed on the stream and consumed (one request serves one call)
static inline int agrs_take_mx(agrs_ctx* ctx, uint32_t** slot) {
  *slot = ctx->next_out_mx;
  ctx->next_out_mx = nullptr;
  if (*slot) return agrs_zero_word(ctx, *slot);
  return AGRS_OK;
}
