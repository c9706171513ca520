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
// the slot of a pending agrs_next_output_absmax request, zeroed on the stream and consumed (one request serves one call)
static inline int agrs_take_mx(agrs_ctx* ctx, uint32_t** slot) {
  *slot = ctx->next_out_mx;
  ctx->next_out_mx = nullptr;
  if (*slot) AGRS_CUDA(ctx, cudaMemsetAsync(*slot, 0, sizeof(uint32_t), ctx->stream));
  return AGRS_OK;
}
