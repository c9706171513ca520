// Internal definitions shared by the kernels of libagrs_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <unordered_set>
#include <vector>

#include "agrs.h"

struct agrs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  cudaMemPool_t pool = nullptr;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  std::string err;
  int sticky = 0;  // sticky CUDA error class
  uint64_t launches = 0;
  int gemm_path = AGRS_GEMM_AUTO;
  int last_gemm_path = 0;
  int tc_bn = 256;         // tcgen05 tile width: 256 (2 smem stages) or 128 (3 stages)
  int tc_kchunk = 1024;    // K elements accumulated inside the tensor core before an RN fold into C (0: whole K)
  int tc_2cta = 1;         // use the cta_group::2 pair kernel when the problem has more than one 128-row block
  int tc_splitk = 0;       // pair kernel: 0 = choose the K split that fills the last wave, n > 0 = force n splits
  int tc_explicit_hi = 0;  // 1: splitter also writes the truncated hi tile back (does not rely on HW truncation)
  int tc_cross = 5;        // pair kernel split scheme (AGRS_TUNE_TC_CROSS): 0 tf32 x3, 1 / 2 tf32 + bf16 cross terms, 3 bf16 x3, 4 fp16 x3 scaled,
                           // 5 (default) = 4 when the caller announced operand magnitudes, else 1
  uint32_t func_attr_done = 0;  // which kernels of conv.cu already have their dynamic shared-memory limit raised on THIS context's device
  int conv_tiled = 1;      // Conv2D backward, single input channel: 1 register-tiled kernel where the shape allows (AGRS_CONV_TILED=0: the generic one)
  int tc_epi_tma = 1;      // pair kernel epilogue: 1 staged in shared memory + TMA store / reduce-add, 0 direct st.global from registers
  int tc_wave_sync = 2;    // pair kernel: the TMA producers of a wave start their tiles together (0 off, 1 always, 2 large problems; AGRS_TC_WAVE_SYNC)
  int tc_act_warps = 0;    // pair kernel, agrs_gemm_bias_act_f32: 1 = the activation warps write act(C) inside the GEMM (AGRS_TUNE_TC_ACT_WARPS), 0 = second pass
  int tc_l2_hint = 0;      // pair kernel: L2 eviction policies on the operand loads (env AGRS_TC_L2_HINT; see launch() in gemm_tcgen05_2cta.cu)
  int tc_reuse_scales = 0; // cross mode 4, measurement aid: 1 = reuse the operand magnitudes of the previous launch (same operands only)
  // operand magnitudes for the next GEMM call (agrs_gemm_operand_absmax) and the slot the next elementwise launch reports into
  // (agrs_next_output_absmax); both are consumed by the call they apply to
  const uint32_t* gemm_a_mx = nullptr;
  const uint32_t* gemm_b_mx = nullptr;
  uint32_t* next_out_mx = nullptr;
  // the activation request of the agrs_gemm_bias_act_f32 call in flight, as the GEMM back ends see it: a back end that can write
  // act(C) next to C from its epilogue does so and sets `fused`; otherwise the entry point runs the activation as a second pass
  struct GemmAct {
    int kind = 0;        // AGRS_ACT_NONE / AGRS_ACT_RELU / AGRS_ACT_SIGMOID
    float* y = nullptr;  // [M, N], row pitch ldy
    int64_t ldy = 0;
    bool fused = false;
  } gemm_act;
  int last_act_fused = 0;  // agrs_last_gemm_act_fused
  int64_t tc_min_flops = (1LL << 24) + 1;  // AUTO path: up to 256^3 MACs the SIMT kernel wins on launch + pipeline-fill latency (7.5 vs 12.1 us
                                           // of device time at 256^3; 512^3: 17.2 vs 13.9 -- profiles/r2_gemm_sweep_small_graph.txt)
  // small scratch for reductions (block partials + completion counter), allocated once
  float* scratch = nullptr;
  size_t scratch_bytes = 0;
  unsigned int* counter = nullptr;
  // per-output-block completion counters of the one-launch column sum (zero between launches: the last block resets its own)
  unsigned int* tickets = nullptr;
  size_t tickets_n = 0;
  // large workspace for split-K partials, grown on demand
  float* workspace = nullptr;
  size_t workspace_bytes = 0;
  // live GEMM timing (agrs_profile_*): event pairs recorded around launches while enabled
  struct ProfRec { cudaEvent_t start = nullptr, stop = nullptr; double flops = 0; int path = 0; };
  cudaEvent_t timer_a = nullptr, timer_b = nullptr;  // agrs_timer_*
  // agrs_capture_*: while capturing, allocations made inside the capture are tracked; frees of older buffers are deferred
  bool capturing = false;
  std::unordered_set<void*> capture_allocs;
  std::vector<void*> deferred_frees;
  uint64_t launches_at_capture = 0;
  cudaStream_t copy_stream = nullptr;                // agrs_upload_prefetch
  cudaEvent_t copy_ready = nullptr, copy_done = nullptr;
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  // driver entry point for TMA descriptor encoding (resolved lazily; no link-time libcuda dependency)
  void* encode_tiled = nullptr;
};

int agrs_fail(agrs_ctx* ctx, int code, const char* fmt, ...);

// NVTX range around a C-ABI entry point (tracing).  nvtx3 is header-only and binds to a profiler's injection
// library at run time: without a tool attached a range costs one predictable branch.
#include <nvtx3/nvToolsExt.h>
struct AgrsRange {
  explicit AgrsRange(const char* name) { nvtxRangePushA(name); }
  ~AgrsRange() { nvtxRangePop(); }
  AgrsRange(const AgrsRange&) = delete;
};
#define AGRS_RANGE(name) AgrsRange agrs_range_(name)

#define AGRS_CHECK_CTX(ctx)                 \
  do {                                      \
    if (!(ctx)) return AGRS_ERR_INVALID;    \
    if ((ctx)->sticky) return (ctx)->sticky; \
  } while (0)

#define AGRS_CUDA(ctx, expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      (ctx)->sticky = AGRS_ERR_CUDA;                                                          \
      return agrs_fail((ctx), AGRS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                                   \
    }                                                                                         \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define AGRS_LAUNCHED(ctx)                        \
  do {                                            \
    (ctx)->launches++;                            \
    AGRS_CUDA((ctx), cudaPeekAtLastError());      \
  } while (0)

// entry points that block or touch the device outside the stream cannot run while the stream is being captured
#define AGRS_NOT_CAPTURING(ctx, what)                                                                         \
  do {                                                                                                        \
    if ((ctx)->capturing) return agrs_fail((ctx), AGRS_ERR_UNSUPPORTED, "%s is not allowed while a CUDA graph is being captured", what); \
  } while (0)

#define AGRS_REQUIRE(ctx, cond, code, ...)                  \
  do {                                                      \
    if (!(cond)) return agrs_fail((ctx), (code), __VA_ARGS__); \
  } while (0)

int agrs_ensure_workspace(agrs_ctx* ctx, size_t nbytes);
// event pair around one GEMM launch while agrs_profile_enable is on (gemm.cu)
int agrs_prof_begin(agrs_ctx* ctx, int path, double flops, agrs_ctx::ProfRec* rec);
int agrs_prof_end(agrs_ctx* ctx, const agrs_ctx::ProfRec& rec);

// gemm back ends (gemm_simt.cu / gemm_tcgen05.cu)
int agrs_gemm_simt(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                   int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias);
bool agrs_gemm_tc_eligible(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                           const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias);
int agrs_gemm_tc(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias);

int agrs_gemm_tc_2cta(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                      int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias);

bool agrs_gemm_skinny_eligible(int transA, int transB, int64_t M, int64_t N, int64_t K);
int agrs_gemm_skinny(agrs_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                     const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias);

static inline bool agrs_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
