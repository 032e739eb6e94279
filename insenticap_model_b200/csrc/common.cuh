// Shared internal declarations for libisc_b200.so (not part of the public ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/isc.h"

namespace isc {

constexpr int H = 512;    // word_emb = feat_emb = rnn_hid = att_hid (opts.py:80-95)
constexpr int G4 = 4 * H; // LSTM gate width

void set_error(const char* fmt, ...);

#define ISC_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      isc::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                         \
    }                                                                         \
  } while (0)

#define ISC_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) {                                                  \
      isc::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (int)_e;                                                         \
    }                                                                         \
  } while (0)

#define ISC_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

#define ISC_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      isc::set_error(__VA_ARGS__);      \
      return ISC_ERR_ARG;               \
    }                                   \
  } while (0)

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

// Per-kernel-class accounting (isc_profile_* in include/isc.h). Every launcher opens a ProfScope:
// it counts the launch and, when profiling is enabled, brackets it with CUDA events on the launch
// stream and books the launch's ALGORITHMIC work (flops for GEMMs, HBM bytes for the rest).
struct ProfScope {
  ProfScope(int kclass, double work, cudaStream_t stream);
  ~ProfScope();
  int slot_;
  cudaStream_t stream_;
};

// A dense operand in both representations: fp32 (SIMT path) and bf16 hi/lo planes (tcgen05).
struct Operand {
  const float* f32 = nullptr;
  int64_t ld = 0;  // elements, fp32 view
  const __nv_bfloat16* hi = nullptr;
  const __nv_bfloat16* lo = nullptr;
  int64_t ldp = 0;  // elements, plane view
};

// Where a GEMM (or a pointwise kernel) writes its result; any member may be null.
struct Dest {
  float* f32 = nullptr;
  int64_t ld = 0;
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int64_t ldp = 0;
};

struct Epilogue {
  const float* bias = nullptr;    // [N]
  const float* rowadd = nullptr;  // [ceil(M / rows_per_group), ld_rowadd]: per-image additive term
  int64_t ld_rowadd = 0;
  int rows_per_group = 1;
  const float* addmat = nullptr;  // [M, ld_addmat]: per-row additive term
  int64_t ld_addmat = 0;
  int act = ACT_NONE;
};

// C = act(A[M,K] · W[N,K]^T + bias + rowadd + addmat)
int gemm_simt(const Operand& A, const Operand& W, const Dest& C, int M, int N, int K,
              const Epilogue& ep, cudaStream_t stream);
int gemm_tc(const Operand& A, const Operand& W, const Dest& C, int M, int N, int K, int passes,
            const Epilogue& ep, cudaStream_t stream);
inline int gemm(int precision, const Operand& A, const Operand& W, const Dest& C, int M, int N,
                int K, const Epilogue& ep, cudaStream_t stream) {
  if (precision == ISC_PREC_FP32) return gemm_simt(A, W, C, M, N, K, ep, stream);
  return gemm_tc(A, W, C, M, N, K, precision == ISC_PREC_BF16X3 ? 3 : 1, ep, stream);
}

// fp32 [rows, cols] -> bf16 hi (and lo = bf16(x - hi) when lo != null)
int split_planes(const float* src, int64_t ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo,
                 int64_t ld_dst, int64_t rows, int cols, cudaStream_t stream);

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// tanh with ~2e-7 ABSOLUTE error: (1 - e) / (1 + e), e = exp(-2|x|). Two MUFU ops (ex2, rcp) and no
// branch; the attention scores are sums of alpha_j * tanh(.), so absolute error is what matters.
__device__ __forceinline__ float tanh_ex2(float x) {
  float e, r;
  // e = 2^(-2|x| log2 e) <= 1: the raw MUFU op needs no range fix-up (underflow flushes to 0 -> tanh = 1)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.8853900817779268f * fabsf(x)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return copysignf((1.0f - e) * r, x);
}
// Four tanh values sharing ONE reciprocal: 1/d_i = (prod_j d_j)^-1 * prod_{j != i} d_j with d_i = 1 + e_i in
// [1,2]. 1.25 MUFU ops per value instead of 2 — the attention score loop is SFU/issue-bound otherwise.
// Inputs are PRE-SCALED: xs = 2*log2(e)*x, so e = 2^-|xs| = exp(-2|x|) needs no multiply.
constexpr float kTanhScale = 2.8853900817779268f;  // 2 * log2(e)
__device__ __forceinline__ void tanh4_ex2_scaled(const float xs[4], float t[4]) {
  float e[4], d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(-fabsf(xs[i])));
    d[i] = 1.0f + e[i];
  }
  const float p01 = d[0] * d[1], p23 = d[2] * d[3];
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
  const float r01 = r * p23, r23 = r * p01;
  const float inv[4] = {r01 * d[1], r01 * d[0], r23 * d[3], r23 * d[2]};
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = copysignf(fmaf(-e[i], inv[i], inv[i]), xs[i]);  // (1 - e) / (1 + e)
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_accurate(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace isc
