// Shared internal declarations for libisc_b200.so (not part of the public ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>

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

// Programmatic dependent launch (PDL) for the kernels of a decode step, which run back to back on one stream: each of them
// is launched with the programmatic-serialization attribute, calls pdl_trigger() at its top (its successor may then be
// scheduled as soon as every CTA of this grid has started) and pdl_wait() before its first access to global memory
// (returns once the predecessor grid has completed and its writes are visible). The successor's launch latency and
// set-up (barrier init, TMEM allocation, tensor-map prefetch) so overlap this kernel's tail. OFF by default: measured
// on the B = 1024 beam-3 call it is 3 % SLOWER (8.97 vs 8.71 ms per call, graph replay) — the early-resident successor
// CTAs cost more than the hidden launch latency. ISC_PDL=1 enables it (results are identical either way).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ISC_PDL");
    return e && atoi(e) != 0;
  }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}


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

// ACT_EXPNEG2_RELU: exp(-2 * relu(v)) — the representation of the projected attention features that the
// e-product tanh of the bf16x3 attention kernel reads (tanh2_eprod below)
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_EXPNEG2_RELU = 3, ACT_SIGMOID = 4 };

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

// 16-bit attention path: projected features are stored as fp16(exp(-2p) * kFastScale); fp16 keeps that a NORMAL number
// down to p = 10.05, so an image whose largest projected feature exceeds kFastPMax is flagged and read full width.
constexpr float kFastScale = 32768.0f;        // 2^15
constexpr float kFastPMax = 10.0f;
constexpr float kFastMinStored = 6.7540e-5f;  // 2^15 * exp(-2 * kFastPMax), above fp16's smallest normal 6.10e-5

// Optional fp16 copy of a GEMM's output for the attention kernel's 16-bit path (isc_feats_t::att16 / p_att16):
//   out[row][col] = fp16((expneg2 ? exp(-2 y) : y) * scale),   y = the epilogue's result,
// and flags[row / rows_per_flag] |= 1 when a stored value leaves [vmin, vmax] (or is NaN): that image is outside the
// 16-bit path's exact domain and the attention kernel reads its full-width features instead.
struct Half16Out {
  void* out = nullptr;  // __half [M, ld]
  int64_t ld = 0;
  float scale = 1.0f;
  int expneg2 = 0;
  float vmin = 0.0f, vmax = 65000.0f;
  int* flags = nullptr;
  int rows_per_flag = 1;
};

// Where a GEMM (or a pointwise kernel) writes its result; any member may be null.
struct Dest {
  float* f32 = nullptr;
  int64_t ld = 0;
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int64_t ldp = 0;
  Half16Out h16;
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

// Fused vocabulary-logit epilogue of the tensor-core GEMM: the [M,V] logits are never written. Every epilogue
// thread owns one row and one 128-column slice of a 128x256 tile and emits one record
//   {slice max, sum exp(x - max), -, -, v[k_sel], idx[k_sel]}     (sel_rec(k_sel) floats)
// with the k_sel largest UNMASKED logits of the slice (value desc, column asc). A small merge kernel turns the
// np = 2 * ceil(V / 256) records of a row into its log-softmax normaliser and its top-K (kernels_select.cu).
// Masks follow Captioner.sample (captioner.py:394-399): PAD/SOS/UNK when pad != eos, the previous word when
// decoding_constraint is set.
// Two record widths: 4 candidates per slice (beams up to 4, greedy) or 8 (beams 5..8).
constexpr int SEL_K_MAX = 8;
__host__ __device__ constexpr int sel_rec(int ks) { return 4 + 2 * ks; }  // floats per record
struct LogitsSelect {
  float* rec = nullptr;             // [M][np][sel_rec(k_sel)]
  int np = 0;
  int k_sel = 4;                    // 4 or 8
  const long long* last = nullptr;  // [M] previous word per row, or null
  int constraint = 0, mask_special = 0, pad_id = 0, sos_id = 0, unk_id = 0;
};

// sorted insertion into a (value desc) list; equal values keep their arrival order
template <int KS>
__device__ __forceinline__ void topk_insert(float (&v)[KS], int (&idx)[KS], float x, int n) {
  bool c[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) c[k] = x > v[k];
#pragma unroll
  for (int k = KS - 1; k > 0; --k) {
    v[k] = c[k] ? (c[k - 1] ? v[k - 1] : x) : v[k];
    idx[k] = c[k] ? (c[k - 1] ? idx[k - 1] : n) : idx[k];
  }
  v[0] = c[0] ? x : v[0];
  idx[0] = c[0] ? n : idx[0];
}

// C = act(A[M,K] · W[N,K]^T + bias + rowadd + addmat)
int gemm_simt(const Operand& A, const Operand& W, const Dest& C, int M, int N, int K,
              const Epilogue& ep, cudaStream_t stream);
int gemm_tc(const Operand& A, const Operand& W, const Dest& C, int M, int N, int K, int passes,
            const Epilogue& ep, cudaStream_t stream);
// LSTM cell fused into the gate GEMM's epilogue (nn.LSTMCell, gate order i,f,g,o): the [M,4H] pre-activations are
// never written. The weight planes W must be GATE-INTERLEAVED (split_planes_gate_interleaved: row = 16-unit group * 64 +
// gate * 16 + unit), so the B rows of an output tile — [i f g o] of units 16g..16g+15 per 64 columns — are contiguous
// (one TMA box per 128 columns) and the thread that owns a row and 64 accumulator columns holds all four gates of 16
// units. Tiles are 256 columns wide where that fills the SMs evenly and 128 wide for the rest of a row (lstm_tile).
struct LstmEpilogue {
  const int* parent = nullptr;   // [M] row of the previous state this row continues from (beam reorder), or null
  const float* c_prev = nullptr; // [*, H]
  float* h_out = nullptr;        // [M, H]
  float* c_out = nullptr;        // [M, H]
  __nv_bfloat16* x_hi = nullptr; // h also goes, as bf16 planes, into the next GEMM's operand [M, ldx] at column x_col
  __nv_bfloat16* x_lo = nullptr;
  long long ldx = 0;
  int x_col = 0;
  const unsigned char* mask = nullptr;  // dropout keep mask [M, H] on that copy only (captioner.py:182), or null
  float scale = 1.0f;
  // optional gathered addend: gates[row] += gather_tab[clamp(gather_idx[row]), :]  ([gather_rows, 4H] fp32)
  const float* gather_tab = nullptr;
  const long long* gather_idx = nullptr;
  int gather_rows = 0;
};
int gemm_tc_lstm(const Operand& A, const Operand& W, int M, int K, int passes, const float* bias, const float* rowadd,
                 int64_t ld_rowadd, int rows_per_group, const LstmEpilogue& lstm, cudaStream_t stream);

// Fused gate of Attention.forward (models/captioner.py:108-117) on the tensor-core GEMM:
//   g3 = tanh([c | s] W3^T + b3 + addmat),  w = sigmoid(alpha . g3 + alpha_b),  out = w c + (1 - w) s.
// A row's 512 gate columns are four 128-column tiles: the four CTAs of a thread-block cluster compute them, exchange
// their partial alpha . g3 sums through distributed shared memory and each mixes its 128 output columns — g3 never
// reaches HBM and the separate gate_mix launch disappears. `out` takes the mixed context (fp32 and / or bf16 planes).
struct GateEpilogue {
  const float* alpha = nullptr;    // [H]
  const float* alpha_b = nullptr;  // [1]
  const float* cs = nullptr;       // [M, ld_cs] fp32: content context at column 0, sentiment context at column H
  long long ld_cs = 0;
  float* gate_w = nullptr;         // optional [M] (stride ld_gate_w): the gate weight w
  long long ld_gate_w = 0;
};
int gemm_tc_gate(const Operand& A, const Operand& W, int M, int K, int passes, const float* bias, const float* addmat,
                 int64_t ld_addmat, const GateEpilogue& gate, const Dest& out, cudaStream_t stream);

// tensor-core GEMM whose A operand is fp32 in HBM: TMA brings fp32 tiles into a staging ring and four converter warps
// split them into the bf16 hi/lo operand tiles in shared memory (no plane-split pass over HBM). 128x256 tiles.
int gemm_tc_af32(const float* A, int64_t lda, const Operand& W, const Dest& C, int M, int N, int K, int passes,
                 const Epilogue& ep, cudaStream_t stream);

// 3x3 convolution (stride 1, padding 1) as ONE tensor-core GEMM. Activations live on a zero-bordered 16x16 grid per
// image: row = image * 256 + y * 16 + x, valid pixels at y, x in 1..14. The K loop walks 9 segments (dy, dx), each
// reading the A rows shifted by dy * 16 + dx through TMA; W is [N][9 * C] with K index ((dy+1)*3 + (dx+1)) * C + c.
// A is either fp32 (A32, split in-kernel by the converter warps) or bf16 planes. zero_border forces the output's border
// rows to zero so that it can feed the next convolution directly.
int gemm_tc_conv3x3(const float* A32, int64_t lda, const Operand& A, const Operand& W, const Dest& C, int M, int N, int C_in,
                    int passes, const Epilogue& ep, int zero_border, cudaStream_t stream);

// tensor-core GEMM whose epilogue emits LogitsSelect records instead of C (bias added; N = vocabulary)
int gemm_tc_logits(const Operand& A, const Operand& W, int M, int N, int K, int passes, const float* bias,
                   const LogitsSelect& sel, cudaStream_t stream);
inline int logits_slices(int V) { return 2 * ((V + 255) / 256); }
inline int gemm(int precision, const Operand& A, const Operand& W, const Dest& C, int M, int N,
                int K, const Epilogue& ep, cudaStream_t stream) {
  if (precision == ISC_PREC_FP32) return gemm_simt(A, W, C, M, N, K, ep, stream);
  return gemm_tc(A, W, C, M, N, K, precision == ISC_PREC_BF16X3 ? 3 : 1, ep, stream);
}

// fp32 [rows, cols] -> bf16 hi (and lo = bf16(x - hi) when lo != null)
// max_blocks > 0 caps the grid (used when the split shares the SMs with a persistent GEMM on another stream)
int split_planes(const float* src, int64_t ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo,
                 int64_t ld_dst, int64_t rows, int cols, cudaStream_t stream, int max_blocks = 0);
// the same for an LSTM weight [4 * hidden, cols] (gate-major rows i, f, g, o), planes written gate-interleaved
int split_planes_gate_interleaved(const float* src, long long ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long ld_dst,
                                  int hidden, int cols, cudaStream_t stream);

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
// tanh(p + q) from PRE-EXPONENTIATED operands: with ea = exp(-2p), eb = exp(-2q), e = ea * eb,
//   tanh(p + q) = (1 - e) / (1 + e).
// No ex2 in the inner loop, and two values share ONE reciprocal (1/d0 = d1 / (d0 d1)), so a value costs
// 0.5 MUFU + 6 FMA-pipe instructions instead of 2 MUFU + ~9. Absolute error ~2e-7 (relative error of e is
// a few ulp and |d tanh / d ln e| <= 1/2). Domain: ea in [0, ~7] (p >= -1; p is a ReLU output) and
// eb <= kExpClamp (q >= -20), so that d0 * d1 cannot overflow; an underflowing ea only occurs where
// tanh has saturated to 1 anyway.
constexpr float kExpClamp = 2.3e17f;  // ~exp(40)
__device__ __forceinline__ float exp_neg2(float x) { return fminf(expf(-2.0f * x), kExpClamp); }
__device__ __forceinline__ void tanh2_eprod(float ea0, float ea1, float eb0, float eb1, float& t0, float& t1) {
  const float e0 = ea0 * eb0, e1 = ea1 * eb1;
  const float d0 = 1.0f + e0, d1 = 1.0f + e1;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  const float i0 = r * d1, i1 = r * d0;
  t0 = fmaf(-e0, i0, i0);
  t1 = fmaf(-e1, i1, i1);
}
// The attention SCORE only enters a softmax over the regions, which ignores anything constant per query:
//   sum_j a_j tanh(p_j + q_j) = sum_j 2 a_j / (1 + e_j)  -  sum_j a_j,     e_j = ea_j * eb_j,
// so the kernel accumulates sum_j (2 a_j) / (1 + e_j) and never forms a tanh. Two terms share one reciprocal:
//   2a0/d0 + 2a1/d1 = (2a0 d1 + 2a1 d0) / (d0 d1),   d = fma(ea, eb, 1)
// — six FMA-pipe instructions and one MUFU per PAIR (tanh2_eprod + two FMAs: eleven). a0x2 / a1x2 are the doubled
// alphas. Same operand domain as tanh2_eprod (d0 * d1 cannot overflow).
__device__ __forceinline__ float score2_eprod(float ea0, float ea1, float eb0, float eb1, float a0x2, float a1x2, float acc) {
  const float d0 = fmaf(ea0, eb0, 1.0f), d1 = fmaf(ea1, eb1, 1.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  return fmaf(r, fmaf(a0x2, d1, a1x2 * d0), acc);
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_accurate(float x) { return 1.0f / (1.0f + expf(-x)); }
// sigmoid from the two raw MUFU ops (ex2, rcp): relative error ~3e-7 (a few ulp), five instructions instead of ~25
// (expf's range reduction and the IEEE division). Used by the LSTM cell fused into the gate GEMMs' epilogue, whose
// warps are otherwise busy for half of the kernel; -x * log2(e) below -126 flushes to 0 -> sigmoid = 1,
// above 128 gives inf -> rcp = 0.
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_TANH) return tanhf(v);
  if (act == ACT_EXPNEG2_RELU) return exp_neg2(fmaxf(v, 0.0f));
  if (act == ACT_SIGMOID) return sigmoid_accurate(v);
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
