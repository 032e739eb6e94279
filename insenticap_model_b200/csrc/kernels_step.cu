// Non-GEMM kernels of one decode step (Captioner.forward_step, /root/reference/models/captioner.py:168-186)
// and of the prologue. All are HBM-bound element-wise / reduction kernels: 128-bit accesses,
// warp-shuffle reductions, one CTA per row (or per image for the attention, so the K beams of an
// image share one pass over its region features).
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include <type_traits>

#include "kernels.cuh"

namespace isc {

// --------------------------------------------------------------------------------------------
// embed + pack: builds the attention-LSTM input  X1[m] = [h_lang_prev | ReLU(E[it]) | h_att_prev]
// (captioner.py:170-174; the step-invariant terms — the fc slice and W_ih[:,2H:3H]·sl of
// xt = ReLU(E[it]) + sl — are hoisted into feats.pre_gates) and copies
// h_lang_prev into X2[m, 2H:3H] for the language LSTM's recurrent term. Rows are read through
// `parent` so that the beam reorder costs no separate gather.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_pack_kernel(const long long* __restrict__ it, const int* __restrict__ parent,
                                                         const float* __restrict__ h_in, long long M, int V,
                                                         const float* __restrict__ emb, RowDest x1, RowDest x2) {
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x;
  const int src = parent ? parent[m] : m;
  long long tok = it[m];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  const int c = threadIdx.x * 4;
  float4 hl = *reinterpret_cast<const float4*>(h_in + (1 * M + src) * H + c);
  float4 ha = *reinterpret_cast<const float4*>(h_in + (0 * M + src) * H + c);
  x1.store4(m, c, hl);
  if (emb) {
    float4 e = *reinterpret_cast<const float4*>(emb + tok * H + c);
    e.x = fmaxf(e.x, 0.f); e.y = fmaxf(e.y, 0.f); e.z = fmaxf(e.z, 0.f); e.w = fmaxf(e.w, 0.f);
    x1.store4(m, H + c, e);
    x1.store4(m, 2 * H + c, ha);
  } else {  // the word term is added from the xt_gates table by the gate GEMM's epilogue: X1 = [h_lang_prev | h_att_prev]
    x1.store4(m, H + c, ha);
  }
  x2.store4(m, 2 * H + c, hl);
}

// --------------------------------------------------------------------------------------------
// LSTM cell pointwise (nn.LSTMCell, gate order i,f,g,o): gates already hold W·x + biases.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) lstm_pointwise_kernel(const float* __restrict__ gates, const int* __restrict__ parent,
                                                             const float* __restrict__ c_prev, float* __restrict__ h_out,
                                                             float* __restrict__ c_out, RowDest extra, int extra_col,
                                                             const unsigned char* __restrict__ mask, float scale) {
  const int m = blockIdx.x;
  const int src = parent ? parent[m] : m;
  const int c = threadIdx.x * 4;
  const float* g = gates + (long long)m * G4;
  float4 gi = *reinterpret_cast<const float4*>(g + c);
  float4 gf = *reinterpret_cast<const float4*>(g + H + c);
  float4 gg = *reinterpret_cast<const float4*>(g + 2 * H + c);
  float4 go = *reinterpret_cast<const float4*>(g + 3 * H + c);
  float4 cp = *reinterpret_cast<const float4*>(c_prev + (long long)src * H + c);
  float4 cn, hn;
#define ISC_LSTM(X)                                                              \
  cn.X = sigmoid_accurate(gf.X) * cp.X + sigmoid_accurate(gi.X) * tanhf(gg.X);   \
  hn.X = sigmoid_accurate(go.X) * tanhf(cn.X);
  ISC_LSTM(x) ISC_LSTM(y) ISC_LSTM(z) ISC_LSTM(w)
#undef ISC_LSTM
  *reinterpret_cast<float4*>(c_out + (long long)m * H + c) = cn;
  *reinterpret_cast<float4*>(h_out + (long long)m * H + c) = hn;
  if (mask) {  // dropout on the copy that feeds the next GEMM (captioner.py:182); the recurrent state stays intact
    const uchar4 k = *reinterpret_cast<const uchar4*>(mask + (long long)m * H + c);
    hn.x *= k.x ? scale : 0.f; hn.y *= k.y ? scale : 0.f; hn.z *= k.z ? scale : 0.f; hn.w *= k.w ? scale : 0.f;
  }
  extra.store4(m, extra_col + c, hn);
}

// --------------------------------------------------------------------------------------------
// Attention (ContentAttention :23-35, SentiAttention :50-62): one CTA per image, R rows (beams)
// of that image processed against one pass over p_att / att (and p_sw / sw).
//   hproj[m] = [h2att(h) | h2word(h) | gate h2att(h)] (+ their biases), from the projection GEMM.
//
// The kernel streams both [L,H] feature tensors of its image once per decode step: their WIDTH is the HBM time of the
// step. Tensor-core precisions therefore read 16-bit copies (isc_feats_t::att16 / p_att16): fp16 keeps 11
// significant bits per value, and because a context vector is an average over L = 196 regions the rounding noise of its
// inputs shrinks by sqrt(L): the context moves by ~1e-5 relative, the size of the split-bf16 GEMMs' own error, and no
// golden token changes (tests/test_gpu_parity.py). Images with a value outside the 16-bit path's exact domain are
// flagged by the prologue and their CTA reads the full-width tensors (the "wide" path below).
// --------------------------------------------------------------------------------------------
template <typename FeatT>
struct FeatLoad;
template <>
struct FeatLoad<float> {
  static constexpr int kChunks = 4;  // 4 x float4 per lane per 512-wide row
  __device__ static __forceinline__ int col(int lane, int i) { return i * 128 + lane * 4; }
  __device__ static __forceinline__ void load_shared(const float* row, int lane, int i, float* v) {
    float4 t = *reinterpret_cast<const float4*>(row + col(lane, i));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void load_shared2(const float* row, int lane, int i, float2* v) {  // kWidth / 2 pairs
    float4 t = *reinterpret_cast<const float4*>(row + col(lane, i));
    v[0] = make_float2(t.x, t.y);
    v[1] = make_float2(t.z, t.w);
  }
  static constexpr int kWidth = 4;
};
template <>
struct FeatLoad<__nv_bfloat16> {
  static constexpr int kChunks = 2;  // 2 x (8 bf16 = 16 B) per lane
  __device__ static __forceinline__ int col(int lane, int i) { return i * 256 + lane * 8; }
  __device__ static __forceinline__ void load_shared(const __nv_bfloat16* row, int lane, int i, float* v) {
    uint4 t = *reinterpret_cast<const uint4*>(row + col(lane, i));
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 f = __bfloat1622float2(p[q]);
      v[2 * q] = f.x;
      v[2 * q + 1] = f.y;
    }
  }
  __device__ static __forceinline__ void load_shared2(const __nv_bfloat16* row, int lane, int i, float2* v) {
    uint4 t = *reinterpret_cast<const uint4*>(row + col(lane, i));
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = __bfloat1622float2(p[q]);
  }
  static constexpr int kWidth = 8;
};
template <>
struct FeatLoad<__half> {
  static constexpr int kChunks = 2;  // 2 x (8 fp16 = 16 B) per lane
  __device__ static __forceinline__ int col(int lane, int i) { return i * 256 + lane * 8; }
  __device__ static __forceinline__ void load_shared(const __half* row, int lane, int i, float* v) {
    uint4 t = *reinterpret_cast<const uint4*>(row + col(lane, i));
    const __half2* p = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 f = __half22float2(p[q]);  // HADD2.F32: full-rate, no conversion pipe
      v[2 * q] = f.x;
      v[2 * q + 1] = f.y;
    }
  }
  __device__ static __forceinline__ void load_shared2(const __half* row, int lane, int i, float2* v) {
    uint4 t = *reinterpret_cast<const uint4*>(row + col(lane, i));
    const __half2* p = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = __half22float2(p[q]);
  }
  static constexpr int kWidth = 8;
};

// Scoring modes (TANH_MODE):
//   0  libdevice tanhf on p + q                                   (ISC_PREC_FP32)
//   1  e-product, fast: pv = exp(-2p) * 2^15 (fp16), q = exp(-2q) * 2^-15 with q clamped at -20 (eb <= e^40). Exact on the
//      domain the prologue's flag guarantees (p <= 10): then p + q <= -10 wherever the clamp bites and tanh is -1 either way.
//   2  tanh.approx.f32 on p + q                                   (ISC_PREC_BF16, wide path)
//   3  e-product, wide: pv = exp(-2p) (fp32), q = exp(-2q) clamped at 1.6e38 (q >= -44), the PRODUCT e clamped at 1e18
//      (tanh is -1 to fp32 precision from e = 1e8 on): no operand combination overflows d0 * d1, and tanh(p + q) is right
//      for every (p, q) with p <= 43 (exp(-2p) a normal float) and (q >= -44 or p <= 34). Two more FMA-pipe instructions
//      per value than mode 1; only flagged images, the sentiment words (11 rows) and the training tape take it.
constexpr float kWideExpClamp = 1.6e38f;
__device__ __forceinline__ float exp_neg2_wide(float x) { return fminf(expf(-2.0f * x), kWideExpClamp); }
__device__ __forceinline__ float score2_eprod_wide(float ea0, float ea1, float eb0, float eb1, float a0x2, float a1x2, float acc) {
  const float e0 = fminf(ea0 * eb0, 1e18f), e1 = fminf(ea1 * eb1, 1e18f);
  const float d0 = 1.0f + e0, d1 = 1.0f + e1;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
  return fmaf(r, fmaf(a0x2, d1, a1x2 * d0), acc);
}

// Packed fp32 arithmetic (Blackwell FFMA2 / FMUL2: two fp32 operations per instruction). A three-register FFMA issues
// every SECOND cycle per scheduler on this architecture (register-port bound), so a kernel whose inner loops are fp32
// multiply-adds — this one: ncu had its fma pipe at 39 % "of peak", i.e. 78 % of what three-register FFMAs can reach —
// doubles its arithmetic rate by pairing them.
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
// Four values of the softmax-equivalent score at once (score2_eprod's algebra on register pairs): with the natural pairs
// (v0, v1), (v2, v3) of a converted half2, d01 = 1 + ea01 * eb01 and d23 likewise; values 0 / 2 and 1 / 3 share a
// reciprocal: D = d01 * d23 = (d0 d2, d1 d3), N = a01 * d23 + a23 * d01 = (a0 d2 + a2 d0, a1 d3 + a3 d1), acc += N / D.
// Six packed instructions and two MUFU for four values (fourteen scalar ones before).
__device__ __forceinline__ float2 score4_eprod(float2 ea01, float2 ea23, float2 eb01, float2 eb23, float2 a01, float2 a23,
                                               float2 acc) {
  const float2 one = make_float2(1.0f, 1.0f);
  const float2 d01 = fma2(ea01, eb01, one), d23 = fma2(ea23, eb23, one);
  const float2 D = mul2(d01, d23);
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(D.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(D.y));
  const float2 n = fma2(a01, d23, mul2(a23, d01));
  return fma2(r, n, acc);
}

// RT > 0: rows per image known at compile time (fully unrolled); RT == 0: runtime R <= 8.
template <typename FeatT, int TANH_MODE, int RT>
__device__ __forceinline__ void score_one(const float (&pv)[FeatLoad<FeatT>::kChunks][FeatLoad<FeatT>::kWidth],
                                          const float (&al)[FeatLoad<FeatT>::kChunks][FeatLoad<FeatT>::kWidth], int l,
                                          int n_items, int R, const float* __restrict__ q_smem,
                                          float* __restrict__ score_smem, int lane) {
  using L = FeatLoad<FeatT>;
  constexpr int RU = RT > 0 ? RT : 8;
#pragma unroll
  for (int r = 0; r < RU; ++r) {
    if (RT > 0 || r < R) {
      const float* q = q_smem + r * H;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < L::kChunks; ++i) {
        const int c0 = L::col(lane, i);
#pragma unroll
        for (int j4 = 0; j4 < L::kWidth; j4 += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(q + c0 + j4);
          if (TANH_MODE == 1) {
            // the softmax-equivalent score sum 2 alpha / (1 + pv q) (score2_eprod, common.cuh), one reciprocal per pair
            acc = score2_eprod(pv[i][j4], pv[i][j4 + 1], qv.x, qv.y, al[i][j4], al[i][j4 + 1], acc);
            acc = score2_eprod(pv[i][j4 + 2], pv[i][j4 + 3], qv.z, qv.w, al[i][j4 + 2], al[i][j4 + 3], acc);
          } else if (TANH_MODE == 3) {
            acc = score2_eprod_wide(pv[i][j4], pv[i][j4 + 1], qv.x, qv.y, al[i][j4], al[i][j4 + 1], acc);
            acc = score2_eprod_wide(pv[i][j4 + 2], pv[i][j4 + 3], qv.z, qv.w, al[i][j4 + 2], al[i][j4 + 3], acc);
          } else {
            float t[4];
            const float x[4] = {pv[i][j4] + qv.x, pv[i][j4 + 1] + qv.y, pv[i][j4 + 2] + qv.z, pv[i][j4 + 3] + qv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = TANH_MODE == 2 ? tanh_fast(x[j]) : tanhf(x[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = fmaf(al[i][j4 + j], t[j], acc);
          }
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) score_smem[r * n_items + l] = acc;
    }
  }
}

// One warp scores items l = warp, warp + n_warps, ...  The item's 512-wide row is staged gmem -> smem with
// cp.async (no registers held across the scoring of the previous row): a RING-slot ring per warp, RING - 1 rows in flight.
// Every lane reads back exactly the 16-byte chunks it copied itself, so no cross-lane barrier is needed.
// L2 prefetch of rows the kernel will read later holds no register and no shared memory, so it adds bytes in flight
// beyond what the ring and the register-held loads of the weighted sum keep outstanding: one instruction per 128-byte
// line, one warp-round beyond the ring for the scores, kSumAheadBytes ahead for the weighted sum. Distances measured in
// round 1 on fp32 rows (B = 1024, beam 3, attention ms per call): none 2.795 | (this) 2.588 | longer ones thrash L2
// (6 rounds / 64 rows: 3.11, 12 / 96: 3.65) — with ~590 images resident their lines evict each other from the 126 MB L2.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr int kRingBytes = 4096;          // per warp, smem-query variant: 2 fp32 rows or 4 16-bit rows (4 CTAs per SM)
constexpr int kRingBytesReg = 8192;       // per warp, register-query variant (2 CTAs per SM): 4 fp32 rows or 8 16-bit rows
constexpr int kSumAheadBytes = 32 * 1024;  // weighted_sum: bytes prefetched into L2 ahead of the loads
template <typename FeatT, int TANH_MODE, int RT>
__device__ __forceinline__ void score_rows(const FeatT* __restrict__ p_feat, int n_items, int R,
                                           const float* __restrict__ q_smem /*[R][H]*/, const float* __restrict__ alpha_smem,
                                           float* __restrict__ score_smem /*[R][n_items]*/, void* __restrict__ ring_raw,
                                           int warp, int lane, int n_warps) {
  using L = FeatLoad<FeatT>;
  constexpr int RING = kRingBytes / (H * (int)sizeof(FeatT));
  FeatT* ring = reinterpret_cast<FeatT*>(ring_raw);
  float al[L::kChunks][L::kWidth];  // this lane's slice of alpha stays in registers
#pragma unroll
  for (int i = 0; i < L::kChunks; ++i)
#pragma unroll
    for (int j = 0; j < L::kWidth; ++j)
      al[i][j] = ((TANH_MODE == 1 || TANH_MODE == 3) ? 2.0f : 1.0f) * alpha_smem[L::col(lane, i) + j];
  auto stage = [&](int l, int slot) {
    if (l < n_items) {
#pragma unroll
      for (int i = 0; i < L::kChunks; ++i) {
        const FeatT* src = p_feat + (long long)l * H + L::col(lane, i);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + slot * H + L::col(lane, i));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // (an empty group past the end keeps the group count uniform)
  };
  constexpr int kLines = H * (int)sizeof(FeatT) / 128;  // 128-byte lines per row (16 fp32, 8 fp16 / bf16)
  auto ahead = [&](int l) {
    if (l < n_items && lane < kLines) prefetch_l2(reinterpret_cast<const char*>(p_feat + (long long)l * H) + lane * 128);
  };
#pragma unroll
  for (int a = 0; a < RING - 1; ++a) stage(warp + a * n_warps, a);
  ahead(warp + (RING - 1) * n_warps);
  int it = 0;
  for (int l = warp; l < n_items; l += n_warps, ++it) {
    ahead(l + RING * n_warps);
    stage(l + (RING - 1) * n_warps, (it + RING - 1) % RING);
    asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");  // the group that holds row l has landed
    float pv[L::kChunks][L::kWidth];
#pragma unroll
    for (int i = 0; i < L::kChunks; ++i) L::load_shared(ring + (it % RING) * H, lane, i, pv[i]);
    score_one<FeatT, TANH_MODE, RT>(pv, al, l, n_items, R, q_smem, score_smem, lane);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Sum over the 32 lanes of EIGHT per-lane values at once: three exchange-halves stages (xor 16, 8, 4: a lane keeps the
// half of the values its lane bit selects and adds the other lane's copy of that half) leave ONE value per lane, two
// plain butterfly stages finish it. 9 shuffles for 8 sums instead of 40, and five dependent shuffle latencies per
// EIGHT rows instead of per row. Lanes with (lane & 3) == 0 end up holding the total of value
// k = 4 * bit4(lane) + 2 * bit3(lane) + bit2(lane).
__device__ __forceinline__ float treduce8(const float (&v)[8], int lane) {
  const unsigned full = 0xffffffffu;
  float w4[4], w2[2];
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) w4[i] = (b16 ? v[i + 4] : v[i]) + __shfl_xor_sync(full, b16 ? v[i] : v[i + 4], 16);
#pragma unroll
  for (int i = 0; i < 2; ++i) w2[i] = (b8 ? w4[i + 2] : w4[i]) + __shfl_xor_sync(full, b8 ? w4[i] : w4[i + 2], 8);
  float w1 = (b4 ? w2[1] : w2[0]) + __shfl_xor_sync(full, b4 ? w2[0] : w2[1], 4);
  w1 += __shfl_xor_sync(full, w1, 2);
  w1 += __shfl_xor_sync(full, w1, 1);
  return w1;
}

// Register-query variant of score_rows for RT = 1 or 3 rows per image (greedy / beam-3, the measured workloads).
// ncu on the smem-query kernel (B = 1024, beam 3, fp16 rows): l1tex at 79 % of peak — every row re-read the RT query
// vectors from shared memory, 6 KB per 1 KB row — and a third of all stall samples on the five dependent shuffle + add
// steps of the per-row warp reduction. Here a lane keeps ITS 16 columns of every query (and of alpha) in registers for
// the whole image (the lane <-> column mapping never changes), and the warp reductions of EIGHT consecutive rows are
// batched (treduce8): shared memory only carries the staged feature rows.
template <typename FeatT, int TANH_MODE, int RT>
__device__ __forceinline__ void score_rows_reg(const FeatT* __restrict__ p_feat, int n_items,
                                               const float* __restrict__ q_smem /*[RT][H]*/, const float* __restrict__ alpha_smem,
                                               float* __restrict__ score_smem /*[RT][n_items]*/, void* __restrict__ ring_raw,
                                               int warp, int lane, int n_warps, bool primed = false) {
  using L = FeatLoad<FeatT>;
  constexpr int NV = L::kChunks * L::kWidth;  // 16 values per lane per row
  constexpr int RING = kRingBytesReg / (H * (int)sizeof(FeatT));  // 4 or 8: divides the 8-row group
  static_assert(8 % RING == 0, "ring slots must divide the row group");
  FeatT* ring = reinterpret_cast<FeatT*>(ring_raw);
  float al[NV], q[RT][NV];
#pragma unroll
  for (int i = 0; i < L::kChunks; ++i)
#pragma unroll
    for (int j = 0; j < L::kWidth; j += 4) {
      const float4 a = *reinterpret_cast<const float4*>(alpha_smem + L::col(lane, i) + j);
      const float sc = (TANH_MODE == 1 || TANH_MODE == 3) ? 2.0f : 1.0f;
      al[i * L::kWidth + j] = sc * a.x; al[i * L::kWidth + j + 1] = sc * a.y;
      al[i * L::kWidth + j + 2] = sc * a.z; al[i * L::kWidth + j + 3] = sc * a.w;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float4 t = *reinterpret_cast<const float4*>(q_smem + r * H + L::col(lane, i) + j);
        q[r][i * L::kWidth + j] = t.x; q[r][i * L::kWidth + j + 1] = t.y;
        q[r][i * L::kWidth + j + 2] = t.z; q[r][i * L::kWidth + j + 3] = t.w;
      }
    }
  auto stage = [&](int l, int slot) {
    if (l < n_items) {
#pragma unroll
      for (int i = 0; i < L::kChunks; ++i) {
        const FeatT* src = p_feat + (long long)l * H + L::col(lane, i);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + slot * H + L::col(lane, i));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  constexpr int kLines = H * (int)sizeof(FeatT) / 128;
  auto ahead = [&](int l) {
    if (l < n_items && lane < kLines) prefetch_l2(reinterpret_cast<const char*>(p_feat + (long long)l * H) + lane * 128);
  };
  if (!primed) {
#pragma unroll
    for (int a = 0; a < RING - 1; ++a) stage(warp + a * n_warps, a);
    ahead(warp + (RING - 1) * n_warps);
  }
  for (int l0 = warp; l0 < n_items; l0 += 8 * n_warps) {  // 8 rows of this warp per group
    float acc[RT][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int l = l0 + k * n_warps;
      ahead(l + RING * n_warps);
      stage(l + (RING - 1) * n_warps, (k + RING - 1) % RING);
      asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r][k] = 0.f;
      if (l < n_items && TANH_MODE == 1) {  // warp-uniform; packed fp32 (score4_eprod)
        float2 pv2[NV / 2];
#pragma unroll
        for (int i = 0; i < L::kChunks; ++i) L::load_shared2(ring + (k % RING) * H, lane, i, pv2 + i * (L::kWidth / 2));
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < NV; j += 4)
            a2 = score4_eprod(pv2[j / 2], pv2[j / 2 + 1], make_float2(q[r][j], q[r][j + 1]), make_float2(q[r][j + 2], q[r][j + 3]),
                              make_float2(al[j], al[j + 1]), make_float2(al[j + 2], al[j + 3]), a2);
          acc[r][k] = a2.x + a2.y;
        }
      } else if (l < n_items) {
        float pv[NV];
#pragma unroll
        for (int i = 0; i < L::kChunks; ++i) L::load_shared(ring + (k % RING) * H, lane, i, pv + i * L::kWidth);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          float a = 0.f;
#pragma unroll
          for (int j = 0; j < NV; j += 2) {
            if (TANH_MODE == 1) a = score2_eprod(pv[j], pv[j + 1], q[r][j], q[r][j + 1], al[j], al[j + 1], a);
            else if (TANH_MODE == 3) a = score2_eprod_wide(pv[j], pv[j + 1], q[r][j], q[r][j + 1], al[j], al[j + 1], a);
            else if (TANH_MODE == 2) a = fmaf(al[j], tanh_fast(pv[j] + q[r][j]), fmaf(al[j + 1], tanh_fast(pv[j + 1] + q[r][j + 1]), a));
            else a = fmaf(al[j], tanhf(pv[j] + q[r][j]), fmaf(al[j + 1], tanhf(pv[j + 1] + q[r][j + 1]), a));
          }
          acc[r][k] = a;
        }
      }
    }
    const int kk = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const int lw = l0 + kk * n_warps;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float tot = treduce8(acc[r], lane);
      if ((lane & 3) == 0 && lw < n_items) score_smem[r * n_items + lw] = tot;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// The first RING - 1 rows of a warp's stream (rows warp, warp + n_warps, ...) requested into its ring, plus the L2
// prefetch of the next one: what score_rows_reg / wsum_rows_reg do before their loops, callable EARLY — before the queries
// are prepared, or right after the scoring pass — so that HBM works through the kernel's set-up and its softmax.
template <typename FeatT>
__device__ __forceinline__ void ring_prime(const FeatT* __restrict__ feat, int n_items, void* __restrict__ ring_raw, int warp,
                                           int lane, int n_warps) {
  using L = FeatLoad<FeatT>;
  constexpr int RING = kRingBytesReg / (H * (int)sizeof(FeatT));
  FeatT* ring = reinterpret_cast<FeatT*>(ring_raw);
#pragma unroll
  for (int a = 0; a < RING - 1; ++a) {
    const int l = warp + a * n_warps;
    if (l < n_items) {
#pragma unroll
      for (int i = 0; i < L::kChunks; ++i) {
        const FeatT* src = feat + (long long)l * H + L::col(lane, i);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + a * H + L::col(lane, i));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  constexpr int kLines = H * (int)sizeof(FeatT) / 128;
  const int la = warp + (RING - 1) * n_warps;
  if (la < n_items && lane < kLines) prefetch_l2(reinterpret_cast<const char*>(feat + (long long)la * H) + lane * 128);
}

// Weighted sum with the scoring pass's structure: warp w streams rows w, w + 8, ... through its cp.async ring (RING - 1
// rows in flight, nothing held in registers), a lane accumulates ITS 16 columns of the RT context vectors, and the eight
// per-warp partial sums meet in shared memory at the end (each warp parks its partial in its own ring area). The
// register-held global loads of weighted_sum (32 KB in flight per SM at two CTAs) sat on the long scoreboard for 30 % of
// the kernel (ncu, B = 1024, beam 3); this keeps 7 KB per warp in flight like the scores.
// Expects ring_prime<FeatT>() to have been called for `feat`. Leaves this warp's partial [RT][H] floats at ring_raw.
template <typename FeatT, int RT>
__device__ __forceinline__ void wsum_rows_reg(const FeatT* __restrict__ feat, int n_items, const float* __restrict__ w_smem,
                                              void* __restrict__ ring_raw, int warp, int lane, int n_warps) {
  using L = FeatLoad<FeatT>;
  constexpr int NV = L::kChunks * L::kWidth;
  constexpr int RING = kRingBytesReg / (H * (int)sizeof(FeatT));
  FeatT* ring = reinterpret_cast<FeatT*>(ring_raw);
  float2 ctx[RT][NV / 2];  // register pairs: the accumulation runs on packed fp32 FMAs (fma2)
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) ctx[r][j] = make_float2(0.f, 0.f);
  constexpr int kLines = H * (int)sizeof(FeatT) / 128;
  int it = 0;
  for (int l = warp; l < n_items; l += n_warps, ++it) {
    const int la = l + RING * n_warps, ls = l + (RING - 1) * n_warps;
    if (la < n_items && lane < kLines) prefetch_l2(reinterpret_cast<const char*>(feat + (long long)la * H) + lane * 128);
    if (ls < n_items) {
      const int slot = (it + RING - 1) % RING;
#pragma unroll
      for (int i = 0; i < L::kChunks; ++i) {
        const FeatT* src = feat + (long long)ls * H + L::col(lane, i);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + slot * H + L::col(lane, i));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
    float2 pv[NV / 2];
#pragma unroll
    for (int i = 0; i < L::kChunks; ++i) L::load_shared2(ring + (it % RING) * H, lane, i, pv + i * (L::kWidth / 2));
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float w = w_smem[r * n_items + l];  // broadcast
      const float2 w2 = make_float2(w, w);
#pragma unroll
      for (int j = 0; j < NV / 2; ++j) ctx[r][j] = fma2(w2, pv[j], ctx[r][j]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();  // every lane has read its last staged row: the ring area now holds this warp's partial sums
  float* part = reinterpret_cast<float*>(ring_raw);
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int i = 0; i < L::kChunks; ++i)
#pragma unroll
      for (int j = 0; j < L::kWidth; j += 4) {
        const float2 lo = ctx[r][(i * L::kWidth + j) / 2], hi = ctx[r][(i * L::kWidth + j) / 2 + 1];
        *reinterpret_cast<float4*>(part + r * H + L::col(lane, i) + j) = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
}

// FAST: exp through the raw MUFU op and one reciprocal per row (weights differ from expf / true division by ~3e-7
// relative; a row's softmax was three serial warps of ~30-instruction expf + division chains while five warps waited).
template <bool FAST = false>
__device__ __forceinline__ void softmax_rows(float* score_smem, int n_items, int R, int warp, int lane, int n_warps,
                                             float* w_out, long long ld_w, long long row0) {
  for (int r = warp; r < R; r += n_warps) {
    float* s = score_smem + r * n_items;
    float mx = -INFINITY;
    for (int l = lane; l < n_items; l += 32) mx = fmaxf(mx, s[l]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int l = lane; l < n_items; l += 32) {
      float e;
      if (FAST) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((s[l] - mx) * 1.4426950408889634f));
      else e = expf(s[l] - mx);
      s[l] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int l = lane; l < n_items; l += 32) {
      float w = FAST ? s[l] * inv : s[l] / sum;
      s[l] = w;
      if (w_out) w_out[(row0 + r) * ld_w + l] = w;
    }
  }
}

template <typename FeatT>
__device__ __forceinline__ float2 load2(const FeatT* p);
template <>
__device__ __forceinline__ float2 load2<float>(const float* p) {
  return __ldg(reinterpret_cast<const float2*>(p));
}
template <>
__device__ __forceinline__ float2 load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  unsigned int u = __ldg(reinterpret_cast<const unsigned int*>(p));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
template <>
__device__ __forceinline__ float2 load2<__half>(const __half* p) {
  unsigned int u = __ldg(reinterpret_cast<const unsigned int*>(p));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// context[r] = sum_l w[r][l] * feat[l] ; thread owns columns (2t, 2t+1); weights are read from smem four
// items at a time (128-bit broadcast loads) when the row length allows it. UNR rows are in flight per thread: 8 fp32
// rows (8 x 8 B) or 16 rows of a 16-bit type (16 x 4 B) — the kernel is bound by bytes in flight, so halving the
// element width must not halve them.
template <typename FeatT, int RT>
__device__ __forceinline__ void weighted_sum(const FeatT* __restrict__ feat, int n_items, int R,
                                             const float* __restrict__ w_smem, RowDest dst, long long row0, int dst_col) {
  constexpr int RU = RT > 0 ? RT : 8;
  constexpr int UNR = sizeof(FeatT) == 4 ? 8 : 16;
  constexpr int kAhead = kSumAheadBytes / (H * (int)sizeof(FeatT));
  const int c = threadIdx.x * 2;
  float2 acc[RU];
#pragma unroll
  for (int r = 0; r < RU; ++r) acc[r] = make_float2(0.f, 0.f);
  int l = 0;
  if ((n_items & 3) == 0 && (reinterpret_cast<uintptr_t>(w_smem) & 15) == 0) {
    for (; l + UNR <= n_items; l += UNR) {
      if ((threadIdx.x & (64 / (int)sizeof(FeatT) - 1)) == 0) {  // one thread per 128-byte line of the UNR rows kAhead ahead
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (l + kAhead + u < n_items) prefetch_l2(feat + (long long)(l + kAhead + u) * H + c);
      }
      float2 a[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) a[u] = load2<FeatT>(feat + (long long)(l + u) * H + c);
#pragma unroll
      for (int r = 0; r < RU; ++r)
        if (RT > 0 || r < R) {
#pragma unroll
          for (int hh = 0; hh < UNR / 4; ++hh) {
            const float4 w = *reinterpret_cast<const float4*>(w_smem + r * n_items + l + 4 * hh);
            acc[r].x = fmaf(w.x, a[4 * hh].x, acc[r].x); acc[r].y = fmaf(w.x, a[4 * hh].y, acc[r].y);
            acc[r].x = fmaf(w.y, a[4 * hh + 1].x, acc[r].x); acc[r].y = fmaf(w.y, a[4 * hh + 1].y, acc[r].y);
            acc[r].x = fmaf(w.z, a[4 * hh + 2].x, acc[r].x); acc[r].y = fmaf(w.z, a[4 * hh + 2].y, acc[r].y);
            acc[r].x = fmaf(w.w, a[4 * hh + 3].x, acc[r].x); acc[r].y = fmaf(w.w, a[4 * hh + 3].y, acc[r].y);
          }
        }
    }
    for (; l + 4 <= n_items; l += 4) {
      float2 a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = load2<FeatT>(feat + (long long)(l + u) * H + c);
#pragma unroll
      for (int r = 0; r < RU; ++r)
        if (RT > 0 || r < R) {
          const float4 w = *reinterpret_cast<const float4*>(w_smem + r * n_items + l);
          acc[r].x = fmaf(w.x, a[0].x, acc[r].x); acc[r].y = fmaf(w.x, a[0].y, acc[r].y);
          acc[r].x = fmaf(w.y, a[1].x, acc[r].x); acc[r].y = fmaf(w.y, a[1].y, acc[r].y);
          acc[r].x = fmaf(w.z, a[2].x, acc[r].x); acc[r].y = fmaf(w.z, a[2].y, acc[r].y);
          acc[r].x = fmaf(w.w, a[3].x, acc[r].x); acc[r].y = fmaf(w.w, a[3].y, acc[r].y);
        }
    }
  }
  for (; l < n_items; ++l) {
    float2 a = load2<FeatT>(feat + (long long)l * H + c);
#pragma unroll
    for (int r = 0; r < RU; ++r)
      if (RT > 0 || r < R) {
        float w = w_smem[r * n_items + l];
        acc[r].x = fmaf(w, a.x, acc[r].x);
        acc[r].y = fmaf(w, a.y, acc[r].y);
      }
  }
#pragma unroll
  for (int r = 0; r < RU; ++r)
    if (RT > 0 || r < R) dst.store2(row0 + r, dst_col + c, acc[r]);
}

template <typename FeatT>
__device__ __forceinline__ void prefetch_first_rows(const FeatT* a0, int L) {
  constexpr int kLines = H * (int)sizeof(FeatT) / 128;
  constexpr int kAhead = kSumAheadBytes / (H * (int)sizeof(FeatT));
  for (int i = threadIdx.x; i < kAhead * kLines && i / kLines < L; i += 256)
    prefetch_l2(reinterpret_cast<const char*>(a0 + (long long)(i / kLines) * H) + (i % kLines) * 128);
}

// PREC = ISC_PREC_*: the representation of the full-width features and the scoring mode of the wide path:
//   FP32: fp32 ReLU(.) + tanhf | BF16X3: fp32 att, fp32 exp(-2 p), e-product (mode 3) | BF16: bf16 att and p, tanh.approx.
// The fast path (p.p_att16 set and the image not flagged; tensor-core precisions): fp16 p_att16 scored in mode 1, att16
// (or, in ISC_PREC_BF16, the bf16 att) summed.
constexpr int kSentiDirectMax = 16;                               // sentiment rows held in registers by the direct path

// sentiment-word scores by direct global loads (S ~ 11 fp32 rows per image: no staging); queries from shared memory
template <int MODE, int RT>
__device__ __forceinline__ void score_senti_direct(const float* __restrict__ p_sw, int S, const float* __restrict__ q_smem,
                                                   const float* __restrict__ alpha_smem, float* __restrict__ score_smem, int warp,
                                                   int lane) {
  for (int l = warp; l < S; l += 8) {
    float pv[4][4], al[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p_sw + (long long)l * H + i * 128 + lane * 4));
      pv[i][0] = t.x; pv[i][1] = t.y; pv[i][2] = t.z; pv[i][3] = t.w;
      const float4 a = *reinterpret_cast<const float4*>(alpha_smem + i * 128 + lane * 4);
      const float sc = (MODE == 1 || MODE == 3) ? 2.0f : 1.0f;
      al[i][0] = sc * a.x; al[i][1] = sc * a.y; al[i][2] = sc * a.z; al[i][3] = sc * a.w;
    }
    score_one<float, MODE, RT>(pv, al, l, S, RT, q_smem, score_smem, lane);
  }
}

template <int RT>
struct AttnVariant {
  static constexpr bool kRegQuery = RT == 1 || RT == 3;  // queries in registers (2 CTAs per SM), else in shared memory (4)
  static constexpr int kRing = kRegQuery ? kRingBytesReg : kRingBytes;
};
template <typename FeatT, int TANH_MODE, int RT>
__device__ __forceinline__ void score_dispatch(const FeatT* __restrict__ p_feat, int n_items, int R, const float* q_smem,
                                               const float* alpha_smem, float* score_smem, void* ring, int warp, int lane) {
  if (AttnVariant<RT>::kRegQuery) score_rows_reg<FeatT, TANH_MODE, (RT > 0 ? RT : 1)>(p_feat, n_items, q_smem, alpha_smem, score_smem, ring, warp, lane, 8);
  else score_rows<FeatT, TANH_MODE, RT>(p_feat, n_items, R, q_smem, alpha_smem, score_smem, ring, warp, lane, 8);
}

template <int PREC, int RT>
__global__ void __launch_bounds__(256, (RT == 1 || RT == 3) ? 2 : 3) attention_kernel(AttnParams p) {
  extern __shared__ __align__(16) float sm[];
  pdl_trigger();
  pdl_wait();
  constexpr int WIDE_MODE = PREC == ISC_PREC_FP32 ? 0 : (PREC == ISC_PREC_BF16X3 ? 3 : 2);
  constexpr bool kReg = AttnVariant<RT>::kRegQuery;
  typedef typename std::conditional<PREC == ISC_PREC_BF16, __nv_bfloat16, float>::type WideT;
  const int R = RT > 0 ? RT : p.R, L = p.L, S = p.S;
  const int Lp = (L + 3) & ~3, Sp = (S + 3) & ~3;  // padded score rows keep 16-byte alignment
  float* q_c = sm;                 // [R][H] content query  h2att(h)
  float* q_s = q_c + R * H;        // [R][H] senti query    h2word(h) + label2word(sl)
  float* alpha_c = q_s + R * H;    // [H]
  float* alpha_s = alpha_c + H;    // [H]
  float* sc_c = alpha_s + H;       // [R][L]
  float* sc_s = sc_c + R * Lp;     // [R][S]
  uint8_t* ring_base = reinterpret_cast<uint8_t*>(sc_s + R * Sp);  // [8 warps][ring bytes of this variant]
  const int img = blockIdx.x;
  const long long row0 = (long long)img * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool fast = PREC != ISC_PREC_FP32 && p.p_att16 != nullptr && (p.flags == nullptr || p.flags[img] == 0);
  const bool att_half = fast && p.att16 != nullptr;
  const long long foff = (long long)img * L * H;
  void* ring = ring_base + warp * AttnVariant<RT>::kRing;
  const __half* p16 = reinterpret_cast<const __half*>(p.p_att16) + foff;
  const __half* a16 = reinterpret_cast<const __half*>(p.att16) + foff;
  const WideT* pw = reinterpret_cast<const WideT*>(p.p_att) + foff;
  const WideT* aw = reinterpret_cast<const WideT*>(p.att) + foff;
  if (kReg && p.sw) {  // the image's 2 x S sentiment-word rows (fp32): on their way to L2 before anything needs them
    const int lines = S * H * 4 / 128;
    for (int i = threadIdx.x; i < 2 * lines; i += 256) {
      const float* base = (i < lines ? p.p_sw : p.sw) + (long long)img * S * H;
      prefetch_l2(reinterpret_cast<const char*>(base) + (i < lines ? i : i - lines) * 128);
    }
  }
  if (kReg && p.att) {  // the projected rows start streaming before the queries exist
    if (fast) ring_prime<__half>(p16, L, ring, warp, lane, 8);
    else ring_prime<WideT>(pw, L, ring, warp, lane, 8);
  }
  // e-product modes keep the queries (like the projected features) as exp(-2 x); the fast path folds its feature scale
  // 2^15 into the query
  auto to_qc = [&](float qc) { return fast ? exp_neg2(qc) * (1.0f / kFastScale) : (WIDE_MODE == 3 ? exp_neg2_wide(qc) : qc); };
  auto to_qs = [&](float qw) { return WIDE_MODE == 3 ? exp_neg2_wide(qw) : qw; };
  if (kReg) {
    // every global load of the set-up is requested before the first one is used (ncu: with the loop rolled, its six
    // dependent load -> exp -> store rounds per thread were 21 % of the kernel, HBM idle)
    constexpr int NQ = (RT > 0 ? RT : 1) * H / 256;
    float qc[NQ], qw[NQ], pwd[NQ], ac[2], as_[2];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int i = threadIdx.x + k * 256, r = i / H, c = i - r * H;
      const float* hp = p.hproj + (row0 + r) * p.ld_hproj;
      qc[k] = __ldg(hp + c);
      qw[k] = __ldg(hp + H + c);
      pwd[k] = p.pre_word ? __ldg(p.pre_word + (long long)img * H + c) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      ac[k] = __ldg(p.alpha_c + threadIdx.x + k * 256);
      as_[k] = __ldg(p.alpha_s + threadIdx.x + k * 256);
    }
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      q_c[threadIdx.x + k * 256] = to_qc(qc[k]);
      q_s[threadIdx.x + k * 256] = to_qs(qw[k] + pwd[k]);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      alpha_c[threadIdx.x + k * 256] = ac[k];
      alpha_s[threadIdx.x + k * 256] = as_[k];
    }
  } else {
    for (int i = threadIdx.x; i < R * H; i += 256) {
      int r = i / H, c = i - r * H;
      const float* hp = p.hproj + (row0 + r) * p.ld_hproj;
      q_c[i] = to_qc(hp[c]);
      q_s[i] = to_qs(hp[H + c] + (p.pre_word ? p.pre_word[(long long)img * H + c] : 0.f));
    }
    for (int i = threadIdx.x; i < H; i += 256) {
      alpha_c[i] = p.alpha_c[i];
      alpha_s[i] = p.alpha_s[i];
    }
  }
  __syncthreads();
  if (kReg) {
    // ---- register-query variant: one stream per warp through both passes
    constexpr int RQ = RT > 0 ? RT : 1;
    if (p.att) {
      if (fast) score_rows_reg<__half, 1, RQ>(p16, L, q_c, alpha_c, sc_c, ring, warp, lane, 8, true);
      else score_rows_reg<WideT, WIDE_MODE, RQ>(pw, L, q_c, alpha_c, sc_c, ring, warp, lane, 8, true);
      // this warp's feature rows start streaming now: they travel through the sentiment scores and the softmax
      if (att_half) ring_prime<__half>(a16, L, ring, warp, lane, 8);
      else ring_prime<WideT>(aw, L, ring, warp, lane, 8);
    }
    if (p.sw) score_senti_direct<WIDE_MODE, RQ>(p.p_sw + (long long)img * S * H, S, q_s, alpha_s, sc_s, warp, lane);
    __syncthreads();
    // the content rows' softmax on warps 0..R-1, the sentiment rows' on warps 7, 6, ...: both at once. (Measured and
    // dropped: no softmax pass, every warp deriving (max, 1 / sum) itself and forming the weights of its rows inside the
    // weighted-sum loop — one barrier and the idle warps less, but 15 us MORE per launch: the ex2 -> multiply in front of
    // each row's 48 FMAs lengthens the loop's dependent chain.)
    if (p.att) softmax_rows<PREC != ISC_PREC_FP32>(sc_c, L, R, warp, lane, 8, p.cont_w, p.ld_cont_w, row0);
    if (p.sw) softmax_rows<PREC != ISC_PREC_FP32>(sc_s, S, R, 7 - warp, lane, 8, p.senti_w, p.ld_senti_w, row0);
    // the sentiment rows of this thread's two columns: requested now, used after the content pass
    const int c2 = threadIdx.x * 2;
    float2 sv[kSentiDirectMax];
    const bool senti_direct = p.sw != nullptr && S <= kSentiDirectMax;
    if (senti_direct) {
#pragma unroll
      for (int l = 0; l < kSentiDirectMax; ++l)
        sv[l] = l < S ? __ldg(reinterpret_cast<const float2*>(p.sw + ((long long)img * S + l) * H + c2)) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (p.att) {
      if (att_half) wsum_rows_reg<__half, RQ>(a16, L, sc_c, ring, warp, lane, 8);
      else wsum_rows_reg<WideT, RQ>(aw, L, sc_c, ring, warp, lane, 8);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < RQ; ++r) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float2 t = *reinterpret_cast<const float2*>(ring_base + w * AttnVariant<RT>::kRing + (r * H + c2) * sizeof(float));
          acc.x += t.x;
          acc.y += t.y;
        }
        p.cont_dst.store2(row0 + r, p.cont_col + c2, acc);
      }
    }
    if (senti_direct) {
#pragma unroll
      for (int r = 0; r < RQ; ++r) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = 0; l < kSentiDirectMax; ++l)
          if (l < S) {
            const float w = sc_s[r * S + l];
            acc.x = fmaf(w, sv[l].x, acc.x);
            acc.y = fmaf(w, sv[l].y, acc.y);
          }
        p.senti_dst.store2(row0 + r, p.senti_col + c2, acc);
      }
    } else if (p.sw) {
      weighted_sum<float, RT>(p.sw + (long long)img * S * H, S, R, sc_s, p.senti_dst, row0, p.senti_col);
    }
    return;
  }
  // ---- shared-memory-query variant (other beam sizes)
  if (p.att) {
    if (fast) score_rows<__half, 1, RT>(p16, L, R, q_c, alpha_c, sc_c, ring, warp, lane, 8);
    else score_rows<WideT, WIDE_MODE, RT>(pw, L, R, q_c, alpha_c, sc_c, ring, warp, lane, 8);
  }
  if (p.sw) score_rows<float, WIDE_MODE, RT>(p.p_sw + (long long)img * S * H, S, R, q_s, alpha_s, sc_s, ring, warp, lane, 8);
  __syncthreads();
  if (p.att) {  // the first rows of the weighted sum travel to L2 while the softmax runs
    if (att_half) prefetch_first_rows(a16, L);
    else prefetch_first_rows(aw, L);
  }
  if (p.att) softmax_rows(sc_c, L, R, warp, lane, 8, p.cont_w, p.ld_cont_w, row0);
  if (p.sw) softmax_rows(sc_s, S, R, warp, lane, 8, p.senti_w, p.ld_senti_w, row0);
  __syncthreads();
  if (p.att) {
    if (att_half) weighted_sum<__half, RT>(a16, L, R, sc_c, p.cont_dst, row0, p.cont_col);
    else weighted_sum<WideT, RT>(aw, L, R, sc_c, p.cont_dst, row0, p.cont_col);
  }
  if (p.sw) weighted_sum<float, RT>(p.sw + (long long)img * S * H, S, R, sc_s, p.senti_dst, row0, p.senti_col);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant (the 16-bit path, RT = 1 or 3 rows per image): the image's projected rows and then its feature rows
// — two contiguous 196 KB blocks — are streamed through ONE 64 KB shared-memory ring of 8 KB chunks (8 rows) by bulk
// asynchronous copies (cp.async.bulk ... mbarrier::complete_tx), issued by a single elected thread: no per-lane copy
// instructions, no address arithmetic, no registers held, 56 KB in flight per CTA from the first instruction of the
// kernel on — the stream starts BEFORE the queries are prepared and runs through the softmax between the two passes.
// ncu on the per-warp cp.async version (B = 1024, beam 3): 16 % of the kernel's time went to the query set-up with HBM
// idle, 39 % to the weighted sums, whose register-held global loads (32 KB in flight per SM) stalled on the long
// scoreboard. Consumers: scoring, warp w takes row 8c + w of chunk c (register-resident queries, treduce8 over the
// 8 chunks of a group); weighted sum, thread t owns columns 2t, 2t + 1 of all 8 rows of a chunk. full[slot] / empty[slot]
// mbarriers; the producer thread (warp 0, lane 0) tops the ring up opportunistically (try_wait) and blocks only for the
// chunk its own warp is about to consume. The sentiment words (11 fp32 rows) are read directly.
// ---------------------------------------------------------------------------------------------------------------
namespace bulk {
__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(s32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {  // bounded: a protocol bug traps, never hangs
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("isc attention: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(s32(bar))
               : "memory");
}
}  // namespace bulk

constexpr int kTmaSlots = 8;
constexpr int kTmaChunkRows = 8;
constexpr int kTmaRowBytes = H * 2;                               // 16-bit rows
constexpr int kTmaChunkBytes = kTmaChunkRows * kTmaRowBytes;      // 8 KB
constexpr int kTmaRingBytes = kTmaSlots * kTmaChunkBytes;         // 64 KB

template <int PREC, int RT>
__global__ void __launch_bounds__(256, 2) attention_tma_kernel(AttnParams p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  static_assert(RT == 1 || RT == 3, "register-resident queries: 1 or 3 rows per image");
  constexpr int WIDE_MODE = PREC == ISC_PREC_BF16X3 ? 3 : 2;
  typedef typename std::conditional<PREC == ISC_PREC_BF16, __nv_bfloat16, float>::type WideT;
  typedef typename std::conditional<PREC == ISC_PREC_BF16, __nv_bfloat16, __half>::type AttT;  // 16-bit feature rows
  constexpr int R = RT;
  const int L = p.L, S = p.S;
  const int Lp = (L + 3) & ~3, Sp = (S + 3) & ~3;
  uint8_t* ring = smraw;                                      // [kTmaSlots][8 KB]; the wide path's per-warp rings otherwise
  float* q_c = reinterpret_cast<float*>(smraw + kTmaRingBytes);  // [R][H]
  float* q_s = q_c + R * H;
  float* alpha_c = q_s + R * H;
  float* alpha_s = alpha_c + H;
  float* sc_c = alpha_s + H;       // [R][L]
  float* sc_s = sc_c + R * Lp;     // [R][S]
  uint64_t* full = reinterpret_cast<uint64_t*>(sc_s + R * Sp + ((R * Sp) & 1));  // 8-byte aligned
  uint64_t* empty = full + kTmaSlots;
  const int img = blockIdx.x;
  const long long row0 = (long long)img * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTmaSlots; ++i) {
      bulk::mbar_init(&full[i], 1);
      bulk::mbar_init(&empty[i], 8);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  const bool fast = p.flags == nullptr || p.flags[img] == 0;
  const long long foff = (long long)img * L * H;
  const uint8_t* src_p = reinterpret_cast<const uint8_t*>(p.p_att16) + foff * 2;
  const uint8_t* src_a = (PREC == ISC_PREC_BF16 ? reinterpret_cast<const uint8_t*>(p.att) : reinterpret_cast<const uint8_t*>(p.att16)) + foff * 2;
  const int n_ca = (L + kTmaChunkRows - 1) / kTmaChunkRows;  // chunks per tensor
  const int n_chunks = (fast && p.att) ? 2 * n_ca : 0;
  int next_issue = 0;  // producer state (warp 0, lane 0)
  // issue every chunk up to `need` (blocking on a free slot) and, while slots are free, up to `want`
  auto pump = [&](int need, int want) {
    while (next_issue < n_chunks && next_issue <= want) {
      const int c = next_issue, slot = c % kTmaSlots, round = c / kTmaSlots;
      if (round > 0) {
        const uint32_t par = (round - 1) & 1;
        if (c <= need) bulk::mbar_wait(&empty[slot], par);
        else if (!bulk::mbar_try_wait(&empty[slot], par)) break;
      }
      const int cc = c < n_ca ? c : c - n_ca;
      const int rows = min(kTmaChunkRows, L - cc * kTmaChunkRows);
      const uint8_t* src = (c < n_ca ? src_p : src_a) + (long long)cc * kTmaChunkBytes;
      bulk::mbar_expect_tx(&full[slot], rows * kTmaRowBytes);
      bulk::copy_g2s(ring + slot * kTmaChunkBytes, src, rows * kTmaRowBytes, &full[slot]);
      ++next_issue;
    }
  };
  const bool producer = warp == 0 && lane == 0;
  if (producer) pump(-1, kTmaSlots - 1);  // the stream starts before the queries exist

  for (int i = threadIdx.x; i < R * H; i += 256) {
    int r = i / H, c = i - r * H;
    const float* hp = p.hproj + (row0 + r) * p.ld_hproj;
    const float qc = hp[c];
    const float qw = hp[H + c] + (p.pre_word ? p.pre_word[(long long)img * H + c] : 0.f);
    q_c[i] = fast ? exp_neg2(qc) * (1.0f / kFastScale) : (WIDE_MODE == 3 ? exp_neg2_wide(qc) : qc);
    q_s[i] = WIDE_MODE == 3 ? exp_neg2_wide(qw) : qw;
  }
  for (int i = threadIdx.x; i < H; i += 256) {
    alpha_c[i] = p.alpha_c[i];
    alpha_s[i] = p.alpha_s[i];
  }
  __syncthreads();
  if (p.sw) score_senti_direct<WIDE_MODE, RT>(p.p_sw + (long long)img * S * H, S, q_s, alpha_s, sc_s, warp, lane);
  if (p.att && !fast) {
    // flagged image: full-width rows through per-warp cp.async rings (the ring memory is free: nothing was streamed)
    score_dispatch<WideT, WIDE_MODE, RT>(reinterpret_cast<const WideT*>(p.p_att) + foff, L, R, q_c, alpha_c, sc_c,
                                         ring + warp * kRingBytesReg, warp, lane);
  } else if (p.att) {
    using LD = FeatLoad<__half>;
    constexpr int NV = 16;
    float al[NV], q[RT][NV];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 a = *reinterpret_cast<const float4*>(alpha_c + LD::col(lane, i) + j);
        al[i * 8 + j] = 2.0f * a.x; al[i * 8 + j + 1] = 2.0f * a.y; al[i * 8 + j + 2] = 2.0f * a.z; al[i * 8 + j + 3] = 2.0f * a.w;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const float4 t = *reinterpret_cast<const float4*>(q_c + r * H + LD::col(lane, i) + j);
          q[r][i * 8 + j] = t.x; q[r][i * 8 + j + 1] = t.y; q[r][i * 8 + j + 2] = t.z; q[r][i * 8 + j + 3] = t.w;
        }
      }
    for (int g = 0; g * 8 < n_ca; ++g) {  // 8 chunks per group: slot = k, parity = g & 1
      float acc[RT][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = g * 8 + k;
#pragma unroll
        for (int r = 0; r < RT; ++r) acc[r][k] = 0.f;
        if (c < n_ca) {  // block-uniform
          if (producer) pump(c, c + kTmaSlots - 1);
          __syncwarp();
          bulk::mbar_wait(&full[k], g & 1);
          const int l = c * kTmaChunkRows + warp;
          if (l < L) {
            float pv[NV];
            const __half* row = reinterpret_cast<const __half*>(ring + k * kTmaChunkBytes + warp * kTmaRowBytes);
            LD::load_shared(row, lane, 0, pv);
            LD::load_shared(row, lane, 1, pv + 8);
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              float a = 0.f;
#pragma unroll
              for (int j = 0; j < NV; j += 2) a = score2_eprod(pv[j], pv[j + 1], q[r][j], q[r][j + 1], al[j], al[j + 1], a);
              acc[r][k] = a;
            }
          }
          __syncwarp();
          if (lane == 0) bulk::mbar_arrive(&empty[k]);
        }
      }
      const int kk = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      const int lw = (g * 8 + kk) * kTmaChunkRows + warp;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float tot = treduce8(acc[r], lane);
        if ((lane & 3) == 0 && lw < L) sc_c[r * L + lw] = tot;
      }
    }
  }
  __syncthreads();
  if (p.att) softmax_rows(sc_c, L, R, warp, lane, 8, p.cont_w, p.ld_cont_w, row0);
  if (p.sw) softmax_rows(sc_s, S, R, warp, lane, 8, p.senti_w, p.ld_senti_w, row0);
  // the sentiment rows of this thread's two columns: requested now, used after the content pass (latency hidden)
  const int c2 = threadIdx.x * 2;
  float2 sv[kSentiDirectMax];
  const bool senti_direct = p.sw != nullptr && S <= kSentiDirectMax;
  if (senti_direct) {
#pragma unroll
    for (int l = 0; l < kSentiDirectMax; ++l)
      sv[l] = l < S ? __ldg(reinterpret_cast<const float2*>(p.sw + ((long long)img * S + l) * H + c2)) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  if (p.att && !fast) {
    weighted_sum<WideT, RT>(reinterpret_cast<const WideT*>(p.att) + foff, L, R, sc_c, p.cont_dst, row0, p.cont_col);
  } else if (p.att) {
    float2 acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = make_float2(0.f, 0.f);
    for (int cc = 0; cc < n_ca; ++cc) {
      const int c = n_ca + cc, slot = c % kTmaSlots;
      if (producer) pump(c, c + kTmaSlots - 1);
      __syncwarp();
      bulk::mbar_wait(&full[slot], (c / kTmaSlots) & 1);
      const AttT* chunk = reinterpret_cast<const AttT*>(ring + slot * kTmaChunkBytes) + c2;
      const int l0 = cc * kTmaChunkRows;
      float2 a[kTmaChunkRows];
      if (l0 + kTmaChunkRows <= L) {
#pragma unroll
        for (int u = 0; u < kTmaChunkRows; ++u) {
          const unsigned v = *reinterpret_cast<const unsigned*>(chunk + u * H);
          a[u] = PREC == ISC_PREC_BF16 ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v))
                                       : __half22float2(*reinterpret_cast<const __half2*>(&v));
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float4 w = *reinterpret_cast<const float4*>(sc_c + r * L + l0 + 4 * hh);
            acc[r].x = fmaf(w.x, a[4 * hh].x, acc[r].x); acc[r].y = fmaf(w.x, a[4 * hh].y, acc[r].y);
            acc[r].x = fmaf(w.y, a[4 * hh + 1].x, acc[r].x); acc[r].y = fmaf(w.y, a[4 * hh + 1].y, acc[r].y);
            acc[r].x = fmaf(w.z, a[4 * hh + 2].x, acc[r].x); acc[r].y = fmaf(w.z, a[4 * hh + 2].y, acc[r].y);
            acc[r].x = fmaf(w.w, a[4 * hh + 3].x, acc[r].x); acc[r].y = fmaf(w.w, a[4 * hh + 3].y, acc[r].y);
          }
        }
      } else {
        for (int u = 0; l0 + u < L; ++u) {
          const unsigned v = *reinterpret_cast<const unsigned*>(chunk + u * H);
          const float2 x = PREC == ISC_PREC_BF16 ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v))
                                                 : __half22float2(*reinterpret_cast<const __half2*>(&v));
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const float w = sc_c[r * L + l0 + u];
            acc[r].x = fmaf(w, x.x, acc[r].x);
            acc[r].y = fmaf(w, x.y, acc[r].y);
          }
        }
      }
      __syncwarp();
      if (lane == 0) bulk::mbar_arrive(&empty[slot]);
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) p.cont_dst.store2(row0 + r, p.cont_col + c2, acc[r]);
  }
  if (senti_direct) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int l = 0; l < kSentiDirectMax; ++l)
        if (l < S) {
          const float w = sc_s[r * S + l];
          acc.x = fmaf(w, sv[l].x, acc.x);
          acc.y = fmaf(w, sv[l].y, acc.y);
        }
      p.senti_dst.store2(row0 + r, p.senti_col + c2, acc);
    }
  } else if (p.sw) {
    weighted_sum<float, RT>(p.sw + (long long)img * S * H, S, R, sc_s, p.senti_dst, row0, p.senti_col);
  }
}

template <int PREC>
static int launch_attention_tma(const AttnParams& p, int B, cudaStream_t stream) {
  const int Lp = (p.L + 3) & ~3, Sp = (p.S + 3) & ~3;
  const size_t smem = kTmaRingBytes + sizeof(float) * (2 * p.R * H + 2 * H + p.R * Lp + p.R * Sp + 2) + 2 * kTmaSlots * sizeof(uint64_t);
  void (*k)(AttnParams) = p.R == 1 ? attention_tma_kernel<PREC, 1> : attention_tma_kernel<PREC, 3>;
  ISC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ISC_CUDA(launch_pdl(k, dim3(B), dim3(256), smem, stream, p));
  ISC_LAUNCH_CHECK();
  return 0;
}

template <int PREC>
static int launch_attention_t(const AttnParams& p, int B, size_t smem, cudaStream_t stream) {
  void (*k)(AttnParams) = nullptr;
  switch (p.R) {
    case 1: k = attention_kernel<PREC, 1>; break;
    case 3: k = attention_kernel<PREC, 3>; break;
    case 5: k = attention_kernel<PREC, 5>; break;
    default: k = attention_kernel<PREC, 0>; break;
  }
  ISC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ISC_CUDA(launch_pdl(k, dim3(B), dim3(256), smem, stream, p));
  ISC_LAUNCH_CHECK();
  return 0;
}

int launch_attention(const AttnParams& p, int B, int precision, cudaStream_t stream) {
  ISC_REQUIRE(p.R >= 1 && p.R <= 8, "attention: rows per image %d not in 1..8", p.R);
  ISC_REQUIRE(!p.att16 || p.p_att16, "attention: att16 without p_att16");
  const int Lp = (p.L + 3) & ~3, Sp = (p.S + 3) & ~3;
  const int ring_bytes = (p.R == 1 || p.R == 3) ? kRingBytesReg : kRingBytes;
  size_t smem = sizeof(float) * (2 * p.R * H + 2 * H + p.R * Lp + p.R * Sp) + 8 * ring_bytes;
  // algorithmic HBM bytes: both feature tensors of every image once per launch (shared by its R rows),
  // sentiment-word features, the R query rows in and the R context rows out
  const bool fastp = precision != ISC_PREC_FP32 && p.p_att16 != nullptr;
  const double wide_b = precision == ISC_PREC_BF16 ? 2.0 : 4.0;
  const double feat_b = (fastp ? 2.0 : wide_b) + ((fastp && p.att16) ? 2.0 : wide_b);
  const double bytes = (double)B * ((p.att ? p.L * H * feat_b : 0.0) + (p.sw ? 2.0 * p.S * H * 4.0 : 0.0) +
                                    (double)p.R * (3.0 * H * 4.0 + 2.0 * H * 4.0));
  ProfScope ps(ISC_K_ATTENTION, bytes, stream);
  // TMA-staged kernel (measured slower than the per-warp rings, kept behind ISC_ATTN_TMA=1): the 16-bit copies exist,
  // 1 or 3 rows per image, 16-byte aligned score rows
  static const bool tma_on = getenv("ISC_ATTN_TMA") && atoi(getenv("ISC_ATTN_TMA")) != 0;
  const bool tma_ok = tma_on && fastp && p.att && (p.R == 1 || p.R == 3) && (p.L % 4) == 0 &&
                      (precision == ISC_PREC_BF16 || p.att16 != nullptr) &&
                      (reinterpret_cast<uintptr_t>(p.p_att16) % 16) == 0 &&
                      (reinterpret_cast<uintptr_t>(precision == ISC_PREC_BF16 ? p.att : p.att16) % 16) == 0;
  if (tma_ok) return precision == ISC_PREC_BF16 ? launch_attention_tma<ISC_PREC_BF16>(p, B, stream) : launch_attention_tma<ISC_PREC_BF16X3>(p, B, stream);
  if (precision == ISC_PREC_BF16) return launch_attention_t<ISC_PREC_BF16>(p, B, smem, stream);
  if (precision == ISC_PREC_BF16X3) return launch_attention_t<ISC_PREC_BF16X3>(p, B, smem, stream);
  return launch_attention_t<ISC_PREC_FP32>(p, B, smem, stream);
}

// --------------------------------------------------------------------------------------------
// Gate (Attention.forward :108-117): g3 = tanh(cont2att(c) + senti2att(s) + h2att(h)) comes from
// the GEMM epilogue; here w = sigmoid(att_alpha · g3 + b), ctx = w*c + (1-w)*s.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gate_mix_kernel(const float* __restrict__ g3, const float* __restrict__ cs,
                                                       const float* __restrict__ alpha, const float* __restrict__ alpha_b,
                                                       RowDest ctx, float* __restrict__ gate_w, long long ld_gate_w) {
  __shared__ float red[4];
  __shared__ float wsh;
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x;
  const int c = threadIdx.x * 4;
  float4 g = *reinterpret_cast<const float4*>(g3 + (long long)m * H + c);
  float4 a = *reinterpret_cast<const float4*>(alpha + c);
  float part = g.x * a.x + g.y * a.y + g.z * a.z + g.w * a.w;
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float w = sigmoid_accurate(red[0] + red[1] + red[2] + red[3] + alpha_b[0]);
    wsh = w;
    if (gate_w) gate_w[(long long)m * ld_gate_w] = w;
  }
  __syncthreads();
  const float w = wsh;
  float4 cv = *reinterpret_cast<const float4*>(cs + (long long)m * 2 * H + c);
  float4 sv = *reinterpret_cast<const float4*>(cs + (long long)m * 2 * H + H + c);
  float4 o;
  o.x = w * cv.x + (1.f - w) * sv.x;
  o.y = w * cv.y + (1.f - w) * sv.y;
  o.z = w * cv.z + (1.f - w) * sv.z;
  o.w = w * cv.w + (1.f - w) * sv.w;
  ctx.store4(m, c, o);
}

// --------------------------------------------------------------------------------------------
// Prologue gathers: ReLU(E[id]) rows; concept mean; senti-word rows with the prepended PAD.
// --------------------------------------------------------------------------------------------
// out[row] = ReLU(emb[ids[row]]);  ids == null -> out[row] = ReLU(emb[row]) (the whole table)
__global__ void __launch_bounds__(128) embed_rows_kernel(const long long* __restrict__ ids, long long n_per_group,
                                                         int prepend_pad, int pad_id, int V, const float* __restrict__ emb,
                                                         RowDest dst) {
  // rows are laid out [group][prepend_pad + n_per_group]; ids are [group][n_per_group]
  const long long row = blockIdx.x;
  const long long per = n_per_group + prepend_pad;
  const long long grp = row / per;
  const long long j = row - grp * per;
  long long tok = (prepend_pad && j == 0) ? pad_id : (ids ? ids[grp * n_per_group + (j - prepend_pad)] : row);
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  const int c = threadIdx.x * 4;
  float4 e = *reinterpret_cast<const float4*>(emb + tok * H + c);
  e.x = fmaxf(e.x, 0.f); e.y = fmaxf(e.y, 0.f); e.z = fmaxf(e.z, 0.f); e.w = fmaxf(e.w, 0.f);
  dst.store4(row, c, e);
}

// out[b] = mean_j ReLU(emb[ids[b][j]])   (captioner.py:297-298)
__global__ void __launch_bounds__(128) embed_mean_kernel(const long long* __restrict__ ids, int n, int V,
                                                         const float* __restrict__ emb, RowDest dst) {
  const long long b = blockIdx.x;
  const int c = threadIdx.x * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < n; ++j) {
    long long tok = ids[b * n + j];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    float4 e = *reinterpret_cast<const float4*>(emb + tok * H + c);
    acc.x += fmaxf(e.x, 0.f); acc.y += fmaxf(e.y, 0.f); acc.z += fmaxf(e.z, 0.f); acc.w += fmaxf(e.w, 0.f);
  }
  const float inv = 1.0f / (float)n;
  // torch.mean divides the fp32 sum by n
  acc.x = acc.x / (float)n; acc.y = acc.y / (float)n; acc.z = acc.z / (float)n; acc.w = acc.w / (float)n;
  (void)inv;
  dst.store4(b, c, acc);
}

__global__ void fill_kernel(float* p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------ host launchers
int launch_embed_pack(const long long* it, const int* parent, const float* h_in, int M, int V, const float* emb,
                      RowDest x1, RowDest x2, cudaStream_t stream) {
  ProfScope ps(ISC_K_POINTWISE, (double)M * H * 4.0 * 7, stream);
  ISC_CUDA(launch_pdl(embed_pack_kernel, dim3((unsigned)M), dim3(128), 0, stream, it, parent, h_in, M, V, emb, x1, x2));
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_lstm_pointwise(const float* gates, const int* parent, const float* c_prev, float* h_out, float* c_out,
                          RowDest extra, int extra_col, int M, cudaStream_t stream, const unsigned char* mask, float scale) {
  ProfScope ps(ISC_K_LSTM, (double)M * H * 4.0 * 8, stream);  // 4H gates + c in, h + c + h-copy out
  lstm_pointwise_kernel<<<M, 128, 0, stream>>>(gates, parent, c_prev, h_out, c_out, extra, extra_col, mask, scale);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_gate_mix(const float* g3, const float* cs, const float* alpha, const float* alpha_b, RowDest ctx,
                    float* gate_w, long long ld_gate_w, int M, cudaStream_t stream) {
  ProfScope ps(ISC_K_POINTWISE, (double)M * H * 4.0 * 4, stream);
  ISC_CUDA(launch_pdl(gate_mix_kernel, dim3((unsigned)M), dim3(128), 0, stream, g3, cs, alpha, alpha_b, ctx, gate_w, ld_gate_w));
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_embed_rows(const long long* ids, long long groups, long long n_per_group, int prepend_pad, int pad_id, int V,
                      const float* emb, RowDest dst, cudaStream_t stream) {
  long long rows = groups * (n_per_group + prepend_pad);
  if (rows <= 0) return 0;
  ProfScope ps(ISC_K_POINTWISE, (double)rows * H * 4.0 * 2, stream);
  embed_rows_kernel<<<(unsigned)rows, 128, 0, stream>>>(ids, n_per_group, prepend_pad, pad_id, V, emb, dst);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_embed_mean(const long long* ids, int B, int n, int V, const float* emb, RowDest dst, cudaStream_t stream) {
  ProfScope ps(ISC_K_POINTWISE, (double)B * H * 4.0 * (n + 1), stream);
  embed_mean_kernel<<<B, 128, 0, stream>>>(ids, n, V, emb, dst);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_fill(float* p, long long n, float v, cudaStream_t stream) {
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope ps(ISC_K_POINTWISE, (double)n * 4.0, stream);
  fill_kernel<<<blocks, 256, 0, stream>>>(p, n, v);
  ISC_LAUNCH_CHECK();
  return 0;
}

}  // namespace isc
