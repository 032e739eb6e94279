// Backward kernels of the teacher-forced decode (forward_xe / forward_seq2seq in train(), and the REINFORCE
// re-score of forward_rl samples): /root/reference/models/captioner.py:168-288 differentiated by hand. The dense
// contractions of the backward pass reuse the tcgen05 GEMM (gemm_tc.cu) on transposed operand planes; everything
// here is the element-wise / reduction glue between them, all HBM-bound: 128-bit accesses, warp-shuffle and
// shared-memory reductions, one CTA per row (per image for the attention).
#include "kernels.cuh"

namespace isc {
namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------------
// d logits = d logp - softmax * sum(d logp)   (log_softmax backward, captioner.py:183), one CTA per row.
// Dense form: dlogp row given. Fused-loss form (dlogp == null): the loss is sum coef[m] * (-logp[m, target[m]]), so
// d logp is -coef at the target and the row sum is -coef.
// ---------------------------------------------------------------------------------------------------------
// Input rows live in the [B, T, V] tensors (row b * T + t); output row blockIdx.x = t * M + b (step-major, the order of
// the tape), so one launch covers every step and the classifier's backward GEMMs contract over all T * B rows at once.
__global__ void __launch_bounds__(256) logsoftmax_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ dlogp,
                                                             const long long* __restrict__ target, long long ld_target,
                                                             const float* __restrict__ coef, int T, int M, int V,
                                                             float* __restrict__ dlogits, long long ld_out) {
  __shared__ float red[8];
  const int t = blockIdx.x / M, b = blockIdx.x - t * M;
  const long long m = (long long)b * T + t;  // row in the [B, T, .] inputs
  const long long ld_logp = V, ld_dlogp = V, ld_coef = 1;
  const float* lp = logp + m * ld_logp;
  float* out = dlogits + (long long)blockIdx.x * ld_out;
  float s = 0.f;
  int tgt = -1;
  float cf = 0.f;
  if (dlogp) {
    const float* g = dlogp + m * ld_dlogp;
    for (int i = threadIdx.x; i < V; i += 256) s += g[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
  } else {
    tgt = (int)target[(long long)b * ld_target + t];
    cf = coef[m * ld_coef];
    s = -cf;
  }
  for (int i = threadIdx.x; i < V; i += 256) {
    const float g = dlogp ? dlogp[m * ld_dlogp + i] : (i == tgt ? -cf : 0.f);
    out[i] = g - expf(lp[i]) * s;
  }
}

// ---------------------------------------------------------------------------------------------------------
// LSTM cell backward (nn.LSTMCell, gate order i,f,g,o). gates = pre-activations of this step.
//   dh = dh_a * (mask ? mask * scale : 1) + dh_b + dh_c     (any of them may be null)
//   dc = dc_carry + dh * o * (1 - tanh(c)^2);  dc_carry <- dc * f   (read-modify-write, zero before the last step)
// ---------------------------------------------------------------------------------------------------------
struct LstmBwd {
  const float* gates;   // [M,4H]
  const float* c_prev;  // [M,H] or null (zeros)
  const float* c_new;   // [M,H]
  const float* dh_a; long long ld_a; const unsigned char* mask; float scale;
  const float* dh_b; long long ld_b;
  const float* dh_c; long long ld_c;
  float* dc_carry;      // [M,H]
  float* dgates;        // [M,4H]
  RowDest dgates_planes;
};
__global__ void __launch_bounds__(128) lstm_bwd_kernel(LstmBwd p) {
  const long long m = blockIdx.x;
  const int c = threadIdx.x * 4;
  const float* g = p.gates + m * G4;
  const float4 gi = *reinterpret_cast<const float4*>(g + c), gf = *reinterpret_cast<const float4*>(g + H + c);
  const float4 gg = *reinterpret_cast<const float4*>(g + 2 * H + c), go = *reinterpret_cast<const float4*>(g + 3 * H + c);
  float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.c_prev) cp = *reinterpret_cast<const float4*>(p.c_prev + m * H + c);
  const float4 cn = *reinterpret_cast<const float4*>(p.c_new + m * H + c);
  float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.dh_a) {
    dh = *reinterpret_cast<const float4*>(p.dh_a + m * p.ld_a + c);
    if (p.mask) {
      const uchar4 k = *reinterpret_cast<const uchar4*>(p.mask + m * H + c);
      dh.x *= k.x ? p.scale : 0.f; dh.y *= k.y ? p.scale : 0.f; dh.z *= k.z ? p.scale : 0.f; dh.w *= k.w ? p.scale : 0.f;
    }
  }
  if (p.dh_b) {
    const float4 t = *reinterpret_cast<const float4*>(p.dh_b + m * p.ld_b + c);
    dh.x += t.x; dh.y += t.y; dh.z += t.z; dh.w += t.w;
  }
  if (p.dh_c) {
    const float4 t = *reinterpret_cast<const float4*>(p.dh_c + m * p.ld_c + c);
    dh.x += t.x; dh.y += t.y; dh.z += t.z; dh.w += t.w;
  }
  float4 dcc = *reinterpret_cast<const float4*>(p.dc_carry + m * H + c);
  float4 di, df, dg, dq, dcp;
#define ISC_LB(X)                                                  \
  {                                                                \
    const float i_ = sigmoidf_(gi.X), f_ = sigmoidf_(gf.X);        \
    const float g_ = tanhf(gg.X), o_ = sigmoidf_(go.X);            \
    const float tc = tanhf(cn.X);                                  \
    const float dc = dcc.X + dh.X * o_ * (1.f - tc * tc);          \
    dq.X = dh.X * tc * o_ * (1.f - o_);                            \
    di.X = dc * g_ * i_ * (1.f - i_);                              \
    df.X = dc * cp.X * f_ * (1.f - f_);                            \
    dg.X = dc * i_ * (1.f - g_ * g_);                              \
    dcp.X = dc * f_;                                               \
  }
  ISC_LB(x) ISC_LB(y) ISC_LB(z) ISC_LB(w)
#undef ISC_LB
  *reinterpret_cast<float4*>(p.dc_carry + m * H + c) = dcp;
  float* dgp = p.dgates + m * G4;
  *reinterpret_cast<float4*>(dgp + c) = di;
  *reinterpret_cast<float4*>(dgp + H + c) = df;
  *reinterpret_cast<float4*>(dgp + 2 * H + c) = dg;
  *reinterpret_cast<float4*>(dgp + 3 * H + c) = dq;
  p.dgates_planes.store4(m, c, di);
  p.dgates_planes.store4(m, H + c, df);
  p.dgates_planes.store4(m, 2 * H + c, dg);
  p.dgates_planes.store4(m, 3 * H + c, dq);
}

// ---------------------------------------------------------------------------------------------------------
// Gate backward (Attention.forward :108-117): ctx = w c + (1 - w) s, w = sigmoid(alpha . g3 + b),
// g3 = tanh(pre3). Writes dcs = [w dctx | (1 - w) dctx] and dpre3; accumulates d alpha, d b.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gate_bwd_kernel(const float* __restrict__ dctx, long long ld_dctx,
                                                       const float* __restrict__ cs, const float* __restrict__ g3,
                                                       const float* __restrict__ gate_w, const float* __restrict__ alpha,
                                                       float* __restrict__ dcs, RowDest dpre3, int dpre3_col,
                                                       float* __restrict__ dalpha, float* __restrict__ dalpha_b, int M) {
  // a CTA walks rows blockIdx.x, + gridDim.x, ... and keeps its d alpha / d b partial sums in registers: one atomicAdd
  // per column and CTA instead of one per column and ROW (2560 rows x 512 columns on 512 addresses took 111 us)
  __shared__ float red[4];
  const int c = threadIdx.x * 4;
  const float4 a = *reinterpret_cast<const float4*>(alpha + c);
  float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
  float db = 0.f;
  for (long long m = blockIdx.x; m < M; m += gridDim.x) {
    const float4 d = *reinterpret_cast<const float4*>(dctx + m * ld_dctx + c);
    const float4 cv = *reinterpret_cast<const float4*>(cs + m * 2 * H + c);
    const float4 sv = *reinterpret_cast<const float4*>(cs + m * 2 * H + H + c);
    float part = d.x * (cv.x - sv.x) + d.y * (cv.y - sv.y) + d.z * (cv.z - sv.z) + d.w * (cv.w - sv.w);
    part = warp_sum(part);
    __syncthreads();  // red[] of the previous row has been read by everyone
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    const float dw = red[0] + red[1] + red[2] + red[3];
    const float w = gate_w[m];
    const float dz = dw * w * (1.f - w);
    float4 o1, o2;
    o1.x = w * d.x; o1.y = w * d.y; o1.z = w * d.z; o1.w = w * d.w;
    o2.x = (1.f - w) * d.x; o2.y = (1.f - w) * d.y; o2.z = (1.f - w) * d.z; o2.w = (1.f - w) * d.w;
    *reinterpret_cast<float4*>(dcs + m * 2 * H + c) = o1;
    *reinterpret_cast<float4*>(dcs + m * 2 * H + H + c) = o2;
    const float4 g = *reinterpret_cast<const float4*>(g3 + m * H + c);
    float4 dp;
    dp.x = dz * a.x * (1.f - g.x * g.x); dp.y = dz * a.y * (1.f - g.y * g.y);
    dp.z = dz * a.z * (1.f - g.z * g.z); dp.w = dz * a.w * (1.f - g.w * g.w);
    dpre3.store4(m, dpre3_col + c, dp);
    da.x += dz * g.x; da.y += dz * g.y; da.z += dz * g.z; da.w += dz * g.w;
    db += dz;
  }
  atomicAdd(dalpha + c, da.x);
  atomicAdd(dalpha + c + 1, da.y);
  atomicAdd(dalpha + c + 2, da.z);
  atomicAdd(dalpha + c + 3, da.w);
  if (threadIdx.x == 0) atomicAdd(dalpha_b, db);
}

// ---------------------------------------------------------------------------------------------------------
// Additive attention backward for one image-row (ContentAttention :23-35 or SentiAttention :50-62):
//   e_l = alpha . tanh(p_l + q) ; w = softmax(e) ; ctx = sum_l w_l feat_l
// Given d ctx: d feat_l += w_l dctx, dw_l = dctx . feat_l, de_l = w_l (dw_l - sum_k w_k dw_k),
//   dpre_lj = de_l alpha_j (1 - t_lj^2): dp_l += dpre_l, dq = sum_l dpre_l, dalpha_j += sum_l de_l t_lj.
// ea holds exp(-2 p) (the e-product representation, common.cuh), q the raw query.
// ---------------------------------------------------------------------------------------------------------
// DEEP: two rows per warp and trip with the next two already requested (128 registers, 2 CTAs per SM): the kernel is
// bound by the bytes it keeps in flight, not by its arithmetic. DEEP = false is the one-row-per-trip form (70 registers,
// 3 CTAs per SM), kept behind ISC_ATTN_BWD_DEEP=0 for comparison.
template <bool DEEP>
__device__ void attn_bwd_part(const float* __restrict__ feat, const float* __restrict__ ea, int n_items,
                              const float* __restrict__ w_saved, const float* dctx_s, const float* eb_s,
                              const float* alpha_s, float* dw_s, float* red_s /*[8][2][H]*/, float* dfeat, float* dp,
                              float* dq_out /*smem [H]*/, float* dalpha_g /*global [H]*/, float* de_out /*[n_items] or null*/) {
  // dfeat / dp == null: deferred accumulation — this step only leaves de_l (and the caller d ctx) behind, and
  // attention_bwd_final_kernel builds d feat / d p for all steps in one pass instead of a read-modify-write per step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // phase 1: dw_l and d feat_l
  constexpr int R = DEEP ? 2 : 1;
  {
    float4 f[R][4], g[R][4];
    auto load_rows = [&](float4 (&dst)[R][4], int l0) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int l = l0 + 8 * u;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[u][i] = l < n_items ? __ldg(reinterpret_cast<const float4*>(feat + (long long)l * H + i * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    load_rows(f, warp);
    for (int l0 = warp; l0 < n_items; l0 += 8 * R) {
      if (DEEP) load_rows(g, l0 + 8 * R);
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int l = l0 + 8 * u;
        float acc = 0.f;
        const float wl = l < n_items ? w_saved[l] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = i * 128 + lane * 4;
          const float4 d = *reinterpret_cast<const float4*>(dctx_s + c);
          acc += f[u][i].x * d.x + f[u][i].y * d.y + f[u][i].z * d.z + f[u][i].w * d.w;
          if (dfeat && l < n_items) {
            float4* dst = reinterpret_cast<float4*>(dfeat + (long long)l * H + c);
            float4 o = *dst;
            o.x += wl * d.x; o.y += wl * d.y; o.z += wl * d.z; o.w += wl * d.w;
            *dst = o;
          }
        }
        acc = warp_sum(acc);
        if (lane == 0 && l < n_items) dw_s[l] = acc;
      }
      if (DEEP) {
#pragma unroll
        for (int u = 0; u < R; ++u)
#pragma unroll
          for (int i = 0; i < 4; ++i) f[u][i] = g[u][i];
      } else {
        load_rows(f, l0 + 8 * R);
      }
    }
  }
  __syncthreads();
  // phase 2: softmax backward (every thread computes the same scalar)
  float sdw = 0.f;
  for (int l = 0; l < n_items; ++l) sdw += w_saved[l] * dw_s[l];
  __syncthreads();
  for (int l = threadIdx.x; l < n_items; l += 256) {
    const float de = w_saved[l] * (dw_s[l] - sdw);  // de_l
    dw_s[l] = de;
    if (de_out) de_out[l] = de;
  }
  __syncthreads();
  // phase 3: through the tanh
  float dq[16], da[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dq[i] = da[i] = 0.f;
  for (int l0 = warp; l0 < n_items; l0 += 8 * R) {
    float4 e4[R][4];  // all rows' loads requested before the first is used
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int l = l0 + 8 * u;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        e4[u][i] = l < n_items ? __ldg(reinterpret_cast<const float4*>(ea + (long long)l * H + i * 128 + lane * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int l = l0 + 8 * u;
      if (l >= n_items) break;
      const float de = dw_s[l];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 b4 = *reinterpret_cast<const float4*>(eb_s + c);
        const float4 a4 = *reinterpret_cast<const float4*>(alpha_s + c);
        float t[4];
        tanh2_eprod(e4[u][i].x, e4[u][i].y, b4.x, b4.y, t[0], t[1]);
        tanh2_eprod(e4[u][i].z, e4[u][i].w, b4.z, b4.w, t[2], t[3]);
        const float al[4] = {a4.x, a4.y, a4.z, a4.w};
        float dpre[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dpre[j] = de * al[j] * (1.f - t[j] * t[j]);
          dq[i * 4 + j] += dpre[j];
          da[i * 4 + j] += de * t[j];
        }
        if (dp) {
          float4* dst = reinterpret_cast<float4*>(dp + (long long)l * H + c);
          float4 o = *dst;
          o.x += dpre[0]; o.y += dpre[1]; o.z += dpre[2]; o.w += dpre[3];
          *dst = o;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red_s[(warp * 2 + 0) * H + i * 128 + lane * 4 + j] = dq[i * 4 + j];
      red_s[(warp * 2 + 1) * H + i * 128 + lane * 4 + j] = da[i * 4 + j];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += 256) {
    float sq = 0.f, sa = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) {
      sq += red_s[(w8 * 2 + 0) * H + c];
      sa += red_s[(w8 * 2 + 1) * H + c];
    }
    dq_out[c] = sq;
    atomicAdd(dalpha_g + c, sa);
  }
  __syncthreads();
}

struct AttnBwd {
  int L, S;
  const float* dcs; long long ld_dcs; int cont_col, senti_col;  // d ctx of the content / sentiment attention
  const float* hproj; long long ld_hproj;                       // [M,3H] queries of this step
  const float* pre_word;                                        // [B,H]
  const float* att; const float* ea_att;                        // [B,L,H]; null -> no content attention
  const float* sw; const float* ea_sw;                          // [B,S,H]; null -> no sentiment attention
  const float* cont_w; const float* senti_w;                    // saved softmax weights [M,L], [M,S]
  const float* alpha_c; const float* alpha_s;
  float* datt; float* dp_att; float* dsw; float* dp_sw;         // accumulators, same shapes as the features
  RowDest dhproj;                                               // [M,3H]: cols 0..H-1 <- dq_c, H..2H-1 <- dq_s
  float* dpre_word;                                             // [B,H] += dq_s
  float* dalpha_c; float* dalpha_s;                             // [H] +=
  // deferred mode (datt == null): this step's softmax gradients and d ctx, for attention_bwd_final_kernel
  float* de_c; float* de_s;                                     // [M,L], [M,S]
  float* dctx_c; float* dctx_s_out;                             // [M,H] each
};
template <bool DEEP>
__global__ void __launch_bounds__(256) attention_bwd_kernel(AttnBwd p) {
  extern __shared__ __align__(16) float sm[];
  float* dctx_s = sm;             // [H]
  float* eb_s = dctx_s + H;       // [H]
  float* alpha_s = eb_s + H;      // [H]
  float* dq_s = alpha_s + H;      // [H]
  float* dw_s = dq_s + H;         // [max(L,S) padded]
  const int nmax = ((p.L > p.S ? p.L : p.S) + 3) & ~3;
  float* red_s = dw_s + nmax;     // [8][2][H]
  const long long m = blockIdx.x;  // one row per image in training
  if (p.att) {
    for (int c = threadIdx.x; c < H; c += 256) {
      dctx_s[c] = p.dcs[m * p.ld_dcs + p.cont_col + c];
      eb_s[c] = exp_neg2(p.hproj[m * p.ld_hproj + c]);
      alpha_s[c] = p.alpha_c[c];
    }
    __syncthreads();
    if (p.dctx_c)
      for (int c = threadIdx.x; c < H; c += 256) p.dctx_c[m * H + c] = dctx_s[c];
    attn_bwd_part<DEEP>(p.att + m * p.L * H, p.ea_att + m * p.L * H, p.L, p.cont_w + m * p.L, dctx_s, eb_s, alpha_s, dw_s, red_s,
                  p.datt ? p.datt + m * p.L * H : nullptr, p.datt ? p.dp_att + m * p.L * H : nullptr, dq_s, p.dalpha_c,
                  p.de_c ? p.de_c + m * p.L : nullptr);
    for (int c = threadIdx.x * 4; c < H; c += 1024) p.dhproj.store4(m, c, *reinterpret_cast<float4*>(dq_s + c));
    __syncthreads();
  }
  if (p.sw) {
    for (int c = threadIdx.x; c < H; c += 256) {
      dctx_s[c] = p.dcs[m * p.ld_dcs + p.senti_col + c];
      eb_s[c] = exp_neg2(p.hproj[m * p.ld_hproj + H + c] + (p.pre_word ? p.pre_word[m * H + c] : 0.f));
      alpha_s[c] = p.alpha_s[c];
    }
    __syncthreads();
    if (p.dctx_s_out)
      for (int c = threadIdx.x; c < H; c += 256) p.dctx_s_out[m * H + c] = dctx_s[c];
    attn_bwd_part<DEEP>(p.sw + m * p.S * H, p.ea_sw + m * p.S * H, p.S, p.senti_w + m * p.S, dctx_s, eb_s, alpha_s, dw_s, red_s,
                  p.dsw ? p.dsw + m * p.S * H : nullptr, p.dsw ? p.dp_sw + m * p.S * H : nullptr, dq_s, p.dalpha_s,
                  p.de_s ? p.de_s + m * p.S : nullptr);
    for (int c = threadIdx.x * 4; c < H; c += 1024) {
      const float4 v = *reinterpret_cast<float4*>(dq_s + c);
      p.dhproj.store4(m, H + c, v);
      if (p.dpre_word) {
        float4* d = reinterpret_cast<float4*>(p.dpre_word + m * H + c);
        float4 o = *d;
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
        *d = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Deferred accumulation of the attention feature gradients over all T steps, one pass per image:
//   d feat[l, j] = sum_t w[t, l] dctx[t, j]
//   d p[l, j]    = sum_t de[t, l] alpha_j (1 - tanh^2(p[l, j] + q[t, j]))
// (plain stores: no zero-init, no per-step read-modify-write of the [B, L, H] accumulators).
// ---------------------------------------------------------------------------------------------------------
struct AttnBwdFinal {
  int T, M, n_items;
  const float* ea;        // [B, n, H] exp(-2 p)
  const float* w_all;     // [T, M, n] softmax weights
  const float* de_all;    // [T, M, n]
  const float* dctx_all;  // [T, M, H]
  const float* hproj;     // [T, M, ld_hproj] queries; column q_col
  long long ld_hproj;
  int q_col;
  const float* pre_word;  // [B, H] added to the query (sentiment attention) or null
  const float* alpha;     // [H]
  float* dfeat;           // [B, n, H]
  float* dp;              // [B, n, H]
};
__global__ void __launch_bounds__(256) attention_bwd_final_kernel(AttnBwdFinal p) {
  extern __shared__ __align__(16) float sm[];
  float* dctx_s = sm;                 // [T][H]
  float* eb_s = dctx_s + p.T * H;     // [T][H]
  float* alpha_s = eb_s + p.T * H;    // [H]
  const long long b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.T * H; i += 256) {
    const int t = i / H, c = i - t * H;
    const long long row = (long long)t * p.M + b;
    dctx_s[i] = p.dctx_all[row * H + c];
    eb_s[i] = exp_neg2(p.hproj[row * p.ld_hproj + p.q_col + c] + (p.pre_word ? p.pre_word[b * H + c] : 0.f));
  }
  for (int i = threadIdx.x; i < H; i += 256) alpha_s[i] = p.alpha[i];
  __syncthreads();
  // gridDim.y CTAs share an image's items (each re-reads the 2 * T * H query / d ctx rows, which are L2 hits): one CTA per
  // image left the 148 SMs with 1.7 CTAs each
  for (int l = blockIdx.y * 8 + warp; l < p.n_items; l += 8 * gridDim.y) {
    float af[16], ap[16], e[16], al[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 128 + lane * 4;
      const float4 e4 = *reinterpret_cast<const float4*>(p.ea + (b * p.n_items + l) * H + c);
      const float4 a4 = *reinterpret_cast<const float4*>(alpha_s + c);
      e[4 * i] = e4.x; e[4 * i + 1] = e4.y; e[4 * i + 2] = e4.z; e[4 * i + 3] = e4.w;
      al[4 * i] = a4.x; al[4 * i + 1] = a4.y; al[4 * i + 2] = a4.z; al[4 * i + 3] = a4.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) af[i] = ap[i] = 0.f;
    // lane t fetches step t's softmax weight and its gradient once (they sit T * n_items apart: read inside the step loop,
    // every trip waited for two L2 round trips), the step loop takes them by shuffle
    for (int t0 = 0; t0 < p.T; t0 += 32) {
    float w_l = 0.f, de_l = 0.f;
    if (t0 + lane < p.T) {
      const long long row = (long long)(t0 + lane) * p.M + b;
      w_l = p.w_all[row * p.n_items + l];
      de_l = p.de_all[row * p.n_items + l];
    }
    const int tn = p.T - t0 < 32 ? p.T - t0 : 32;
    for (int tt_ = 0; tt_ < tn; ++tt_) {
      const int t = t0 + tt_;
      const float w = __shfl_sync(0xffffffffu, w_l, tt_);
      const float de = __shfl_sync(0xffffffffu, de_l, tt_);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 d4 = *reinterpret_cast<const float4*>(dctx_s + t * H + c);
        const float4 b4 = *reinterpret_cast<const float4*>(eb_s + t * H + c);
        float tt[4];
        tanh2_eprod(e[4 * i], e[4 * i + 1], b4.x, b4.y, tt[0], tt[1]);
        tanh2_eprod(e[4 * i + 2], e[4 * i + 3], b4.z, b4.w, tt[2], tt[3]);
        const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          af[4 * i + j] = fmaf(w, dd[j], af[4 * i + j]);
          ap[4 * i + j] = fmaf(de * al[4 * i + j], 1.f - tt[j] * tt[j], ap[4 * i + j]);
        }
      }
    }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 128 + lane * 4;
      *reinterpret_cast<float4*>(p.dfeat + (b * p.n_items + l) * H + c) = make_float4(af[4 * i], af[4 * i + 1], af[4 * i + 2], af[4 * i + 3]);
      *reinterpret_cast<float4*>(p.dp + (b * p.n_items + l) * H + c) = make_float4(ap[4 * i], ap[4 * i + 1], ap[4 * i + 2], ap[4 * i + 3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Scatter-add into an embedding table: dE[ids[r]] += relu'(E[ids[r]]) * mask * scale * g[r] * gscale
// (word_embed / senti_label_embed are Sequential(Embedding, ReLU); the PAD row of word_embed gets no gradient).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_bwd_kernel(const long long* __restrict__ ids, long long ld_ids,
                                                        long long ids_per_group, int prepend_pad, int pad_id, int skip_pad,
                                                        int V, const float* __restrict__ emb, const float* __restrict__ g,
                                                        long long ld_g, long long g_rows_per_group,
                                                        const unsigned char* __restrict__ mask, float scale, float gscale,
                                                        float* __restrict__ demb) {
  // rows are [group][prepend_pad + ids_per_group]; g row = group * g_rows_per_group + (g_rows_per_group > 1 ? j : 0)
  const long long row = blockIdx.x;
  const long long per = ids_per_group + prepend_pad;
  const long long grp = row / per, j = row - grp * per;
  long long tok = (prepend_pad && j == 0) ? pad_id : ids[grp * ld_ids + (j - prepend_pad)];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  if (skip_pad && tok == pad_id) return;
  const long long grow = grp * g_rows_per_group + (g_rows_per_group > 1 ? j : 0);
  const int c = threadIdx.x * 4;
  const float4 e = *reinterpret_cast<const float4*>(emb + tok * H + c);
  float4 d = *reinterpret_cast<const float4*>(g + grow * ld_g + c);
  float k[4] = {gscale, gscale, gscale, gscale};
  if (mask) {
    const uchar4 mk = *reinterpret_cast<const uchar4*>(mask + grow * H + c);
    k[0] *= mk.x ? scale : 0.f; k[1] *= mk.y ? scale : 0.f; k[2] *= mk.z ? scale : 0.f; k[3] *= mk.w ? scale : 0.f;
  }
  float* dst = demb + tok * H + c;
  if (e.x > 0.f) atomicAdd(dst, d.x * k[0]);
  if (e.y > 0.f) atomicAdd(dst + 1, d.y * k[1]);
  if (e.z > 0.f) atomicAdd(dst + 2, d.z * k[2]);
  if (e.w > 0.f) atomicAdd(dst + 3, d.w * k[3]);
}

// ---------------------------------------------------------------------------------------------------------
// Generic pointwise helpers over fp32 [rows, cols] matrices (cols % 4 == 0).
// ---------------------------------------------------------------------------------------------------------
// out = (a + b) * [gate > 0 (or ea-form: gate < 1)] * (mask ? mask * scale : 1), also to planes
__global__ void relu_mask_bwd_kernel(const float* __restrict__ a, long long ld_a, const float* __restrict__ b, long long ld_b,
                                     const float* __restrict__ gate, long long ld_gate, int gate_is_exp,
                                     const unsigned char* __restrict__ mask, long long ld_mask, float scale, long long rows,
                                     int cols, float* __restrict__ out, long long ld_out, RowDest planes) {
  const int c4n = cols >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * c4n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    float4 v = *reinterpret_cast<const float4*>(a + r * ld_a + c);
    if (b) {
      const float4 t = *reinterpret_cast<const float4*>(b + r * ld_b + c);
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (mask) {
      const uchar4 k = *reinterpret_cast<const uchar4*>(mask + r * ld_mask + c);
      v.x *= k.x ? scale : 0.f; v.y *= k.y ? scale : 0.f; v.z *= k.z ? scale : 0.f; v.w *= k.w ? scale : 0.f;
    }
    if (gate) {
      const float4 g = *reinterpret_cast<const float4*>(gate + r * ld_gate + c);
      if (gate_is_exp) {
        v.x = g.x < 1.f ? v.x : 0.f; v.y = g.y < 1.f ? v.y : 0.f; v.z = g.z < 1.f ? v.z : 0.f; v.w = g.w < 1.f ? v.w : 0.f;
      } else {
        v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f; v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
      }
    }
    if (out) *reinterpret_cast<float4*>(out + r * ld_out + c) = v;
    planes.store4(r, c, v);
  }
}

// x *= mask * scale in place (fp32) and refresh the planes (dropout forward)
__global__ void apply_mask_kernel(float* __restrict__ x, long long ld, const unsigned char* __restrict__ mask, float scale,
                                  long long rows, int cols, RowDest planes) {
  const int c4n = cols >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * c4n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
    const uchar4 k = *reinterpret_cast<const uchar4*>(mask + r * cols + c);
    v.x *= k.x ? scale : 0.f; v.y *= k.y ? scale : 0.f; v.z *= k.z ? scale : 0.f; v.w *= k.w ? scale : 0.f;
    *reinterpret_cast<float4*>(x + r * ld + c) = v;
    planes.store4(r, c, v);
  }
}

// Tiled batches (isc_dims_t::att_tile = R: R consecutive rows share one image): dst row block b = src row block b / R,
// times the row's own dropout mask (or unmasked when mask is null); dst is fp32 + planes. `per` = elements per image.
__global__ void expand_mask_kernel(const float* __restrict__ src, int R, long long per, long long n_dst_images,
                                   const unsigned char* __restrict__ mask, float scale, float* __restrict__ dst, int cols,
                                   RowDest planes) {
  const long long per4 = per >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_dst_images * per4; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per4, e = (i - b * per4) * 4;
    float4 v = *reinterpret_cast<const float4*>(src + (b / R) * per + e);
    if (mask) {
      const uchar4 k = *reinterpret_cast<const uchar4*>(mask + b * per + e);
      v.x *= k.x ? scale : 0.f; v.y *= k.y ? scale : 0.f; v.z *= k.z ? scale : 0.f; v.w *= k.w ? scale : 0.f;
    }
    *reinterpret_cast<float4*>(dst + b * per + e) = v;
    const long long flat = b * per + e;
    planes.store4(flat / cols, (int)(flat % cols), v);
  }
}
// dst image i = sum of the R consecutive src images i * R .. i * R + R - 1 (in that order)
__global__ void sum_tiles_kernel(const float* __restrict__ src, int R, long long per, long long n_images, float* __restrict__ dst) {
  const long long per4 = per >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_images * per4; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per4, e = (i - b * per4) * 4;
    float4 a = *reinterpret_cast<const float4*>(src + (b * R) * per + e);
    for (int r = 1; r < R; ++r) {
      const float4 t = *reinterpret_cast<const float4*>(src + (b * R + r) * per + e);
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    *reinterpret_cast<float4*>(dst + b * per + e) = a;
  }
}

// dst[c] += sum_r src[r, c]; a block sums 32 columns x up to 512 rows (8 row-lanes) and adds its partial atomically
__global__ void __launch_bounds__(256) colsum_add_kernel(const float* __restrict__ src, long long ld, long long rows, int cols,
                                                         float* __restrict__ dst, float* __restrict__ dst2) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const long long r_begin = (long long)blockIdx.y * 512;
  const long long r_end = r_begin + 512 < rows ? r_begin + 512 : rows;
  float s = 0.f;
  if (c < cols)
    for (long long r = r_begin + ry; r < r_end; r += 8) s += src[r * ld + c];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    atomicAdd(dst + c, t);
    if (dst2) atomicAdd(dst2 + c, t);
  }
}

// dst[m, :] = sum_t src[t, m, :]
__global__ void sum_steps_kernel(const float* __restrict__ src, int T, long long stride_t, long long n4, float* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float4 v = reinterpret_cast<const float4*>(src + t * stride_t)[i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(dst)[i] = s;
  }
}

// dst[r, c] += src[r, c] (2D, different pitches)
__global__ void add2d_kernel(float* __restrict__ dst, long long ld_dst, const float* __restrict__ src, long long ld_src,
                             long long rows, int cols) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * cols; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ld_dst + c] += src[r * ld_src + c];
  }
}

// Transposed planes: dst_hi/lo[c, col0 + r] = split(src[r, c]); 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ src, long long ld_src, long long rows,
                                                              int cols, __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo, long long ld_dst, long long col0) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + ty + 8 * i;
    const int c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    const long long r = r0 + tx;
    if (c < cols && r < rows) {
      __nv_bfloat16 h, l;
      split_bf16(tile[tx][ty + 8 * i], h, l);
      hi[(long long)c * ld_dst + col0 + r] = h;
      if (lo) lo[(long long)c * ld_dst + col0 + r] = l;
    }
  }
}

// bf16 plane transpose: dst[c, col0 + r] = src[r, c]
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src,
                                                             long long rows, int cols, __nv_bfloat16* __restrict__ dst,
                                                             long long ld_dst, long long col0) {
  __shared__ __nv_bfloat16 tile[32][34];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + ty + 8 * i;
    const int c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    const long long r = r0 + tx;
    if (c < cols && r < rows) dst[(long long)c * ld_dst + col0 + r] = tile[tx][ty + 8 * i];
  }
}

// fused element-wise clamp + Adam (train_xe.py:19-23 clip_gradient, then torch.optim.Adam with weight decay)
__global__ void adam_clamp_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                  long long n, float clip, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  float bias1, float bias2, float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    const float pi = p[i];
    gi += weight_decay * pi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bias2) + eps;
    p[i] = pi - (lr / bias1) * (mi / denom);
  }
}

int grid_for(long long n, int per_block = 256) {
  long long b = (n + per_block - 1) / per_block;
  if (b > 148 * 16) b = 148 * 16;
  return b < 1 ? 1 : (int)b;
}

}  // namespace

// ------------------------------------------------------------------ host launchers
int launch_logsoftmax_bwd(const float* logp, const float* dlogp, const long long* target, long long ld_target,
                          const float* coef, int T, int M, int V, float* dlogits, long long ld_out, cudaStream_t s) {
  ProfScope ps(ISC_K_TRAIN, (double)T * M * V * 4.0 * (dlogp ? 4 : 2), s);
  logsoftmax_bwd_kernel<<<T * M, 256, 0, s>>>(logp, dlogp, target, ld_target, coef, T, M, V, dlogits, ld_out);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_lstm_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh_a, long long ld_a,
                    const unsigned char* mask, float scale, const float* dh_b, long long ld_b, const float* dh_c,
                    long long ld_c, float* dc_carry, float* dgates, RowDest planes, int M, cudaStream_t s) {
  LstmBwd p;
  p.gates = gates; p.c_prev = c_prev; p.c_new = c_new;
  p.dh_a = dh_a; p.ld_a = ld_a; p.mask = mask; p.scale = scale;
  p.dh_b = dh_b; p.ld_b = ld_b; p.dh_c = dh_c; p.ld_c = ld_c;
  p.dc_carry = dc_carry; p.dgates = dgates; p.dgates_planes = planes;
  ProfScope ps(ISC_K_TRAIN, (double)M * H * 4.0 * 14, s);
  lstm_bwd_kernel<<<M, 128, 0, s>>>(p);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_gate_bwd(const float* dctx, long long ld_dctx, const float* cs, const float* g3, const float* gate_w,
                    const float* alpha, float* dcs, RowDest dpre3, int dpre3_col, float* dalpha, float* dalpha_b, int M,
                    cudaStream_t s) {
  ProfScope ps(ISC_K_TRAIN, (double)M * H * 4.0 * 8, s);
  const int grid = M < 592 ? M : 592;  // 4 CTAs per SM of a B200; every CTA adds its partial sums once
  gate_bwd_kernel<<<grid, 128, 0, s>>>(dctx, ld_dctx, cs, g3, gate_w, alpha, dcs, dpre3, dpre3_col, dalpha, dalpha_b, M);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_attention_bwd(const AttnBwdParams& a, int M, cudaStream_t s) {
  AttnBwd p;
  p.L = a.L; p.S = a.S;
  p.dcs = a.dcs; p.ld_dcs = a.ld_dcs; p.cont_col = a.cont_col; p.senti_col = a.senti_col;
  p.hproj = a.hproj; p.ld_hproj = a.ld_hproj; p.pre_word = a.pre_word;
  p.att = a.att; p.ea_att = a.ea_att; p.sw = a.sw; p.ea_sw = a.ea_sw;
  p.cont_w = a.cont_w; p.senti_w = a.senti_w; p.alpha_c = a.alpha_c; p.alpha_s = a.alpha_s;
  p.datt = a.datt; p.dp_att = a.dp_att; p.dsw = a.dsw; p.dp_sw = a.dp_sw;
  p.dhproj = a.dhproj; p.dpre_word = a.dpre_word; p.dalpha_c = a.dalpha_c; p.dalpha_s = a.dalpha_s;
  p.de_c = a.de_c; p.de_s = a.de_s; p.dctx_c = a.dctx_c; p.dctx_s_out = a.dctx_s_out;
  const int nmax = ((a.L > a.S ? a.L : a.S) + 3) & ~3;
  const size_t smem = sizeof(float) * (4 * H + nmax + 16 * H);
  static const int deep_env = getenv("ISC_ATTN_BWD_DEEP") ? atoi(getenv("ISC_ATTN_BWD_DEEP")) : -1;  // experiments: force the variant
  const bool deep = deep_env != 0;  // measured faster at M = 256 (XE: 12.9 -> 12.3 ms) and at M = 2560 (SCST: 96 -> 92 ms)
  ISC_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ISC_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const double bytes = (double)M * ((a.att ? 6.0 * a.L * H * 4.0 : 0.0) + (a.sw ? 6.0 * a.S * H * 4.0 : 0.0));
  ProfScope ps(ISC_K_TRAIN, bytes, s);
  if (deep) attention_bwd_kernel<true><<<M, 256, smem, s>>>(p);
  else attention_bwd_kernel<false><<<M, 256, smem, s>>>(p);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_attention_bwd_final(int T, int M, int B, int n_items, const float* ea, const float* w_all, const float* de_all,
                               const float* dctx_all, const float* hproj, long long ld_hproj, int q_col, const float* pre_word,
                               const float* alpha, float* dfeat, float* dp, cudaStream_t s) {
  AttnBwdFinal p;
  p.T = T; p.M = M; p.n_items = n_items; p.ea = ea; p.w_all = w_all; p.de_all = de_all; p.dctx_all = dctx_all;
  p.hproj = hproj; p.ld_hproj = ld_hproj; p.q_col = q_col; p.pre_word = pre_word; p.alpha = alpha; p.dfeat = dfeat; p.dp = dp;
  const size_t smem = sizeof(float) * ((size_t)2 * T * H + H);
  ISC_REQUIRE(smem <= 200 * 1024, "attention_bwd_final: %d steps do not fit shared memory", T);
  ISC_CUDA(cudaFuncSetAttribute(attention_bwd_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(ISC_K_TRAIN, (double)B * n_items * H * 4.0 * 3, s);
  const int split = n_items >= 64 ? 3 : 1;  // region features (196 items): 3 CTAs per image, 3 resident per SM
  attention_bwd_final_kernel<<<dim3(B, split), 256, smem, s>>>(p);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_embed_bwd(const long long* ids, long long ld_ids, long long groups, long long ids_per_group, int prepend_pad,
                     int pad_id, int skip_pad, int V, const float* emb, const float* g, long long ld_g,
                     long long g_rows_per_group, const unsigned char* mask, float scale, float gscale, float* demb,
                     cudaStream_t s) {
  const long long rows = groups * (ids_per_group + prepend_pad);
  if (rows <= 0) return 0;
  ProfScope ps(ISC_K_TRAIN, (double)rows * H * 4.0 * 3, s);
  embed_bwd_kernel<<<(unsigned)rows, 128, 0, s>>>(ids, ld_ids, ids_per_group, prepend_pad, pad_id, skip_pad, V, emb, g, ld_g,
                                                   g_rows_per_group, mask, scale, gscale, demb);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_relu_mask_bwd(const float* a, long long ld_a, const float* b, long long ld_b, const float* gate, long long ld_gate,
                         int gate_is_exp, const unsigned char* mask, long long ld_mask, float scale, long long rows, int cols,
                         float* out, long long ld_out, RowDest planes, cudaStream_t s) {
  if (rows <= 0) return 0;
  ProfScope ps(ISC_K_TRAIN, (double)rows * cols * 4.0 * 4, s);
  relu_mask_bwd_kernel<<<grid_for(rows * (cols / 4)), 256, 0, s>>>(a, ld_a, b, ld_b, gate, ld_gate, gate_is_exp, mask, ld_mask,
                                                                    scale, rows, cols, out, ld_out, planes);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_apply_mask(float* x, long long ld, const unsigned char* mask, float scale, long long rows, int cols, RowDest planes,
                      cudaStream_t s) {
  if (rows <= 0) return 0;
  ProfScope ps(ISC_K_TRAIN, (double)rows * cols * 9.0, s);
  apply_mask_kernel<<<grid_for(rows * (cols / 4)), 256, 0, s>>>(x, ld, mask, scale, rows, cols, planes);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_expand_mask(const float* src, int R, long long per, long long n_dst_images, const unsigned char* mask, float scale,
                       float* dst, int cols, RowDest planes, cudaStream_t s) {
  if (n_dst_images <= 0) return 0;
  ProfScope ps(ISC_K_TRAIN, (double)n_dst_images * per * 9.0, s);
  expand_mask_kernel<<<grid_for(n_dst_images * (per / 4)), 256, 0, s>>>(src, R, per, n_dst_images, mask, scale, dst, cols, planes);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_sum_tiles(const float* src, int R, long long per, long long n_images, float* dst, cudaStream_t s) {
  if (n_images <= 0) return 0;
  ProfScope ps(ISC_K_TRAIN, (double)n_images * per * 4.0 * (R + 1), s);
  sum_tiles_kernel<<<grid_for(n_images * (per / 4)), 256, 0, s>>>(src, R, per, n_images, dst);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_colsum_add(const float* src, long long ld, long long rows, int cols, float* dst, float* dst2, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  ProfScope ps(ISC_K_TRAIN, (double)rows * cols * 4.0, s);
  dim3 grid((cols + 31) / 32, (unsigned)((rows + 511) / 512));
  colsum_add_kernel<<<grid, 256, 0, s>>>(src, ld, rows, cols, dst, dst2);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_sum_steps(const float* src, int T, long long stride_t, long long n, float* dst, cudaStream_t s) {
  ProfScope ps(ISC_K_TRAIN, (double)n * 4.0 * (T + 1), s);
  sum_steps_kernel<<<grid_for(n / 4), 256, 0, s>>>(src, T, stride_t, n / 4, dst);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_add2d(float* dst, long long ld_dst, const float* src, long long ld_src, long long rows, int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  ProfScope ps(ISC_K_TRAIN, (double)rows * cols * 12.0, s);
  add2d_kernel<<<grid_for(rows * cols), 256, 0, s>>>(dst, ld_dst, src, ld_src, rows, cols);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_split_transpose(const float* src, long long ld_src, long long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo,
                           long long ld_dst, long long col0, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  ProfScope ps(ISC_K_TRAIN, (double)rows * cols * 8.0, s);
  split_transpose_kernel<<<grid, 256, 0, s>>>(src, ld_src, rows, cols, hi, lo, ld_dst, col0);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_transpose_bf16(const __nv_bfloat16* src, long long ld_src, long long rows, int cols, __nv_bfloat16* dst,
                          long long ld_dst, long long col0, cudaStream_t s) {
  if (rows <= 0 || cols <= 0 || !src) return 0;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  ProfScope ps(ISC_K_TRAIN, (double)rows * cols * 4.0, s);
  transpose_bf16_kernel<<<grid, 256, 0, s>>>(src, ld_src, rows, cols, dst, ld_dst, col0);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_adam_clamp(float* p, const float* g, float* m, float* v, long long n, float clip, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int step, float grad_scale, cudaStream_t s) {
  if (n <= 0) return 0;
  const float bias1 = 1.f - powf(beta1, (float)step), bias2 = 1.f - powf(beta2, (float)step);
  ProfScope ps(ISC_K_TRAIN, (double)n * 4.0 * 7, s);
  adam_clamp_kernel<<<grid_for(n), 256, 0, s>>>(p, g, m, v, n, clip, lr, beta1, beta2, eps, weight_decay, bias1, bias2, grad_scale);
  ISC_LAUNCH_CHECK();
  return 0;
}

}  // namespace isc
