// Host-side engine shared by the C-ABI translation units (api.cu: decode, api_train.cu: training): argument checks,
// the packed-weight and workspace layouts, the prologue (Captioner.forward_* :198-214 / :247-261 / :294-315) and one
// decode step (Captioner.forward_step :168-186) as launch sequences. Everything lives in an anonymous namespace:
// each translation unit gets its own copy of these small host functions (internal to libisc_b200.so, not ABI).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"

namespace isc {
namespace {

typedef __nv_bfloat16 bf16;
constexpr int T_MAX = 64;  // longest caption the beam/greedy bookkeeping buffers are sized for

int check_device() {
  int dev = 0, major = 0, minor = 0;
  ISC_CUDA(cudaGetDevice(&dev));
  ISC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  ISC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("device %d is sm_%d%d; libisc_b200 only contains sm_100a code and has no fallback", dev, major, minor);
    return ISC_ERR_DEVICE;
  }
  return 0;
}

int check_dims(const isc_dims_t* d) {
  ISC_REQUIRE(d != nullptr, "dims is NULL");
  ISC_REQUIRE(d->hidden == H, "hidden=%d: only %d is compiled in", d->hidden, H);
  ISC_REQUIRE(d->vocab >= 8 && d->vocab <= 65534, "vocab=%d out of range [8, 65534]", d->vocab);
  ISC_REQUIRE(d->feat_dim > 0 && d->feat_dim % 8 == 0, "feat_dim=%d must be a positive multiple of 8", d->feat_dim);
  ISC_REQUIRE(d->n_regions > 0 && d->n_senti > 0 && d->n_labels > 0, "n_regions/n_senti/n_labels must be positive");
  ISC_REQUIRE(d->att_tile >= 0 && d->att_tile <= 64, "att_tile=%d out of range [0, 64]", d->att_tile);
  return 0;
}
int check_precision(int p) {
  ISC_REQUIRE(p == ISC_PREC_FP32 || p == ISC_PREC_BF16X3 || p == ISC_PREC_BF16, "unknown precision %d", p);
  return 0;
}

// Representation of the PROJECTED attention features (feats.p_att, feats.p_sw): ReLU(.) as in the reference,
// except in ISC_PREC_BF16X3 where the attention kernel's e-product tanh reads exp(-2 * ReLU(.)).
int proj_act(int precision) { return precision == ISC_PREC_BF16X3 ? ACT_EXPNEG2_RELU : ACT_RELU; }

__global__ void proj_convert_kernel(const float* __restrict__ src, float* __restrict__ dst_f32,
                                    __nv_bfloat16* __restrict__ dst_bf16, long long n, int expneg2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i];
    if (dst_bf16) dst_bf16[i] = __float2bfloat16_rn(v);
    else dst_f32[i] = expneg2 ? exp_neg2(fmaxf(v, -1.0f)) : v;
  }
}

// fp16 -> fp32, 8 elements (one 16-byte load, two 16-byte stores) per thread and iteration
__global__ void expand_f16_kernel(const __half* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n8 = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&t);
    const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n8 * 8)) dst[n8 * 8 + threadIdx.x] = __half2float(src[n8 * 8 + threadIdx.x]);
}

// ------------------------------------------------------------------ bump allocator
struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// ------------------------------------------------------------------ packed weights
struct Mat {
  float* f32 = nullptr;
  bf16* hi = nullptr;
  bf16* lo = nullptr;
  int rows = 0, cols = 0;
  Operand op() const {
    Operand o;
    o.f32 = f32;
    o.ld = cols;
    o.hi = hi;
    o.lo = lo;
    o.ldp = cols;
    return o;
  }
};

struct Packed {
  Mat W1, Wpre, W2, W3, W4, W5, Wfc, Watt, Wa2a, Ws2a, Wcpt, Wl2w;
  // decode-only shortcut for the attention LSTM's word term: xt_gates[v] = W_ih[:, 2H:3H] . ReLU(E[v]) for every word
  // (tensor-core precisions). The gate GEMM then contracts over [h_lang_prev | h_att_prev] only (W1b, K = 2H) and its
  // fused LSTM epilogue adds the row xt_gates[it] — a third of that GEMM's flops becomes an 8 KB gather per row.
  Mat W1b;           // [4H, 2H] = W1 columns [h_lang_prev | h_att_prev]; planes GATE-INTERLEAVED (fused LSTM GEMM only)
  Mat W4g;           // gate-interleaved planes of W4 for the fused LSTM GEMM (no fp32 copy)
  float* xt_gates;   // [V, 4H]
  bf16 *erelu_hi, *erelu_lo;  // [V, H] planes of ReLU(E), operand of the GEMM that builds xt_gates
  float *b1, *b2, *b3, *b4, *b5, *bfc, *batt, *ba2a, *bs2a, *bcpt, *bl2w;
  float *emb, *lab_emb, *alpha_c, *alpha_s, *alpha_g, *alpha_g_b;
  size_t total = 0;
};

Packed carve_packed(const isc_dims_t& d, int precision, void* base) {
  Bump b(base);
  Packed p;
  const int V = d.vocab, D = d.feat_dim;
  auto mat = [&](Mat& m, int rows, int cols) {
    m.rows = rows;
    m.cols = cols;
    m.f32 = b.take<float>((size_t)rows * cols);
    if (precision != ISC_PREC_FP32) {
      m.hi = b.take<bf16>((size_t)rows * cols);
      if (precision == ISC_PREC_BF16X3) m.lo = b.take<bf16>((size_t)rows * cols);
    }
  };
  mat(p.W1, G4, 3 * H);
  mat(p.Wpre, G4, 2 * H);
  mat(p.W2, 3 * H, H);
  mat(p.W3, H, 2 * H);
  mat(p.W4, G4, 3 * H);
  mat(p.W5, V, H);
  p.xt_gates = nullptr;
  p.erelu_hi = p.erelu_lo = nullptr;
  if (precision != ISC_PREC_FP32) {
    mat(p.W1b, G4, 2 * H);
    p.W4g.rows = G4;
    p.W4g.cols = 3 * H;
    p.W4g.hi = b.take<bf16>((size_t)G4 * 3 * H);
    if (precision == ISC_PREC_BF16X3) p.W4g.lo = b.take<bf16>((size_t)G4 * 3 * H);
    p.xt_gates = b.take<float>((size_t)V * G4);
    p.erelu_hi = b.take<bf16>((size_t)V * H);
    if (precision == ISC_PREC_BF16X3) p.erelu_lo = b.take<bf16>((size_t)V * H);
  }
  mat(p.Wfc, H, D);
  mat(p.Watt, H, D);
  mat(p.Wa2a, H, H);
  mat(p.Ws2a, H, H);
  mat(p.Wcpt, H, H);
  mat(p.Wl2w, H, H);
  p.b1 = b.take<float>(G4);
  p.b2 = b.take<float>(3 * H);
  p.b3 = b.take<float>(H);
  p.b4 = b.take<float>(G4);
  p.b5 = b.take<float>(V);
  p.bfc = b.take<float>(H);
  p.batt = b.take<float>(H);
  p.ba2a = b.take<float>(H);
  p.bs2a = b.take<float>(H);
  p.bcpt = b.take<float>(H);
  p.bl2w = b.take<float>(H);
  p.emb = b.take<float>((size_t)V * H);
  p.lab_emb = b.take<float>((size_t)d.n_labels * H);
  p.alpha_c = b.take<float>(H);
  p.alpha_s = b.take<float>(H);
  p.alpha_g = b.take<float>(H);
  p.alpha_g_b = b.take<float>(4);
  p.total = (b.off + 255) & ~size_t(255);
  return p;
}

__global__ void add_vec_kernel(float* dst, const float* a, const float* b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a[i] + (b ? b[i] : 0.f);
}

int copy_block(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int rows, int cols, cudaStream_t s) {
  ISC_CUDA(cudaMemcpy2DAsync(dst, ld_dst * sizeof(float), src, ld_src * sizeof(float), (size_t)cols * sizeof(float), rows,
                             cudaMemcpyDeviceToDevice, s));
  return 0;
}
int add_vec(float* dst, const float* a, const float* b, int n, cudaStream_t s) {
  ProfScope ps(ISC_K_POINTWISE, 3.0 * n * sizeof(float), s);
  add_vec_kernel<<<(n + 255) / 256, 256, 0, s>>>(dst, a, b, n);
  ISC_LAUNCH_CHECK();
  return 0;
}
int finish_mat(const Mat& m, int precision, cudaStream_t s) {
  if (precision == ISC_PREC_FP32) return 0;
  return split_planes(m.f32, m.cols, m.hi, m.lo, m.cols, m.rows, m.cols, s);
}

// ------------------------------------------------------------------ decode workspace
struct Planes {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};
struct DecodeWs {
  float *X1, *X2, *gates, *gates2, *hproj, *cs, *g3, *logits;  // gates2 == gates unless a tape keeps both
  bool tape;  // training: the gate pre-activations must be kept, so the LSTM cell is not fused into the GEMM
  Planes pX1, pX2, pcs, phL;
  float* state_h[2];
  float* state_c[2];
  long long* it;
  int* parent;
  int* unfinished;
  int* alive_count;
  int* tok[2];
  int* len[2];
  double* score[2];
  int* alive[2];
  float* fcsl;
  Planes pfcsl;
  float* rec;  // LogitsSelect records of the fused logits epilogue [M][np][sel_rec(8)] (tensor-core precisions)
  int np;
  float* cand_lp;
  int* cand_word;
  int* cand_count;
  int* ticket;
  long long ld_logits;
  size_t total;
};

DecodeWs carve_decode(const isc_dims_t& d, int precision, int M, void* base) {
  Bump b(base);
  DecodeWs w;
  const bool tc = precision != ISC_PREC_FP32;
  const bool x3 = precision == ISC_PREC_BF16X3;
  const size_t m = (size_t)M;
  auto planes = [&](Planes& p, size_t n) {
    if (tc) {
      p.hi = b.take<bf16>(n);
      if (x3) p.lo = b.take<bf16>(n);
    }
  };
  w.X1 = tc ? nullptr : b.take<float>(m * 3 * H);
  w.X2 = tc ? nullptr : b.take<float>(m * 3 * H);
  planes(w.pX1, m * 3 * H);
  planes(w.pX2, m * 3 * H);
  w.gates = b.take<float>(m * G4);
  w.gates2 = w.gates;
  w.tape = false;
  w.hproj = b.take<float>(m * 3 * H);
  w.cs = b.take<float>(m * 2 * H);
  planes(w.pcs, m * 2 * H);
  w.g3 = b.take<float>(m * H);
  planes(w.phL, m * H);
  w.ld_logits = (d.vocab + 3) & ~3;
  w.logits = b.take<float>(m * w.ld_logits);
  for (int i = 0; i < 2; ++i) {
    w.state_h[i] = b.take<float>(2 * m * H);
    w.state_c[i] = b.take<float>(2 * m * H);
  }
  w.it = b.take<long long>(m);
  w.parent = b.take<int>(m);
  w.unfinished = b.take<int>(m);
  w.alive_count = b.take<int>(T_MAX);
  for (int i = 0; i < 2; ++i) {
    w.tok[i] = b.take<int>(m * T_MAX);
    w.len[i] = b.take<int>(m);
    w.score[i] = b.take<double>(m);
    w.alive[i] = b.take<int>(m);
  }
  w.fcsl = b.take<float>(m * 2 * H);
  planes(w.pfcsl, m * 2 * H);
  w.np = logits_slices(d.vocab);
  w.rec = tc ? b.take<float>(m * w.np * sel_rec(SEL_K_MAX)) : nullptr;
  w.cand_lp = b.take<float>(m * 8);
  w.cand_word = b.take<int>(m * 8);
  w.cand_count = b.take<int>(m);
  w.ticket = b.take<int>(m);
  w.total = (b.off + 255) & ~size_t(255);
  return w;
}

RowDest rowdest(float* f32, long long ld, const Planes& p, long long ldp) {
  RowDest r;
  r.f32 = f32;
  r.ld = ld;
  r.hi = p.hi;
  r.lo = p.lo;
  r.ldp = ldp;
  return r;
}
Operand operand(const float* f32, long long ld, const Planes& p, long long ldp, long long col = 0) {
  Operand o;
  o.f32 = f32 ? f32 + col : nullptr;
  o.ld = ld;
  o.hi = p.hi ? p.hi + col : nullptr;
  o.lo = p.lo ? p.lo + col : nullptr;
  o.ldp = ldp;
  return o;
}

struct Ctx {
  isc_dims_t d;
  Packed pk;
  int precision;
  const isc_feats_t* f;
  cudaStream_t s;
};

// hoisted step-invariant terms: pre_gates = [fc | sl] · Wpre^T + (b_ih + b_hh), pre_word = label2word(sl)
int run_hoist(const Ctx& c, int B, float* fcsl, const Planes& pfcsl) {
  const isc_feats_t& f = *c.f;
  ISC_REQUIRE(f.fc && f.pre_gates, "feats.fc and feats.pre_gates are required");
  const bool tc = c.precision != ISC_PREC_FP32;
  const int K = f.sl ? 2 * H : H;
  ISC_TRY(copy_block(fcsl, 2 * H, f.fc, H, B, H, c.s));
  if (f.sl) ISC_TRY(copy_block(fcsl + H, 2 * H, f.sl, H, B, H, c.s));
  if (tc) ISC_TRY(split_planes(fcsl, 2 * H, pfcsl.hi, pfcsl.lo, 2 * H, B, K, c.s));
  Epilogue ep;
  ep.bias = c.pk.b1;
  Dest dst;
  dst.f32 = f.pre_gates;
  dst.ld = G4;
  ISC_TRY(gemm(c.precision, operand(fcsl, 2 * H, pfcsl, 2 * H), c.pk.Wpre.op(), dst, B, G4, K, ep, c.s));
  if (f.sl && f.pre_word) {
    Epilogue e2;
    e2.bias = c.pk.bl2w;
    Dest d2;
    d2.f32 = f.pre_word;
    d2.ld = H;
    ISC_TRY(gemm(c.precision, operand(fcsl, 2 * H, pfcsl, 2 * H, H), c.pk.Wl2w.op(), d2, B, H, H, e2, c.s));
  }
  return 0;
}

struct StepIO {
  const long long* it;
  const int* parent;
  const float* h_in;
  const float* c_in;
  float* h_out;
  float* c_out;
  float* logits;
  long long ld_logits;
  float* cont_w = nullptr;
  long long ld_cont_w = 0;
  float* senti_w = nullptr;
  long long ld_senti_w = 0;
  float* gate_w = nullptr;
  long long ld_gate_w = 0;
  LogitsSelect sel;  // sel.rec != null: the classifier GEMM emits selection records instead of logits
  const unsigned char* out_mask = nullptr;  // dropout keep-mask [M,H] on h_lang before the classifier (captioner.py:182)
  float drop_scale = 1.0f;
  // rows of one plane of the OUTPUT state ([2][state_rows][H]: attention LSTM, language LSTM); 0 = M. Larger than M when
  // a beam search's first step runs on one row per image but the following steps index B * K rows.
  long long state_rows = 0;
  bool skip_pack = false;  // X1 / X2's recurrent rows were already written (by the previous step's beam merge)
};

// One decode step over M rows (captioner.py:168-186), raw classifier logits out.
int run_step(const Ctx& c, const DecodeWs& w, int M, int R, const StepIO& io) {
  const isc_feats_t& f = *c.f;
  const Packed& pk = c.pk;
  const int B = M / R;
  const bool has_att = f.att != nullptr, has_sw = f.sw != nullptr;
  ISC_REQUIRE(has_att || has_sw, "feats: att and sw are both NULL");
  ISC_REQUIRE(!has_att || f.p_att, "feats.p_att missing");
  ISC_REQUIRE(!has_sw || (f.p_sw && f.sl && f.pre_word), "feats.p_sw / sl / pre_word missing");
  const bool rl = has_att && has_sw;
  const long long m = M;
  const long long so = io.state_rows > 0 ? io.state_rows : m;  // rows per plane of the output state

  const bool fuse_lstm = c.precision != ISC_PREC_FP32 && !w.tape;  // LSTM cell inside the gate GEMM's epilogue
  const int passes = c.precision == ISC_PREC_BF16X3 ? 3 : 1;
  RowDest x1 = rowdest(w.X1, 3 * H, w.pX1, 3 * H);
  RowDest x2 = rowdest(w.X2, 3 * H, w.pX2, 3 * H);
  // fused path: X1 = [h_lang_prev | h_att_prev] only, the word term comes from the xt_gates table in the epilogue
  if (!io.skip_pack)
    ISC_TRY(launch_embed_pack(io.it, io.parent, io.h_in, M, c.d.vocab, fuse_lstm ? nullptr : pk.emb, x1, x2, c.s));

  // attention LSTM
  if (fuse_lstm) {
    LstmEpilogue le;
    le.parent = io.parent;
    le.c_prev = io.c_in;
    le.h_out = io.h_out;
    le.c_out = io.c_out;
    le.x_hi = w.pX2.hi;  // h_att is the middle third of the language LSTM's operand
    le.x_lo = w.pX2.lo;
    le.ldx = 3 * H;
    le.x_col = H;
    le.gather_tab = pk.xt_gates;
    le.gather_idx = io.it;
    le.gather_rows = c.d.vocab;
    ISC_TRY(gemm_tc_lstm(operand(nullptr, 0, w.pX1, 3 * H), pk.W1b.op(), M, 2 * H, passes, nullptr, f.pre_gates, G4, R, le, c.s));
  } else {
    Epilogue ep;
    ep.rowadd = f.pre_gates;
    ep.ld_rowadd = G4;
    ep.rows_per_group = R;
    Dest dst;
    dst.f32 = w.gates;
    dst.ld = G4;
    ISC_TRY(gemm(c.precision, operand(w.X1, 3 * H, w.pX1, 3 * H), pk.W1.op(), dst, M, G4, 3 * H, ep, c.s));
    ISC_TRY(launch_lstm_pointwise(w.gates, io.parent, io.c_in, io.h_out, io.c_out, x2, H, M, c.s));
  }
  // h projections: [cont h2att | senti h2word | gate h2att]
  {
    Epilogue ep;
    ep.bias = pk.b2;
    Dest dst;
    dst.f32 = w.hproj;
    dst.ld = 3 * H;
    Operand a = c.precision == ISC_PREC_FP32 ? operand(io.h_out, H, Planes(), H) : operand(nullptr, 0, w.pX2, 3 * H, H);
    ISC_TRY(gemm(c.precision, a, pk.W2.op(), dst, M, 3 * H, H, ep, c.s));
  }
  // attention
  {
    AttnParams ap;
    ap.R = R;
    ap.L = c.d.n_regions;
    ap.S = c.d.n_senti;
    ap.hproj = w.hproj;
    ap.ld_hproj = 3 * H;
    ap.att = f.att;
    ap.p_att = f.p_att;
    if (!w.tape && f.p_att16) {
      // 16-bit copies (decode loops); not with the training tape, whose backward recomputes from the fp32 tensors
      ap.att16 = f.att16;
      ap.p_att16 = f.p_att16;
      ap.flags = f.feat_flags;
    }
    ap.sw = f.sw;
    ap.p_sw = f.p_sw;
    ap.pre_word = f.pre_word;
    ap.alpha_c = pk.alpha_c;
    ap.alpha_s = pk.alpha_s;
    RowDest cs = rowdest(w.cs, 2 * H, w.pcs, 2 * H);
    if (rl) {
      ap.cont_dst = cs;
      ap.cont_col = 0;
      ap.senti_dst = cs;
      ap.senti_col = H;
    } else {  // xe: content only; seq2seq: sentiment only -> straight into the language-LSTM input
      ap.cont_dst = x2;
      ap.cont_col = 0;
      ap.senti_dst = x2;
      ap.senti_col = 0;
    }
    ap.cont_w = io.cont_w;
    ap.ld_cont_w = io.ld_cont_w;
    ap.senti_w = io.senti_w;
    ap.ld_senti_w = io.ld_senti_w;
    ISC_TRY(launch_attention(ap, B, c.precision, c.s));
  }
  if (rl) {
    // ISC_GATE_FUSED=1: gate GEMM with the gate weight and the context mix in its epilogue (a cluster of 4 CTAs per row
    // block exchanging their alpha . g3 partial sums through distributed shared memory; not when the tape needs g3 or the
    // GEMM runs on CUDA cores). Measured on the B = 1024 beam-3 call: the gate_mix launch (7 us) and the g3 round trip
    // disappear, the cluster barrier, the exchange and the mix inside the epilogue add 4-6 us to the GEMM, and cluster
    // launches start later: 7.47-7.49 ms unfused, 7.50-7.58 ms fused per call — off by default, results identical.
    static const bool gate_fused_on = getenv("ISC_GATE_FUSED") && atoi(getenv("ISC_GATE_FUSED")) != 0;
    if (gate_fused_on && c.precision != ISC_PREC_FP32 && !w.tape) {
      GateEpilogue ge;
      ge.alpha = pk.alpha_g;
      ge.alpha_b = pk.alpha_g_b;
      ge.cs = w.cs;
      ge.ld_cs = 2 * H;
      ge.gate_w = io.gate_w;
      ge.ld_gate_w = io.ld_gate_w;
      Dest dst;
      dst.f32 = x2.f32;
      dst.ld = x2.ld;
      dst.hi = x2.hi;
      dst.lo = x2.lo;
      dst.ldp = x2.ldp;
      ISC_TRY(gemm_tc_gate(operand(w.cs, 2 * H, w.pcs, 2 * H), pk.W3.op(), M, 2 * H, passes, pk.b3, w.hproj + 2 * H, 3 * H, ge, dst, c.s));
    } else {
      Epilogue ep;
      ep.bias = pk.b3;
      ep.addmat = w.hproj + 2 * H;
      ep.ld_addmat = 3 * H;
      ep.act = ACT_TANH;
      Dest dst;
      dst.f32 = w.g3;
      dst.ld = H;
      ISC_TRY(gemm(c.precision, operand(w.cs, 2 * H, w.pcs, 2 * H), pk.W3.op(), dst, M, H, 2 * H, ep, c.s));
      ISC_TRY(launch_gate_mix(w.g3, w.cs, pk.alpha_g, pk.alpha_g_b, x2, io.gate_w, io.ld_gate_w, M, c.s));
    }
  }

  // language LSTM
  if (fuse_lstm) {
    LstmEpilogue le;
    le.parent = io.parent;
    le.c_prev = io.c_in + m * H;
    le.h_out = io.h_out + so * H;
    le.c_out = io.c_out + so * H;
    le.x_hi = w.phL.hi;  // h_lang (after dropout, if any) is the classifier's operand
    le.x_lo = w.phL.lo;
    le.ldx = H;
    le.x_col = 0;
    le.mask = io.out_mask;
    le.scale = io.drop_scale;
    ISC_TRY(gemm_tc_lstm(operand(nullptr, 0, w.pX2, 3 * H), pk.W4g.op(), M, 3 * H, passes, pk.b4, nullptr, 0, 1, le, c.s));
  } else {
    Epilogue ep;
    ep.bias = pk.b4;
    Dest dst;
    dst.f32 = w.gates2;
    dst.ld = G4;
    ISC_TRY(gemm(c.precision, operand(w.X2, 3 * H, w.pX2, 3 * H), pk.W4.op(), dst, M, G4, 3 * H, ep, c.s));
    RowDest hl = rowdest(nullptr, 0, w.phL, H);
    ISC_TRY(launch_lstm_pointwise(w.gates2, io.parent, io.c_in + m * H, io.h_out + so * H, io.c_out + so * H, hl, 0, M, c.s,
                                  io.out_mask, io.drop_scale));
  }
  // classifier logits
  {
    Epilogue ep;
    ep.bias = pk.b5;
    Dest dst;
    dst.f32 = io.logits;
    dst.ld = io.ld_logits;
    Operand a = operand(io.h_out + so * H, H, w.phL, H);
    if (io.sel.rec) {
      ISC_TRY(gemm_tc_logits(a, pk.W5.op(), M, c.d.vocab, H, c.precision == ISC_PREC_BF16X3 ? 3 : 1, pk.b5, io.sel, c.s));
    } else {
      ISC_TRY(gemm(c.precision, a, pk.W5.op(), dst, M, c.d.vocab, H, ep, c.s));
    }
  }
  return 0;
}

int make_ctx(Ctx& c, const isc_dims_t* dims, const void* packed, int precision, const isc_feats_t* feats,
             isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_TRY(check_dims(dims));
  ISC_TRY(check_precision(precision));
  ISC_REQUIRE(packed != nullptr, "packed weights pointer is NULL");
  c.d = *dims;
  c.pk = carve_packed(*dims, precision, const_cast<void*>(packed));
  c.precision = precision;
  c.f = feats;
  c.s = static_cast<cudaStream_t>(stream);
  return 0;
}

// prologue workspace
struct ProWs {
  bf16 *raw_hi, *raw_lo;  // [chunk*L, D] raw region features as planes
  bf16 *att_hi, *att_lo;  // [chunk*L, H]
  bf16 *fc_hi, *fc_lo;    // [B, D]
  float* att_pre;         // [chunk*L, H] tiled batches: ReLU(att_embed) of a chunk's images, before expansion + dropout
  float* tmp;             // [B*S, H] scratch rows (cpt mean / sw)
  bf16 *tmp_hi, *tmp_lo;
  float* fcsl;
  Planes pfcsl;
  int chunk;
  size_t total;
};
ProWs carve_prologue(const isc_dims_t& d, int precision, int B, void* base) {
  Bump b(base);
  ProWs w;
  memset(&w, 0, sizeof(w));
  const bool tc = precision != ISC_PREC_FP32, x3 = precision == ISC_PREC_BF16X3;
  w.chunk = B < 96 ? B : 96;  // 96 x 196 rows = 147 row tiles x 2 wide column tiles = 2 waves on 148 SMs
  const size_t rows = (size_t)w.chunk * d.n_regions;
  if (tc) {
    w.raw_hi = b.take<bf16>(rows * d.feat_dim);
    if (x3) w.raw_lo = b.take<bf16>(rows * d.feat_dim);
    if (x3) {  // in ISC_PREC_BF16 the bf16 feature tensor itself is the next GEMM's operand
      w.att_hi = b.take<bf16>(rows * H);
      w.att_lo = b.take<bf16>(rows * H);
    }
    w.fc_hi = b.take<bf16>((size_t)B * d.feat_dim);
    if (x3) w.fc_lo = b.take<bf16>((size_t)B * d.feat_dim);
  }
  if (d.att_tile > 1) w.att_pre = b.take<float>(rows * H);
  const size_t trow = (size_t)B * (d.n_senti > 1 ? d.n_senti : 1);
  w.tmp = b.take<float>(trow * H);
  if (tc) {
    w.tmp_hi = b.take<bf16>(trow * H);
    if (x3) w.tmp_lo = b.take<bf16>(trow * H);
  }
  w.fcsl = b.take<float>((size_t)B * 2 * H);
  if (tc) {
    w.pfcsl.hi = b.take<bf16>((size_t)B * 2 * H);
    if (x3) w.pfcsl.lo = b.take<bf16>((size_t)B * 2 * H);
  }
  w.total = (b.off + 255) & ~size_t(255);
  return w;
}

int run_prologue(const isc_dims_t* dims, const void* packed, int precision, const float* fc_feats, const float* att_feats,
                 const int64_t* cpt_words, int n_cpt, const int64_t* senti_words, const int64_t* senti_labels, int B,
                 int seq2seq, const isc_feats_t* out, void* workspace, size_t workspace_bytes, isc_stream_t stream,
                 const isc_dropout_t* drop, float* fc_embedded, bool raw_bf16 = false);

}  // namespace
}  // namespace isc


namespace isc {
namespace {

// The prologue proper. drop != null applies the training-mode dropout masks (keep flags, captioner.py:200/210/214,
// :250/258, :296/304/311/315) right after each ReLU; fc_embedded receives the pre-dropout fc embedding
// (the reference's self.fc_feats, used by the domain-alignment loss).
int run_prologue(const isc_dims_t* dims, const void* packed, int precision, const float* fc_feats,
                 const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                 const int64_t* senti_labels, int B, int seq2seq, const isc_feats_t* out, void* workspace,
                 size_t workspace_bytes, isc_stream_t stream, const isc_dropout_t* drop, float* fc_embedded, bool raw_bf16) {
  // raw_bf16: fc_feats / att_feats point at bf16 values (a bf16 feature shard); ISC_PREC_BF16 only, where the fp32
  // inputs are rounded to exactly these numbers before the GEMMs anyway.
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, out, stream));
  ISC_REQUIRE(out != nullptr && B > 0, "out is NULL or B <= 0");
  ISC_REQUIRE(!raw_bf16 || (precision == ISC_PREC_BF16 && !drop && !seq2seq), "bf16 input features need ISC_PREC_BF16, eval mode");
  const float dscale = drop ? drop->scale : 1.0f;
  ProWs w = carve_prologue(*dims, precision, B, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("prologue workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  const Packed& pk = c.pk;
  const int D = dims->feat_dim, L = dims->n_regions, S = dims->n_senti, V = dims->vocab;
  const bool tc = precision != ISC_PREC_FP32;
  Planes ptmp;
  ptmp.hi = w.tmp_hi;
  ptmp.lo = w.tmp_lo;
  // concept branch (captioner.py:297-300): mean of ReLU(word_embed) -> cpt2fc -> ReLU
  float* cpt_dst = seq2seq ? out->fc : out->cpt_feats;
  if (cpt_words && cpt_dst) {
    ISC_TRY(launch_embed_mean(reinterpret_cast<const long long*>(cpt_words), B, n_cpt, V, pk.emb,
                              rowdest(w.tmp, H, ptmp, H), c.s));
    Epilogue ep;
    ep.bias = pk.bcpt;
    ep.act = ACT_RELU;
    Dest dst;
    dst.f32 = cpt_dst;
    dst.ld = H;
    ISC_TRY(gemm(precision, operand(w.tmp, H, ptmp, H), pk.Wcpt.op(), dst, B, H, H, ep, c.s));
    if (seq2seq && out->cpt_feats && out->cpt_feats != out->fc)
      ISC_TRY(copy_block(out->cpt_feats, H, out->fc, H, B, H, c.s));
    if (seq2seq && drop && drop->fc)  // captioner.py:250: dropout on cpt_feats, which then stands in for fc_feats
      ISC_TRY(launch_apply_mask(out->fc, H, drop->fc, dscale, B, H, RowDest(), c.s));
  } else {
    ISC_REQUIRE(!seq2seq, "seq2seq prologue needs cpt_words and out->fc");
  }
  if (!seq2seq) {
    ISC_REQUIRE(fc_feats && att_feats && out->fc && out->att && out->p_att, "fc/att inputs or outputs missing");
    // fc_embed (captioner.py:294)
    {
      Planes pfc;
      pfc.hi = w.fc_hi;
      pfc.lo = w.fc_lo;
      if (raw_bf16) {
        pfc.hi = reinterpret_cast<bf16*>(const_cast<float*>(fc_feats));
        pfc.lo = nullptr;
      } else if (tc) {
        ISC_TRY(split_planes(fc_feats, D, w.fc_hi, w.fc_lo, D, B, D, c.s));
      }
      Epilogue ep;
      ep.bias = pk.bfc;
      ep.act = ACT_RELU;
      Dest dst;
      dst.f32 = out->fc;
      dst.ld = H;
      ISC_TRY(gemm(precision, operand(raw_bf16 ? nullptr : fc_feats, D, pfc, D), pk.Wfc.op(), dst, B, H, D, ep, c.s));
      if (fc_embedded) ISC_TRY(copy_block(fc_embedded, H, out->fc, H, B, H, c.s));
      if (drop && drop->fc) ISC_TRY(launch_apply_mask(out->fc, H, drop->fc, dscale, B, H, RowDest(), c.s));
    }
    // att_embed + att2att (captioner.py:302-305), chunked over images so the operand planes stay small (a chunk's
    // planes fit the 126 MB L2). Running the fp32 -> bf16 plane split of chunk i+1 on a second stream beside the GEMMs
    // of chunk i was measured and gains nothing: both are bound by the same L2/HBM traffic.
    static const bool fused_split = getenv("ISC_PROLOGUE_SPLIT_KERNEL") == nullptr;  // set to use the separate split pass
    const int R = dims->att_tile > 1 ? dims->att_tile : 1;  // tiled batch: R rows per image, att_feats holds B / R images
    ISC_REQUIRE(R == 1 || (B % R == 0 && precision != ISC_PREC_BF16 && !raw_bf16 && w.att_pre && w.chunk >= R),
                "att_tile=%d needs B %% att_tile == 0 (B=%d) and fp32 features", R, B);
    const int chunk = R == 1 ? w.chunk : (w.chunk / R) * R;  // a chunk holds whole images' tiles
    for (int b0 = 0; b0 < B; b0 += chunk) {
      const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
      const long long rows = (long long)nb * L;           // rows of the (tiled) feature tensors in this chunk
      const long long rows_in = rows / R;                 // rows of the raw features behind them
      const float* raw = att_feats + (long long)(b0 / R) * L * D;
      Planes praw;
      praw.hi = w.raw_hi;
      praw.lo = w.raw_lo;
      Epilogue ep;
      ep.bias = pk.batt;
      ep.act = ACT_RELU;
      Dest dst;
      Planes patt;
      if (precision == ISC_PREC_BF16) {
        dst.hi = reinterpret_cast<bf16*>(out->att) + (long long)b0 * L * H;
        dst.ldp = H;
        patt.hi = dst.hi;
      } else {
        dst.f32 = reinterpret_cast<float*>(out->att) + (long long)b0 * L * H;
        dst.ld = H;
        if (tc) {
          dst.hi = w.att_hi;
          dst.lo = w.att_lo;
          dst.ldp = H;
          patt.hi = w.att_hi;
          patt.lo = w.att_lo;
        }
      }
      float* att_out = dst.f32;
      if (R > 1) {  // the GEMM writes the images' embedding once, the expansion below forms the tiled rows and their planes
        dst.f32 = w.att_pre;
        dst.hi = dst.lo = nullptr;
      }
      const bool want16 = tc && out->p_att16 != nullptr;
      if (want16) {
        ISC_REQUIRE(out->feat_flags && !drop && R == 1, "out->p_att16 needs out->feat_flags (and no dropout, no tiling)");
        if (b0 == 0) ISC_CUDA(cudaMemsetAsync(out->feat_flags, 0, (size_t)B * sizeof(int), c.s));
        if (out->att16 && precision == ISC_PREC_BF16X3) {  // fp16(att) beside the fp32 tensor
          dst.h16.out = static_cast<__half*>(out->att16) + (long long)b0 * L * H;
          dst.h16.ld = H;
          dst.h16.flags = out->feat_flags + b0;
          dst.h16.rows_per_flag = L;
        }
      }
      if (raw_bf16) {
        Planes pin;
        pin.hi = reinterpret_cast<bf16*>(const_cast<float*>(att_feats)) + (long long)b0 * L * D;
        pin.lo = nullptr;
        ISC_TRY(gemm_tc(operand(nullptr, 0, pin, D), pk.Watt.op(), dst, (int)rows, H, D, 1, ep, c.s));
      } else if (tc && fused_split) {
        // the fp32 region features go straight into the GEMM: its converter warps split them in shared memory
        ISC_TRY(gemm_tc_af32(raw, D, pk.Watt.op(), dst, (int)rows_in, H, D, precision == ISC_PREC_BF16X3 ? 3 : 1, ep, c.s));
      } else {
        if (tc) ISC_TRY(split_planes(raw, D, w.raw_hi, w.raw_lo, D, rows_in, D, c.s));
        ISC_TRY(gemm(precision, operand(raw, D, praw, D), pk.Watt.op(), dst, (int)rows_in, H, D, ep, c.s));
      }
      if (R > 1) {
        ISC_TRY(launch_expand_mask(w.att_pre, R, (long long)L * H, nb, (drop && drop->att) ? drop->att + (long long)b0 * L * H : nullptr,
                                   dscale, att_out, H, rowdest(nullptr, 0, patt, H), c.s));
        dst.f32 = att_out;
      } else if (drop && drop->att) {
        ISC_REQUIRE(dst.f32 != nullptr, "dropout needs fp32 features (ISC_PREC_FP32 / ISC_PREC_BF16X3)");
        ISC_TRY(launch_apply_mask(dst.f32, H, drop->att + (long long)b0 * L * H, dscale, rows, H, rowdest(nullptr, 0, patt, H),
                                  c.s));
      }
      Epilogue ep2;
      ep2.bias = pk.ba2a;
      ep2.act = proj_act(precision);
      Dest d2;
      if (precision == ISC_PREC_BF16) {
        d2.hi = reinterpret_cast<bf16*>(out->p_att) + (long long)b0 * L * H;
        d2.ldp = H;
      } else {
        d2.f32 = reinterpret_cast<float*>(out->p_att) + (long long)b0 * L * H;
        d2.ld = H;
      }
      if (want16) {  // fp16(exp(-2 p) * 2^15) beside the full-width projected features
        d2.h16.out = static_cast<__half*>(out->p_att16) + (long long)b0 * L * H;
        d2.h16.ld = H;
        d2.h16.scale = kFastScale;
        d2.h16.expneg2 = precision == ISC_PREC_BF16 ? 1 : 0;  // ISC_PREC_BF16 keeps p itself in its bf16 tensor
        d2.h16.vmin = kFastMinStored;
        d2.h16.flags = out->feat_flags + b0;
        d2.h16.rows_per_flag = L;
      }
      ISC_TRY(gemm(precision, operand(dst.f32, H, patt, H), pk.Wa2a.op(), d2, (int)rows, H, H, ep2, c.s));
    }
  }
  // sentiment words with the PAD prepended (captioner.py:307-312)
  if (senti_words) {
    ISC_REQUIRE(out->sw && out->p_sw, "out->sw / out->p_sw missing");
    ISC_TRY(launch_embed_rows(reinterpret_cast<const long long*>(senti_words), B, S - 1, 1, dims->pad_id, V, pk.emb,
                              rowdest(out->sw, H, ptmp, H), c.s));
    if (drop && drop->sw) ISC_TRY(launch_apply_mask(out->sw, H, drop->sw, dscale, (long long)B * S, H, rowdest(nullptr, 0, ptmp, H), c.s));
    Epilogue ep;
    ep.bias = pk.bs2a;
    ep.act = proj_act(precision);
    Dest dst;
    dst.f32 = out->p_sw;
    dst.ld = H;
    ISC_TRY(gemm(precision, operand(out->sw, H, ptmp, H), pk.Ws2a.op(), dst, B * S, H, H, ep, c.s));
  }
  if (senti_labels) {
    ISC_REQUIRE(out->sl, "out->sl missing");
    ISC_TRY(launch_embed_rows(reinterpret_cast<const long long*>(senti_labels), B, 1, 0, 0, dims->n_labels, pk.lab_emb,
                              rowdest(out->sl, H, Planes(), H), c.s));
    if (drop && drop->sl) ISC_TRY(launch_apply_mask(out->sl, H, drop->sl, dscale, B, H, RowDest(), c.s));
  }
  // hoisted terms
  isc_feats_t f = *out;
  if (!senti_labels) f.sl = nullptr;
  Ctx c2 = c;
  c2.f = &f;
  ISC_TRY(run_hoist(c2, B, w.fcsl, w.pfcsl));
  return 0;
}

}  // namespace
}  // namespace isc
