// C-ABI entry points (include/isc.h) and the host-side launch sequences of the decode path.
// No device allocation, no synchronisation: everything is enqueued on the caller's stream into
// caller-owned buffers, so a whole decode call can be captured in a CUDA graph.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace isc {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------ profiler
namespace prof {
struct Rec {
  int kclass;
  double work;
  cudaEvent_t e0, e1;
};
static std::mutex mu;
static std::atomic<unsigned long long> launches{0};
static std::atomic<int> enabled{0};
static std::vector<Rec> recs;             // records of the current measurement window
static std::vector<cudaEvent_t> pool;     // recycled events
static cudaEvent_t get_event() {
  if (!pool.empty()) {
    cudaEvent_t e = pool.back();
    pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace prof

ProfScope::ProfScope(int kclass, double work, cudaStream_t stream) : slot_(-1), stream_(stream) {
  prof::launches.fetch_add(1, std::memory_order_relaxed);
  if (!prof::enabled.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(prof::mu);
  prof::Rec r;
  r.kclass = kclass;
  r.work = work;
  r.e0 = prof::get_event();
  r.e1 = prof::get_event();
  cudaEventRecord(r.e0, stream);
  prof::recs.push_back(r);
  slot_ = (int)prof::recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(prof::mu);
  if (slot_ < (int)prof::recs.size()) cudaEventRecord(prof::recs[slot_].e1, stream_);
}

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

  RowDest x1 = rowdest(w.X1, 3 * H, w.pX1, 3 * H);
  RowDest x2 = rowdest(w.X2, 3 * H, w.pX2, 3 * H);
  ISC_TRY(launch_embed_pack(io.it, io.parent, io.h_in, M, c.d.vocab, pk.emb, x1, x2, c.s));

  const bool fuse_lstm = c.precision != ISC_PREC_FP32 && !w.tape;  // LSTM cell inside the gate GEMM's epilogue
  const int passes = c.precision == ISC_PREC_BF16X3 ? 3 : 1;
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
    ISC_TRY(gemm_tc_lstm(operand(nullptr, 0, w.pX1, 3 * H), pk.W1.op(), M, 3 * H, passes, nullptr, f.pre_gates, G4, R, le, c.s));
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
    ISC_TRY(launch_attention(ap, B, c.precision == ISC_PREC_BF16, c.precision, c.s));  // tanh mode == precision id
  }
  if (rl) {
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
  // language LSTM
  if (fuse_lstm) {
    LstmEpilogue le;
    le.parent = io.parent;
    le.c_prev = io.c_in + m * H;
    le.h_out = io.h_out + m * H;
    le.c_out = io.c_out + m * H;
    le.x_hi = w.phL.hi;  // h_lang (after dropout, if any) is the classifier's operand
    le.x_lo = w.phL.lo;
    le.ldx = H;
    le.x_col = 0;
    le.mask = io.out_mask;
    le.scale = io.drop_scale;
    ISC_TRY(gemm_tc_lstm(operand(nullptr, 0, w.pX2, 3 * H), pk.W4.op(), M, 3 * H, passes, pk.b4, nullptr, 0, 1, le, c.s));
  } else {
    Epilogue ep;
    ep.bias = pk.b4;
    Dest dst;
    dst.f32 = w.gates2;
    dst.ld = G4;
    ISC_TRY(gemm(c.precision, operand(w.X2, 3 * H, w.pX2, 3 * H), pk.W4.op(), dst, M, G4, 3 * H, ep, c.s));
    RowDest hl = rowdest(nullptr, 0, w.phL, H);
    ISC_TRY(launch_lstm_pointwise(w.gates2, io.parent, io.c_in + m * H, io.h_out + m * H, io.c_out + m * H, hl, 0, M, c.s,
                                  io.out_mask, io.drop_scale));
  }
  // classifier logits
  {
    Epilogue ep;
    ep.bias = pk.b5;
    Dest dst;
    dst.f32 = io.logits;
    dst.ld = io.ld_logits;
    Operand a = operand(io.h_out + m * H, H, w.phL, H);
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
                 const isc_dropout_t* drop, float* fc_embedded);

}  // namespace
}  // namespace isc

using namespace isc;

extern "C" {

const char* isc_version(void) { return "insenticap-b200 0.1 (sm_100a)"; }

uint64_t isc_launch_count(void) { return prof::launches.load(); }
int isc_profile_enable(int on) {
  prof::enabled.store(on ? 1 : 0);
  return 0;
}
int isc_profile_reset(void) {
  std::lock_guard<std::mutex> lk(prof::mu);
  for (auto& r : prof::recs) {
    cudaEventSynchronize(r.e1);
    prof::pool.push_back(r.e0);
    prof::pool.push_back(r.e1);
  }
  prof::recs.clear();
  return 0;
}
int isc_profile_read(int kernel_class, double* total_ms, double* total_work, int64_t* launches) {
  ISC_REQUIRE(kernel_class >= 0 && kernel_class < ISC_K_NUM && total_ms && total_work && launches, "bad profile_read args");
  std::lock_guard<std::mutex> lk(prof::mu);
  double ms = 0.0, work = 0.0;
  long long n = 0;
  for (auto& r : prof::recs) {
    if (r.kclass != kernel_class) continue;
    ISC_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    ISC_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t;
    work += r.work;
    ++n;
  }
  *total_ms = ms;
  *total_work = work;
  *launches = n;
  return 0;
}
const char* isc_last_error(void) { return g_err; }
int isc_check_device(void) { return check_device(); }

size_t isc_packed_weights_bytes(const isc_dims_t* dims, int precision) {
  if (check_dims(dims) != 0 || check_precision(precision) != 0) return 0;
  return carve_packed(*dims, precision, nullptr).total;
}

int isc_pack_weights(const isc_dims_t* dims, const isc_weights_t* w, int precision, void* packed, size_t packed_bytes,
                     isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_TRY(check_dims(dims));
  ISC_TRY(check_precision(precision));
  ISC_REQUIRE(w && packed, "weights / packed pointer is NULL");
  Packed p = carve_packed(*dims, precision, packed);
  if (packed_bytes < p.total) {
    set_error("packed buffer too small: %zu < %zu", packed_bytes, p.total);
    return ISC_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int V = dims->vocab, D = dims->feat_dim;
  // attention LSTM: reference input order is [h_lang | fc | xt] (captioner.py:174); the fc slice is
  // step-invariant and moves to Wpre together with the xt slice applied to the sentiment label.
  ISC_TRY(copy_block(p.W1.f32, 3 * H, w->att_lstm_w_ih, 3 * H, G4, H, s));                  // h_lang_prev
  ISC_TRY(copy_block(p.W1.f32 + H, 3 * H, w->att_lstm_w_ih + 2 * H, 3 * H, G4, H, s));      // xt
  ISC_TRY(copy_block(p.W1.f32 + 2 * H, 3 * H, w->att_lstm_w_hh, H, G4, H, s));              // h_att_prev
  ISC_TRY(copy_block(p.Wpre.f32, 2 * H, w->att_lstm_w_ih + H, 3 * H, G4, H, s));            // fc
  ISC_TRY(copy_block(p.Wpre.f32 + H, 2 * H, w->att_lstm_w_ih + 2 * H, 3 * H, G4, H, s));    // sl (added to xt)
  ISC_TRY(add_vec(p.b1, w->att_lstm_b_ih, w->att_lstm_b_hh, G4, s));
  // the three projections of h_att
  ISC_TRY(copy_block(p.W2.f32, H, w->ca_h2att_w, H, H, H, s));
  ISC_TRY(copy_block(p.W2.f32 + (size_t)H * H, H, w->sa_h2word_w, H, H, H, s));
  ISC_TRY(copy_block(p.W2.f32 + (size_t)2 * H * H, H, w->g_h2att_w, H, H, H, s));
  ISC_TRY(add_vec(p.b2, w->ca_h2att_b, nullptr, H, s));
  ISC_TRY(add_vec(p.b2 + H, w->sa_h2word_b, nullptr, H, s));
  ISC_TRY(add_vec(p.b2 + 2 * H, w->g_h2att_b, nullptr, H, s));
  // gate: cont2att(c) + senti2att(s)
  ISC_TRY(copy_block(p.W3.f32, 2 * H, w->g_cont2att_w, H, H, H, s));
  ISC_TRY(copy_block(p.W3.f32 + H, 2 * H, w->g_senti2att_w, H, H, H, s));
  ISC_TRY(add_vec(p.b3, w->g_cont2att_b, w->g_senti2att_b, H, s));
  // language LSTM: [att | h_att] then h_lang_prev
  ISC_TRY(copy_block(p.W4.f32, 3 * H, w->lang_lstm_w_ih, 2 * H, G4, 2 * H, s));
  ISC_TRY(copy_block(p.W4.f32 + 2 * H, 3 * H, w->lang_lstm_w_hh, H, G4, H, s));
  ISC_TRY(add_vec(p.b4, w->lang_lstm_b_ih, w->lang_lstm_b_hh, G4, s));
  ISC_TRY(copy_block(p.W5.f32, H, w->classifier_w, H, V, H, s));
  ISC_TRY(add_vec(p.b5, w->classifier_b, nullptr, V, s));
  ISC_TRY(copy_block(p.Wfc.f32, D, w->fc_embed_w, D, H, D, s));
  ISC_TRY(copy_block(p.Watt.f32, D, w->att_embed_w, D, H, D, s));
  ISC_TRY(copy_block(p.Wa2a.f32, H, w->att2att_w, H, H, H, s));
  ISC_TRY(copy_block(p.Ws2a.f32, H, w->senti2att_w, H, H, H, s));
  ISC_TRY(copy_block(p.Wcpt.f32, H, w->cpt2fc_w, H, H, H, s));
  ISC_TRY(copy_block(p.Wl2w.f32, H, w->sa_label2word_w, H, H, H, s));
  ISC_TRY(add_vec(p.bfc, w->fc_embed_b, nullptr, H, s));
  ISC_TRY(add_vec(p.batt, w->att_embed_b, nullptr, H, s));
  ISC_TRY(add_vec(p.ba2a, w->att2att_b, nullptr, H, s));
  ISC_TRY(add_vec(p.bs2a, w->senti2att_b, nullptr, H, s));
  ISC_TRY(add_vec(p.bcpt, w->cpt2fc_b, nullptr, H, s));
  ISC_TRY(add_vec(p.bl2w, w->sa_label2word_b, nullptr, H, s));
  ISC_TRY(copy_block(p.emb, H, w->word_embed, H, V, H, s));
  ISC_TRY(copy_block(p.lab_emb, H, w->senti_label_embed, H, dims->n_labels, H, s));
  ISC_TRY(add_vec(p.alpha_c, w->ca_alpha_w, nullptr, H, s));
  ISC_TRY(add_vec(p.alpha_s, w->sa_alpha_w, nullptr, H, s));
  ISC_TRY(add_vec(p.alpha_g, w->g_alpha_w, nullptr, H, s));
  ISC_TRY(add_vec(p.alpha_g_b, w->g_alpha_b, nullptr, 1, s));
  const Mat* mats[] = {&p.W1, &p.Wpre, &p.W2, &p.W3, &p.W4, &p.W5, &p.Wfc, &p.Watt, &p.Wa2a, &p.Ws2a, &p.Wcpt, &p.Wl2w};
  for (const Mat* m : mats) ISC_TRY(finish_mat(*m, precision, s));
  return 0;
}

size_t isc_prologue_workspace_bytes(const isc_dims_t* dims, int precision, int B) {
  if (check_dims(dims) != 0 || check_precision(precision) != 0 || B <= 0) return 0;
  return carve_prologue(*dims, precision, B, nullptr).total;
}

int isc_prologue(const isc_dims_t* dims, const void* packed, int precision, const float* fc_feats,
                 const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                 const int64_t* senti_labels, int B, int seq2seq, const isc_feats_t* out, void* workspace,
                 size_t workspace_bytes, isc_stream_t stream, const isc_dropout_t* dropout) {
  return run_prologue(dims, packed, precision, fc_feats, att_feats, cpt_words, n_cpt, senti_words, senti_labels, B, seq2seq,
                      out, workspace, workspace_bytes, stream, dropout, nullptr);
}

}  // extern "C"

namespace isc {
namespace {

// The prologue proper. drop != null applies the training-mode dropout masks (keep flags, captioner.py:200/210/214,
// :250/258, :296/304/311/315) right after each ReLU; fc_embedded receives the pre-dropout fc embedding
// (the reference's self.fc_feats, used by the domain-alignment loss).
int run_prologue(const isc_dims_t* dims, const void* packed, int precision, const float* fc_feats,
                 const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                 const int64_t* senti_labels, int B, int seq2seq, const isc_feats_t* out, void* workspace,
                 size_t workspace_bytes, isc_stream_t stream, const isc_dropout_t* drop, float* fc_embedded) {
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, out, stream));
  ISC_REQUIRE(out != nullptr && B > 0, "out is NULL or B <= 0");
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
      if (tc) ISC_TRY(split_planes(fc_feats, D, w.fc_hi, w.fc_lo, D, B, D, c.s));
      Epilogue ep;
      ep.bias = pk.bfc;
      ep.act = ACT_RELU;
      Dest dst;
      dst.f32 = out->fc;
      dst.ld = H;
      ISC_TRY(gemm(precision, operand(fc_feats, D, pfc, D), pk.Wfc.op(), dst, B, H, D, ep, c.s));
      if (fc_embedded) ISC_TRY(copy_block(fc_embedded, H, out->fc, H, B, H, c.s));
      if (drop && drop->fc) ISC_TRY(launch_apply_mask(out->fc, H, drop->fc, dscale, B, H, RowDest(), c.s));
    }
    // att_embed + att2att (captioner.py:302-305), chunked over images so the operand planes stay small (a chunk's
    // planes fit the 126 MB L2). Running the fp32 -> bf16 plane split of chunk i+1 on a second stream beside the GEMMs
    // of chunk i was measured and gains nothing: both are bound by the same L2/HBM traffic.
    static const bool fused_split = getenv("ISC_PROLOGUE_SPLIT_KERNEL") == nullptr;  // set to use the separate split pass
    for (int b0 = 0; b0 < B; b0 += w.chunk) {
      const int nb = (B - b0 < w.chunk) ? (B - b0) : w.chunk;
      const long long rows = (long long)nb * L;
      const float* raw = att_feats + (long long)b0 * L * D;
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
      if (tc && fused_split) {
        // the fp32 region features go straight into the GEMM: its converter warps split them in shared memory
        ISC_TRY(gemm_tc_af32(raw, D, pk.Watt.op(), dst, (int)rows, H, D, precision == ISC_PREC_BF16X3 ? 3 : 1, ep, c.s));
      } else {
        if (tc) ISC_TRY(split_planes(raw, D, w.raw_hi, w.raw_lo, D, rows, D, c.s));
        ISC_TRY(gemm(precision, operand(raw, D, praw, D), pk.Watt.op(), dst, (int)rows, H, D, ep, c.s));
      }
      if (drop && drop->att) {
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

extern "C" {

size_t isc_decode_workspace_bytes(const isc_dims_t* dims, int precision, int M) {
  if (check_dims(dims) != 0 || check_precision(precision) != 0 || M <= 0) return 0;
  return carve_decode(*dims, precision, M, nullptr).total;
}

int isc_hoist(const isc_dims_t* dims, const void* packed, int precision, int B, const isc_feats_t* feats,
              void* workspace, size_t workspace_bytes, isc_stream_t stream) {
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, feats, stream));
  ISC_REQUIRE(feats && B > 0, "feats is NULL or B <= 0");
  DecodeWs w = carve_decode(*dims, precision, B, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("decode workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  return run_hoist(c, B, w.fcsl, w.pfcsl);
}

int isc_decode_step(const isc_dims_t* dims, const void* packed, int precision, const isc_feats_t* feats,
                    int rows_per_image, int M, const int64_t* it, const float* h_in, const float* c_in, float* h_out,
                    float* c_out, float* logprobs, int64_t ld_logprobs, float* cont_w, float* senti_w, float* gate_w,
                    void* workspace, size_t workspace_bytes, isc_stream_t stream) {
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, feats, stream));
  ISC_REQUIRE(feats && M > 0 && rows_per_image >= 1 && M % rows_per_image == 0, "bad M / rows_per_image");
  ISC_REQUIRE(it && h_in && c_in && h_out && c_out && logprobs, "NULL step buffer");
  ISC_REQUIRE(h_in != h_out && c_in != c_out, "step state must not alias");
  ISC_REQUIRE(ld_logprobs >= dims->vocab, "ld_logprobs < vocab");
  DecodeWs w = carve_decode(*dims, precision, M, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("decode workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  StepIO io;
  io.it = reinterpret_cast<const long long*>(it);
  io.parent = nullptr;
  io.h_in = h_in;
  io.c_in = c_in;
  io.h_out = h_out;
  io.c_out = c_out;
  io.logits = logprobs;
  io.ld_logits = ld_logprobs;
  io.cont_w = cont_w;
  io.ld_cont_w = dims->n_regions;
  io.senti_w = senti_w;
  io.ld_senti_w = dims->n_senti;
  io.gate_w = gate_w;
  io.ld_gate_w = 1;
  ISC_TRY(run_step(c, w, M, rows_per_image, io));
  return launch_log_softmax(logprobs, ld_logprobs, M, dims->vocab, c.s);
}

int isc_decode_greedy(const isc_dims_t* dims, const void* packed, int precision, const isc_feats_t* feats, int B, int T,
                      int sample_mode, const float* noise, uint64_t seed, int64_t* seq, float* seq_logprobs,
                      float* seq_masks, float* cont_w, float* senti_w, float* gate_w, void* workspace,
                      size_t workspace_bytes, isc_stream_t stream, const uint8_t* out_mask, float drop_scale) {
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, feats, stream));
  ISC_REQUIRE(feats && B > 0 && T > 0 && T <= T_MAX, "bad B or T (T <= %d)", T_MAX);
  ISC_REQUIRE(seq && seq_logprobs && seq_masks, "NULL output");
  ISC_REQUIRE(sample_mode >= 0 && sample_mode <= 2 && (sample_mode != 1 || noise), "bad sample_mode / noise");
  DecodeWs w = carve_decode(*dims, precision, B, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("decode workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  const size_t st = (size_t)2 * B * H * sizeof(float);
  ISC_CUDA(cudaMemsetAsync(w.state_h[0], 0, st, c.s));
  ISC_CUDA(cudaMemsetAsync(w.state_c[0], 0, st, c.s));
  ISC_CUDA(cudaMemsetAsync(seq, 0, (size_t)B * T * sizeof(int64_t), c.s));
  ISC_CUDA(cudaMemsetAsync(seq_logprobs, 0, (size_t)B * T * sizeof(float), c.s));
  ISC_CUDA(cudaMemsetAsync(seq_masks, 0, (size_t)B * T * sizeof(float), c.s));
  ISC_CUDA(cudaMemsetAsync(w.alive_count, 0, T_MAX * sizeof(int), c.s));
  ISC_TRY(launch_greedy_init(w.it, w.unfinished, B, dims->sos_id, c.s));
  const int L = dims->n_regions, S = dims->n_senti, V = dims->vocab;
  const bool fused = precision != ISC_PREC_FP32 && sample_mode == 0;  // argmax straight from the GEMM epilogue
  for (int t = 0; t < T; ++t) {
    StepIO io;
    if (fused) {
      io.sel.rec = w.rec;
      io.sel.np = w.np;
    }
    io.it = w.it;
    io.parent = nullptr;
    io.h_in = w.state_h[t & 1];
    io.c_in = w.state_c[t & 1];
    io.h_out = w.state_h[(t + 1) & 1];
    io.c_out = w.state_c[(t + 1) & 1];
    io.logits = w.logits;
    io.ld_logits = w.ld_logits;
    if (cont_w) {
      io.cont_w = cont_w + (long long)t * L;
      io.ld_cont_w = (long long)T * L;
    }
    if (senti_w) {
      io.senti_w = senti_w + (long long)t * S;
      io.ld_senti_w = (long long)T * S;
    }
    if (gate_w) {
      io.gate_w = gate_w + t;
      io.ld_gate_w = T;
    }
    if (out_mask) {
      io.out_mask = out_mask + (size_t)t * B * H;
      io.drop_scale = drop_scale;
    }
    ISC_TRY(run_step(c, w, B, 1, io));
    GreedyParams g;
    g.logits = fused ? nullptr : w.logits;
    g.ld = w.ld_logits;
    g.rec = io.sel.rec;
    g.np = io.sel.np;
    g.B = B;
    g.V = V;
    g.T = T;
    g.t = t;
    g.sample_mode = sample_mode;
    g.noise = noise ? noise + (long long)t * B * V : nullptr;
    g.seed = seed;
    g.eos_id = dims->eos_id;
    g.it = w.it;
    g.unfinished = w.unfinished;
    g.alive_count = w.alive_count;
    g.seq = reinterpret_cast<long long*>(seq);
    g.seq_logprobs = seq_logprobs;
    g.seq_masks = seq_masks;
    ISC_TRY(launch_greedy_select(g, c.s));
  }
  return 0;
}

int isc_decode_beam(const isc_dims_t* dims, const void* packed, int precision, const isc_feats_t* feats, int B, int K,
                    int T, int decoding_constraint, int64_t* tokens, double* scores, int32_t* lengths, void* workspace,
                    size_t workspace_bytes, isc_stream_t stream) {
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, feats, stream));
  ISC_REQUIRE(feats && B > 0 && T > 0 && T <= T_MAX && K >= 1 && K <= 8, "bad B / K (1..8) / T (<= %d)", T_MAX);
  ISC_REQUIRE(tokens && scores && lengths, "NULL output");
  ISC_REQUIRE(dims->vocab > K + 4, "vocab too small for beam size");
  const int M = B * K;
  DecodeWs w = carve_decode(*dims, precision, M, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("decode workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  const size_t st = (size_t)2 * M * H * sizeof(float);
  ISC_CUDA(cudaMemsetAsync(w.state_h[0], 0, st, c.s));
  ISC_CUDA(cudaMemsetAsync(w.state_c[0], 0, st, c.s));
  ISC_CUDA(cudaMemsetAsync(w.tok[0], 0, (size_t)M * T * sizeof(int), c.s));
  ISC_CUDA(cudaMemsetAsync(w.ticket, 0, (size_t)B * sizeof(int), c.s));
  ISC_TRY(launch_beam_init(w.it, w.alive[0], w.len[0], w.score[0], w.parent, B, K, dims->sos_id, c.s));
  const bool fused = precision != ISC_PREC_FP32;  // masks + top-K inside the logits GEMM epilogue
  const int k_sel = K <= 4 ? 4 : 8;               // candidates kept per 128-column slice
  for (int t = 0; t < T; ++t) {
    StepIO io;
    if (fused) {
      io.sel.rec = w.rec;
      io.sel.np = w.np;
      io.sel.k_sel = k_sel;
      io.sel.last = w.it;
      io.sel.constraint = decoding_constraint ? 1 : 0;
      io.sel.mask_special = dims->pad_id != dims->eos_id;
      io.sel.pad_id = dims->pad_id;
      io.sel.sos_id = dims->sos_id;
      io.sel.unk_id = dims->unk_id;
    }
    io.it = w.it;
    io.parent = t > 0 ? w.parent : nullptr;
    io.h_in = w.state_h[t & 1];
    io.c_in = w.state_c[t & 1];
    io.h_out = w.state_h[(t + 1) & 1];
    io.c_out = w.state_c[(t + 1) & 1];
    io.logits = w.logits;
    io.ld_logits = w.ld_logits;
    ISC_TRY(run_step(c, w, M, K, io));
    BeamParams bp;
    bp.logits = fused ? nullptr : w.logits;
    bp.ld = w.ld_logits;
    bp.rec = io.sel.rec;
    bp.np = io.sel.np;
    bp.k_sel = k_sel;
    bp.B = B;
    bp.K = K;
    bp.V = dims->vocab;
    bp.T = T;
    bp.t = t;
    bp.constraint = decoding_constraint;
    bp.pad_id = dims->pad_id;
    bp.sos_id = dims->sos_id;
    bp.eos_id = dims->eos_id;
    bp.unk_id = dims->unk_id;
    bp.tok_in = w.tok[t & 1];
    bp.tok_out = w.tok[(t + 1) & 1];
    bp.len_in = w.len[t & 1];
    bp.len_out = w.len[(t + 1) & 1];
    bp.score_in = w.score[t & 1];
    bp.score_out = w.score[(t + 1) & 1];
    bp.alive_in = w.alive[t & 1];
    bp.alive_out = w.alive[(t + 1) & 1];
    bp.it = w.it;
    bp.parent = w.parent;
    bp.cand_lp = w.cand_lp;
    bp.cand_word = w.cand_word;
    bp.cand_count = w.cand_count;
    bp.ticket = w.ticket;
    ISC_TRY(launch_beam_select(bp, c.s));
  }
  return launch_beam_finalize(w.tok[T & 1], w.len[T & 1], w.score[T & 1], reinterpret_cast<long long*>(tokens), scores,
                              lengths, B, K, T, c.s);
}

int isc_teacher_forced(const isc_dims_t* dims, const void* packed, int precision, const isc_feats_t* feats, int B,
                       int n_steps, const int64_t* inputs, int64_t ld_inputs, float* logprobs, void* workspace,
                       size_t workspace_bytes, isc_stream_t stream) {
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, feats, stream));
  ISC_REQUIRE(feats && B > 0 && n_steps > 0 && inputs && logprobs && ld_inputs >= n_steps, "bad teacher-forcing args");
  DecodeWs w = carve_decode(*dims, precision, B, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("decode workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  const size_t st = (size_t)2 * B * H * sizeof(float);
  ISC_CUDA(cudaMemsetAsync(w.state_h[0], 0, st, c.s));
  ISC_CUDA(cudaMemsetAsync(w.state_c[0], 0, st, c.s));
  const long long V = dims->vocab;
  for (int t = 0; t < n_steps; ++t) {
    ISC_CUDA(cudaMemcpy2DAsync(w.it, sizeof(long long), inputs + t, ld_inputs * sizeof(long long), sizeof(long long), B,
                               cudaMemcpyDeviceToDevice, c.s));
    StepIO io;
    io.it = w.it;
    io.parent = nullptr;
    io.h_in = w.state_h[t & 1];
    io.c_in = w.state_c[t & 1];
    io.h_out = w.state_h[(t + 1) & 1];
    io.c_out = w.state_c[(t + 1) & 1];
    io.logits = logprobs + (long long)t * V;
    io.ld_logits = (long long)n_steps * V;
    ISC_TRY(run_step(c, w, B, 1, io));
    ISC_TRY(launch_log_softmax(io.logits, io.ld_logits, B, (int)V, c.s));
  }
  return 0;
}

int isc_convert_features(int precision, int projected, const float* src, void* dst, int64_t n, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_TRY(check_precision(precision));
  ISC_REQUIRE(src && dst && n >= 0, "bad convert_features arguments");
  if (n == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const bool bf = precision == ISC_PREC_BF16;
  ProfScope ps(ISC_K_POINTWISE, (double)n * (bf ? 6.0 : 8.0), s);
  proj_convert_kernel<<<blocks, 256, 0, s>>>(src, bf ? nullptr : static_cast<float*>(dst),
                                             bf ? static_cast<__nv_bfloat16*>(dst) : nullptr, n,
                                             projected && precision == ISC_PREC_BF16X3);
  ISC_LAUNCH_CHECK();
  return 0;
}

size_t isc_gemm_workspace_bytes(int precision, int M, int N, int K) {
  if (precision == ISC_PREC_FP32) return 256;
  size_t planes = precision == ISC_PREC_BF16X3 ? 2 : 1;
  return planes * ((size_t)M * K + (size_t)N * K) * sizeof(bf16) + 4 * 256;
}

int isc_gemm_tn(int precision, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C,
                int64_t ldc, int M, int N, int K, int act, void* workspace, size_t workspace_bytes,
                isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_TRY(check_precision(precision));
  ISC_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, "bad gemm arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Operand a, b;
  a.f32 = A;
  a.ld = lda;
  b.f32 = W;
  b.ld = ldw;
  if (precision != ISC_PREC_FP32) {
    if (!workspace || workspace_bytes < isc_gemm_workspace_bytes(precision, M, N, K)) {
      set_error("gemm workspace too small");
      return ISC_ERR_WORKSPACE;
    }
    Bump bump(workspace);
    const bool x3 = precision == ISC_PREC_BF16X3;
    bf16* ahi = bump.take<bf16>((size_t)M * K);
    bf16* alo = x3 ? bump.take<bf16>((size_t)M * K) : nullptr;
    bf16* bhi = bump.take<bf16>((size_t)N * K);
    bf16* blo = x3 ? bump.take<bf16>((size_t)N * K) : nullptr;
    ISC_TRY(split_planes(A, lda, ahi, alo, K, M, K, s));
    ISC_TRY(split_planes(W, ldw, bhi, blo, K, N, K, s));
    a.hi = ahi;
    a.lo = alo;
    a.ldp = K;
    b.hi = bhi;
    b.lo = blo;
    b.ldp = K;
  }
  Epilogue ep;
  ep.bias = bias;
  ep.act = act;
  Dest d;
  d.f32 = C;
  d.ld = ldc;
  return gemm(precision, a, b, d, M, N, K, ep, s);
}

}  // extern "C"

// =====================================================================================================
// Training: teacher-forced forward with a tape + hand-written backward (forward_xe :194-240,
// forward_seq2seq :242-288 and the REINFORCE re-score of forward_rl samples, train_xe.py:144-196,
// models/decoder.py:86-176). Dense contractions of the backward pass run on the tcgen05 GEMM with
// transposed operand planes; dW terms contract over ALL steps at once (K = T * B).
// =====================================================================================================
namespace isc {
namespace {

struct PM {  // bf16 hi/lo plane matrix
  bf16* hi = nullptr;
  bf16* lo = nullptr;
  long long ld = 0;
};
Operand op_of(const PM& p, long long row0 = 0, long long col0 = 0) {
  Operand o;
  o.hi = p.hi + row0 * p.ld + col0;
  o.lo = p.lo ? p.lo + row0 * p.ld + col0 : nullptr;
  o.ldp = p.ld;
  return o;
}
RowDest rd_of(float* f32, long long ld, const PM& p) {
  RowDest r;
  r.f32 = f32;
  r.ld = ld;
  r.hi = p.hi;
  r.lo = p.lo;
  r.ldp = p.ld;
  return r;
}
PM rows_from(const PM& p, long long row0) {
  PM q = p;
  q.hi = p.hi + row0 * p.ld;
  q.lo = p.lo ? p.lo + row0 * p.ld : nullptr;
  return q;
}
long long pad8(long long n) { return (n + 7) & ~7LL; }

struct TrainWs {
  isc_feats_t f;
  void* pro_ws;
  size_t pro_bytes;
  float* cpt_mean;  // inside pro_ws
  long long* it;  // [T][M] tokens actually fed (ground truth, or scheduled-sampling draws)
  // tape
  float *state_h, *state_c;  // [T+1][2][M][H]
  PM pX1, pX2, pcs, phL;     // [T*M][.]
  float *gates1, *gates2, *hproj, *cs, *g3, *gate_w, *cont_w, *senti_w;
  // backward scratch
  float* dlogits;  // [T*M][Vp], step-major
  PM pdlog;        // [T*M][Vp]
  PM dlogT;        // [V][TMp]
  float* dh_all;   // [T*M][H] d h_lang through the classifier, all steps
  float* dh_tmp;   // [M][H]
  float* dX1[2];   // [M][3H]
  float* dX2[2];
  float *dg1, *dg2;  // [T*M][4H]: all steps (bias colsums and transposed planes are taken once, after the loop)
  PM pdg1, pdg2;
  PM dg1T, dg2T;  // [4H][TMp]
  float* dhproj;  // [T*M][3H], all steps
  PM pdhproj;
  PM dhprojT;  // [3H][TMp]
  float* dcs;  // [M][2H]
  float *dc_att, *dc_lang;
  float *datt, *dp_att, *dsw, *dp_sw, *dpre_word, *dpre_gates;
  PM X1T, X2T, csT, hLT;  // [.][TMp]
  PM W1T, W2T, W3T, W4T, W5T, WpreT, Wl2wT, Wa2aT, Ws2aT, WcptT;
  // end phase
  PM pB4;   // [B][4H] planes of dpre_gates
  PM B4T;   // [4H][Bp]
  float* dfcsl;  // [B][2H]
  float *tB1, *tB2;  // [B][H] temporaries
  PM ptB;   // [B][H] planes
  PM tBT;   // [H][Bp]
  PM fcT, slT, cptT;  // [H][Bp]
  PM rawfcT;  // [D][Bp]
  float* dz;  // [B*L][H] (also used for [B*S][H])
  PM pdz;
  PM dzT;   // [H][BLp]
  PM attT;  // [H][BLp]
  PM rawT;  // [D][BLp]
  float* datt2;  // [B*L][H]
  long long TMp, Bp, BLp, Vp;
  size_t total;
};

TrainWs carve_train(const isc_dims_t& d, int precision, int B, int T, void* base) {
  Bump b(base);
  TrainWs w;
  memset(&w, 0, sizeof(w));
  const bool x3 = precision == ISC_PREC_BF16X3;
  const size_t m = (size_t)B, L = d.n_regions, S = d.n_senti, V = d.vocab, D = d.feat_dim;
  w.TMp = pad8((long long)T * B);
  w.Bp = pad8(B);
  const size_t rowsBL = m * (L > S ? L : S);
  w.BLp = pad8((long long)rowsBL);
  w.Vp = pad8(V);
  auto pm = [&](PM& p, size_t rows, size_t ld) {
    p.ld = (long long)ld;
    p.hi = b.take<bf16>(rows * ld);
    p.lo = x3 ? b.take<bf16>(rows * ld) : nullptr;
  };
  w.f.fc = b.take<float>(m * H);
  w.f.att = b.take<float>(m * L * H);
  w.f.p_att = b.take<float>(m * L * H);
  w.f.sw = b.take<float>(m * S * H);
  w.f.p_sw = b.take<float>(m * S * H);
  w.f.sl = b.take<float>(m * H);
  w.f.pre_gates = b.take<float>(m * G4);
  w.f.pre_word = b.take<float>(m * H);
  w.f.cpt_feats = b.take<float>(m * H);
  w.pro_bytes = carve_prologue(d, precision, B, nullptr).total;
  w.pro_ws = b.take<uint8_t>(w.pro_bytes);
  w.cpt_mean = base ? carve_prologue(d, precision, B, w.pro_ws).tmp : nullptr;
  const size_t tm = (size_t)T * m;
  w.it = b.take<long long>(tm);
  w.state_h = b.take<float>((size_t)(T + 1) * 2 * m * H);
  w.state_c = b.take<float>((size_t)(T + 1) * 2 * m * H);
  pm(w.pX1, tm, 3 * H);
  pm(w.pX2, tm, 3 * H);
  pm(w.pcs, tm, 2 * H);
  pm(w.phL, tm, H);
  w.gates1 = b.take<float>(tm * G4);
  w.gates2 = b.take<float>(tm * G4);
  w.hproj = b.take<float>(tm * 3 * H);
  w.cs = b.take<float>(tm * 2 * H);
  w.g3 = b.take<float>(tm * H);
  w.gate_w = b.take<float>(tm);
  w.cont_w = b.take<float>(tm * L);
  w.senti_w = b.take<float>(tm * S);
  w.dlogits = b.take<float>((size_t)T * m * w.Vp);
  pm(w.pdlog, (size_t)T * m, w.Vp);
  w.dh_all = b.take<float>((size_t)T * m * H);
  pm(w.dlogT, V, w.TMp);
  w.dh_tmp = b.take<float>(m * H);
  for (int i = 0; i < 2; ++i) {
    w.dX1[i] = b.take<float>(m * 3 * H);
    w.dX2[i] = b.take<float>(m * 3 * H);
  }
  w.dg1 = b.take<float>((size_t)T * m * G4);
  w.dg2 = b.take<float>((size_t)T * m * G4);
  pm(w.pdg1, m, G4);
  pm(w.pdg2, m, G4);
  pm(w.dg1T, G4, w.TMp);
  pm(w.dg2T, G4, w.TMp);
  w.dhproj = b.take<float>((size_t)T * m * 3 * H);
  pm(w.pdhproj, m, 3 * H);
  pm(w.dhprojT, 3 * H, w.TMp);
  w.dcs = b.take<float>(m * 2 * H);
  w.dc_att = b.take<float>(m * H);
  w.dc_lang = b.take<float>(m * H);
  w.datt = b.take<float>(m * L * H);
  w.dp_att = b.take<float>(m * L * H);
  w.dsw = b.take<float>(m * S * H);
  w.dp_sw = b.take<float>(m * S * H);
  w.dpre_word = b.take<float>(m * H);
  w.dpre_gates = b.take<float>(m * G4);
  pm(w.X1T, 3 * H, w.TMp);
  pm(w.X2T, 3 * H, w.TMp);
  pm(w.csT, 2 * H, w.TMp);
  pm(w.hLT, H, w.TMp);
  pm(w.W1T, 3 * H, G4);
  pm(w.W2T, H, 3 * H);
  pm(w.W3T, 2 * H, H);
  pm(w.W4T, 3 * H, G4);
  pm(w.W5T, H, w.Vp);
  pm(w.WpreT, 2 * H, G4);
  pm(w.Wl2wT, H, H);
  pm(w.Wa2aT, H, H);
  pm(w.Ws2aT, H, H);
  pm(w.WcptT, H, H);
  pm(w.pB4, m, G4);
  pm(w.B4T, G4, w.Bp);
  w.dfcsl = b.take<float>(m * 2 * H);
  w.tB1 = b.take<float>(m * H);
  w.tB2 = b.take<float>(m * H);
  pm(w.ptB, m, H);
  pm(w.tBT, H, w.Bp);
  pm(w.fcT, H, w.Bp);
  pm(w.slT, H, w.Bp);
  pm(w.cptT, H, w.Bp);
  pm(w.rawfcT, D, w.Bp);
  w.dz = b.take<float>(rowsBL * H);
  pm(w.pdz, rowsBL, H);
  pm(w.dzT, H, w.BLp);
  pm(w.attT, H, w.BLp);
  pm(w.rawT, D, w.BLp);
  w.datt2 = b.take<float>(rowsBL * H);
  w.total = (b.off + 255) & ~size_t(255);
  return w;
}

int zero_pm(const PM& p, size_t rows, cudaStream_t s) {
  ISC_CUDA(cudaMemsetAsync(p.hi, 0, rows * p.ld * sizeof(bf16), s));
  if (p.lo) ISC_CUDA(cudaMemsetAsync(p.lo, 0, rows * p.ld * sizeof(bf16), s));
  return 0;
}
// fp32 [rows][cols] -> transposed planes [cols][ldT] at column col0
int to_T(const float* src, long long ld, long long rows, int cols, const PM& dst, long long col0, cudaStream_t s) {
  return launch_split_transpose(src, ld, rows, cols, dst.hi, dst.lo, dst.ld, col0, s);
}
int planes_T(const PM& src, long long rows, int cols, const PM& dst, cudaStream_t s) {
  ISC_TRY(launch_transpose_bf16(src.hi, src.ld, rows, cols, dst.hi, dst.ld, 0, s));
  if (src.lo && dst.lo) ISC_TRY(launch_transpose_bf16(src.lo, src.ld, rows, cols, dst.lo, dst.ld, 0, s));
  return 0;
}
int passes_of(int precision) { return precision == ISC_PREC_BF16X3 ? 3 : 1; }

// C[M][N] (+)= A[M][K] . W[N][K]^T on planes
int pgemm(int precision, const Operand& A, const Operand& W, float* C, long long ldc, int M, int N, int K, bool accumulate,
          cudaStream_t s, const PM* planes_out = nullptr) {
  if (!C && !planes_out) return 0;
  Epilogue ep;
  if (accumulate) {
    ep.addmat = C;
    ep.ld_addmat = ldc;
  }
  Dest d;
  d.f32 = C;
  d.ld = ldc;
  if (planes_out) {
    d.hi = planes_out->hi;
    d.lo = planes_out->lo;
    d.ldp = planes_out->ld;
  }
  return gemm_tc(A, W, d, M, N, K, passes_of(precision), ep, s);
}

struct TrainMode {
  bool att, sw;  // content / sentiment attention active
};
TrainMode mode_of(int mode) {
  TrainMode m;
  m.att = mode != ISC_MODE_SEQ2SEQ;
  m.sw = mode != ISC_MODE_XE;
  return m;
}

DecodeWs tape_view(const TrainWs& w, int t, int B) {
  DecodeWs v;
  memset(&v, 0, sizeof(v));
  const long long r0 = (long long)t * B;
  auto pl = [&](const PM& p, long long cols) {
    Planes q;
    q.hi = p.hi + r0 * cols;
    q.lo = p.lo ? p.lo + r0 * cols : nullptr;
    return q;
  };
  v.pX1 = pl(w.pX1, 3 * H);
  v.pX2 = pl(w.pX2, 3 * H);
  v.pcs = pl(w.pcs, 2 * H);
  v.phL = pl(w.phL, H);
  v.tape = true;
  v.gates = w.gates1 + r0 * G4;
  v.gates2 = w.gates2 + r0 * G4;
  v.hproj = w.hproj + r0 * 3 * H;
  v.cs = w.cs + r0 * 2 * H;
  v.g3 = w.g3 + r0 * H;
  return v;
}

}  // namespace
}  // namespace isc

extern "C" {

size_t isc_train_workspace_bytes(const isc_dims_t* dims, int precision, int B, int n_steps) {
  if (check_dims(dims) != 0 || check_precision(precision) != 0 || B <= 0 || n_steps <= 0) return 0;
  return carve_train(*dims, precision, B, n_steps, nullptr).total;
}

int isc_train_forward(const isc_dims_t* dims, const void* packed, int precision, int mode, const float* fc_feats,
                      const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                      const int64_t* senti_labels, int B, const int64_t* inputs, int64_t ld_inputs, int n_steps,
                      const isc_dropout_t* dropout, const isc_sched_sampling_t* ss, float* logprobs, float* fc_embedded,
                      float* cpt_feats, void* workspace, size_t workspace_bytes, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_TRY(check_dims(dims));
  ISC_REQUIRE(precision == ISC_PREC_BF16X3, "training runs in ISC_PREC_BF16X3 only");
  ISC_REQUIRE(mode == ISC_MODE_XE || mode == ISC_MODE_SEQ2SEQ || mode == ISC_MODE_RL, "unknown mode %d", mode);
  ISC_REQUIRE(B > 0 && n_steps > 0 && n_steps <= T_MAX && inputs && logprobs && ld_inputs >= n_steps && senti_labels,
              "bad train_forward arguments");
  TrainWs w = carve_train(*dims, precision, B, n_steps, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("train workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  const TrainMode tm = mode_of(mode);
  ISC_REQUIRE(!tm.att || (fc_feats && att_feats), "fc_feats / att_feats missing");
  ISC_REQUIRE(!tm.sw || senti_words, "senti_words missing");
  ISC_REQUIRE(tm.att || cpt_words, "seq2seq needs cpt_words");
  isc_feats_t f = w.f;
  if (!tm.att) f.att = f.p_att = nullptr;
  if (!tm.sw) f.sw = f.p_sw = f.pre_word = nullptr;
  if (!cpt_words) f.cpt_feats = nullptr;
  ISC_TRY(run_prologue(dims, packed, precision, fc_feats, att_feats, cpt_words, n_cpt, tm.sw ? senti_words : nullptr,
                       senti_labels, B, tm.att ? 0 : 1, &f, w.pro_ws, w.pro_bytes, stream, dropout, fc_embedded));
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, &f, stream));
  if (cpt_feats && f.cpt_feats) ISC_TRY(copy_block(cpt_feats, H, f.cpt_feats, H, B, H, c.s));
  const size_t st = (size_t)2 * B * H * sizeof(float);
  ISC_CUDA(cudaMemsetAsync(w.state_h, 0, st, c.s));
  ISC_CUDA(cudaMemsetAsync(w.state_c, 0, st, c.s));
  const long long V = dims->vocab, L = dims->n_regions, S = dims->n_senti;
  for (int t = 0; t < n_steps; ++t) {
    long long* it_t = w.it + (size_t)t * B;
    if (ss && ss->prob > 0.f && ss->uniform && t >= 1) {
      ISC_TRY(launch_ss_select(logprobs + (long long)(t - 1) * V, (long long)n_steps * V,
                               reinterpret_cast<const long long*>(inputs) + t, ld_inputs, ss->uniform + (size_t)t * B, ss->prob,
                               ss->noise ? ss->noise + (size_t)t * B * V : nullptr, ss->seed, t, B, (int)V, it_t, c.s));
    } else {
      ISC_CUDA(cudaMemcpy2DAsync(it_t, sizeof(long long), inputs + t, ld_inputs * sizeof(long long), sizeof(long long), B,
                                 cudaMemcpyDeviceToDevice, c.s));
    }
    DecodeWs v = tape_view(w, t, B);
    StepIO io;
    io.it = it_t;
    io.parent = nullptr;
    io.h_in = w.state_h + (size_t)t * 2 * B * H;
    io.c_in = w.state_c + (size_t)t * 2 * B * H;
    io.h_out = w.state_h + (size_t)(t + 1) * 2 * B * H;
    io.c_out = w.state_c + (size_t)(t + 1) * 2 * B * H;
    io.logits = logprobs + (long long)t * V;
    io.ld_logits = (long long)n_steps * V;
    if (tm.att) {
      io.cont_w = w.cont_w + (long long)t * B * L;
      io.ld_cont_w = L;
    }
    if (tm.sw) {
      io.senti_w = w.senti_w + (long long)t * B * S;
      io.ld_senti_w = S;
    }
    if (tm.att && tm.sw) {
      io.gate_w = w.gate_w + (long long)t * B;
      io.ld_gate_w = 1;
    }
    if (dropout && dropout->out) {
      io.out_mask = dropout->out + (size_t)t * B * H;
      io.drop_scale = dropout->scale;
    }
    ISC_TRY(run_step(c, v, B, 1, io));
    ISC_TRY(launch_log_softmax(io.logits, io.ld_logits, B, (int)V, c.s));
  }
  return 0;
}

int isc_train_backward(const isc_dims_t* dims, const void* packed, int precision, int mode, const float* fc_feats,
                       const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                       const int64_t* senti_labels, int B, const int64_t* inputs, int64_t ld_inputs, int n_steps,
                       const isc_dropout_t* dropout, const float* logprobs, const float* dlogprobs, const int64_t* targets,
                       int64_t ld_targets, const float* coef, const float* d_cpt_feats, const isc_grads_t* g, void* workspace,
                       size_t workspace_bytes, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_TRY(check_dims(dims));
  ISC_REQUIRE(precision == ISC_PREC_BF16X3, "training runs in ISC_PREC_BF16X3 only");
  ISC_REQUIRE(mode == ISC_MODE_XE || mode == ISC_MODE_SEQ2SEQ || mode == ISC_MODE_RL, "unknown mode %d", mode);
  ISC_REQUIRE(B > 0 && n_steps > 0 && inputs && logprobs && g && senti_labels, "bad train_backward arguments");
  ISC_REQUIRE(dlogprobs || (targets && coef) || d_cpt_feats, "no incoming gradient");
  TrainWs w = carve_train(*dims, precision, B, n_steps, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("train workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  const TrainMode tm = mode_of(mode);
  isc_feats_t f = w.f;
  Ctx c;
  ISC_TRY(make_ctx(c, dims, packed, precision, &f, stream));
  const Packed& pk = c.pk;
  cudaStream_t s = c.s;
  const int T = n_steps, M = B;
  const long long V = dims->vocab, L = dims->n_regions, S = dims->n_senti, D = dims->feat_dim;
  const long long TM = (long long)T * M;
  const int TMk = (int)w.TMp, Bk = (int)w.Bp;
  const float dscale = dropout ? dropout->scale : 1.0f;
  const bool have_seq = dlogprobs || (targets && coef);
  auto zf = [&](float* p, size_t n) { return cudaMemsetAsync(p, 0, n * sizeof(float), s); };

  // ---- transposed weight planes (the dX GEMMs contract over the weights' output dimension)
  ISC_TRY(to_T(pk.W1.f32, 3 * H, G4, 3 * H, w.W1T, 0, s));
  ISC_TRY(to_T(pk.W2.f32, H, 3 * H, H, w.W2T, 0, s));
  ISC_TRY(to_T(pk.W3.f32, 2 * H, H, 2 * H, w.W3T, 0, s));
  ISC_TRY(to_T(pk.W4.f32, 3 * H, G4, 3 * H, w.W4T, 0, s));
  ISC_TRY(zero_pm(w.W5T, H, s));
  ISC_TRY(to_T(pk.W5.f32, H, V, H, w.W5T, 0, s));
  ISC_TRY(to_T(pk.Wpre.f32, 2 * H, G4, 2 * H, w.WpreT, 0, s));
  ISC_TRY(to_T(pk.Wl2w.f32, H, H, H, w.Wl2wT, 0, s));
  ISC_TRY(to_T(pk.Wa2a.f32, H, H, H, w.Wa2aT, 0, s));
  ISC_TRY(to_T(pk.Ws2a.f32, H, H, H, w.Ws2aT, 0, s));
  ISC_TRY(to_T(pk.Wcpt.f32, H, H, H, w.WcptT, 0, s));

  if (have_seq) {
    // ---- zero the carries and accumulators
    for (int i = 0; i < 2; ++i) {
      ISC_CUDA(zf(w.dX1[i], (size_t)M * 3 * H));
      ISC_CUDA(zf(w.dX2[i], (size_t)M * 3 * H));
    }
    ISC_CUDA(zf(w.dc_att, (size_t)M * H));
    ISC_CUDA(zf(w.dc_lang, (size_t)M * H));
    ISC_CUDA(zf(w.datt, (size_t)M * L * H));
    ISC_CUDA(zf(w.dp_att, (size_t)M * L * H));
    ISC_CUDA(zf(w.dsw, (size_t)M * S * H));
    ISC_CUDA(zf(w.dp_sw, (size_t)M * S * H));
    ISC_CUDA(zf(w.dpre_word, (size_t)M * H));
    ISC_CUDA(zf(w.dpre_gates, (size_t)M * G4));
    ISC_CUDA(zf(w.dhproj, (size_t)TM * 3 * H));
    if (w.Vp != V) ISC_CUDA(zf(w.dlogits, (size_t)TM * w.Vp));  // pad columns must read as zero
    ISC_TRY(zero_pm(w.pdhproj, M, s));

    ISC_TRY(zero_pm(w.dlogT, V, s));
    ISC_TRY(zero_pm(w.dg1T, G4, s));
    ISC_TRY(zero_pm(w.dg2T, G4, s));
    ISC_TRY(zero_pm(w.dhprojT, 3 * H, s));
    ISC_TRY(zero_pm(w.X1T, 3 * H, s));
    ISC_TRY(zero_pm(w.X2T, 3 * H, s));
    ISC_TRY(zero_pm(w.csT, 2 * H, s));
    ISC_TRY(zero_pm(w.hLT, H, s));

    // ---- classifier backward for ALL steps at once (none of it depends on the recurrence): d logits from the
    // log-softmax backward, their planes (operand of d h_lang = d logits . W5) and transposed planes (operand of dW5)
    ISC_TRY(launch_logsoftmax_bwd(logprobs, dlogprobs, reinterpret_cast<const long long*>(targets), ld_targets, coef, T, M, (int)V,
                                  w.dlogits, w.Vp, s));
    ISC_TRY(split_planes(w.dlogits, w.Vp, w.pdlog.hi, w.pdlog.lo, w.Vp, TM, (int)w.Vp, s));
    ISC_TRY(to_T(w.dlogits, w.Vp, TM, (int)V, w.dlogT, 0, s));
    if (g->classifier_b) ISC_TRY(launch_colsum_add(w.dlogits, w.Vp, TM, (int)V, g->classifier_b, nullptr, s));
    ISC_TRY(pgemm(precision, op_of(w.pdlog), op_of(w.W5T), w.dh_all, H, (int)TM, H, (int)w.Vp, false, s));

    // which slice of the h projections is live: [cont h2att | senti h2word | gate h2att]
    const int hp0 = tm.att ? 0 : H, hpn = (tm.att && tm.sw) ? 3 * H : H;

    for (int t = T - 1; t >= 0; --t) {
      const int cur = t & 1, nxt = cur ^ 1;
      const long long r0 = (long long)t * M;
      const float* h_prev_c = w.state_c + (size_t)t * 2 * M * H;
      const float* c_new = w.state_c + (size_t)(t + 1) * 2 * M * H;
      float* dg1 = w.dg1 + r0 * G4;        // this step's slices of the all-steps gradient matrices
      float* dg2 = w.dg2 + r0 * G4;
      float* dhproj = w.dhproj + r0 * 3 * H;
      // 3. language LSTM backward
      ISC_TRY(launch_lstm_bwd(w.gates2 + r0 * G4, h_prev_c + (size_t)M * H, c_new + (size_t)M * H, w.dh_all + r0 * H, H,
                              (dropout && dropout->out) ? dropout->out + (size_t)t * M * H : nullptr, dscale, w.dX1[nxt], 3 * H,
                              w.dX2[nxt] + 2 * H, 3 * H, w.dc_lang, dg2, rd_of(nullptr, 0, w.pdg2), M, s));
      // 4. d [ctx | h_att | h_lang_prev]
      ISC_TRY(pgemm(precision, op_of(w.pdg2), op_of(w.W4T), w.dX2[cur], 3 * H, M, 3 * H, G4, false, s));
      // 5. gate
      const float* dcs = w.dX2[cur];
      long long ld_dcs = 3 * H;
      int cont_col = 0, senti_col = 0;
      if (tm.att && tm.sw) {
        ISC_TRY(launch_gate_bwd(w.dX2[cur], 3 * H, w.cs + r0 * 2 * H, w.g3 + r0 * H, w.gate_w + r0, pk.alpha_g, w.dcs,
                                rd_of(dhproj, 3 * H, w.pdhproj), 2 * H, g->g_alpha_w, g->g_alpha_b, M, s));
        ISC_TRY(pgemm(precision, op_of(w.pdhproj, 0, 2 * H), op_of(w.W3T), w.dcs, 2 * H, M, 2 * H, H, true, s));
        dcs = w.dcs;
        ld_dcs = 2 * H;
        senti_col = H;
      }
      // 6. attention
      AttnBwdParams ab;
      ab.L = (int)L;
      ab.S = (int)S;
      ab.dcs = dcs;
      ab.ld_dcs = ld_dcs;
      ab.cont_col = cont_col;
      ab.senti_col = senti_col;
      ab.hproj = w.hproj + r0 * 3 * H;
      ab.ld_hproj = 3 * H;
      ab.pre_word = tm.sw ? f.pre_word : nullptr;
      if (tm.att) {
        ab.att = static_cast<const float*>(f.att);
        ab.ea_att = static_cast<const float*>(f.p_att);
        ab.cont_w = w.cont_w + r0 * L;
        ab.datt = w.datt;
        ab.dp_att = w.dp_att;
        ab.alpha_c = pk.alpha_c;
        ab.dalpha_c = g->ca_alpha_w;
      }
      if (tm.sw) {
        ab.sw = f.sw;
        ab.ea_sw = f.p_sw;
        ab.senti_w = w.senti_w + r0 * S;
        ab.dsw = w.dsw;
        ab.dp_sw = w.dp_sw;
        ab.alpha_s = pk.alpha_s;
        ab.dalpha_s = g->sa_alpha_w;
        ab.dpre_word = w.dpre_word;
      }
      ab.dhproj = rd_of(dhproj, 3 * H, w.pdhproj);
      ISC_TRY(launch_attention_bwd(ab, M, s));
      // 7. d h_att through the projections
      ISC_TRY(pgemm(precision, op_of(w.pdhproj, 0, hp0), op_of(w.W2T, 0, hp0), w.dh_tmp, H, M, H, hpn, false, s));
      // 8. attention LSTM backward
      ISC_TRY(launch_lstm_bwd(w.gates1 + r0 * G4, h_prev_c, c_new, w.dh_tmp, H, nullptr, 1.f, w.dX2[cur] + H, 3 * H,
                              w.dX1[nxt] + 2 * H, 3 * H, w.dc_att, dg1, rd_of(nullptr, 0, w.pdg1), M, s));
      // 9. d [h_lang_prev | xt | h_att_prev]
      ISC_TRY(pgemm(precision, op_of(w.pdg1), op_of(w.W1T), w.dX1[cur], 3 * H, M, 3 * H, G4, false, s));
      // 10. word embedding of this step's input tokens
      ISC_TRY(launch_embed_bwd(w.it + (size_t)t * M, 1, M, 1, 0, dims->pad_id, 1, (int)V, pk.emb, w.dX1[cur] + H, 3 * H, 1, nullptr,
                               1.f, 1.f, g->word_embed, s));
    }

    // ---- bias gradients and transposed planes of the all-steps gradient matrices, once
    ISC_TRY(launch_colsum_add(w.dg2, G4, TM, G4, g->lang_lstm_b_ih, g->lang_lstm_b_hh, s));
    ISC_TRY(launch_colsum_add(w.dg1, G4, TM, G4, g->att_lstm_b_ih, g->att_lstm_b_hh, s));
    ISC_TRY(to_T(w.dg2, G4, TM, G4, w.dg2T, 0, s));
    ISC_TRY(to_T(w.dg1, G4, TM, G4, w.dg1T, 0, s));
    ISC_TRY(to_T(w.dhproj + hp0, 3 * H, TM, hpn, rows_from(w.dhprojT, hp0), 0, s));
    if (tm.att) ISC_TRY(launch_colsum_add(w.dhproj, 3 * H, TM, H, g->ca_h2att_b, nullptr, s));
    if (tm.sw) ISC_TRY(launch_colsum_add(w.dhproj + H, 3 * H, TM, H, g->sa_h2word_b, nullptr, s));
    if (tm.att && tm.sw) {
      ISC_TRY(launch_colsum_add(w.dhproj + 2 * H, 3 * H, TM, H, g->g_h2att_b, g->g_cont2att_b, s));
      ISC_TRY(launch_colsum_add(w.dhproj + 2 * H, 3 * H, TM, H, g->g_senti2att_b, nullptr, s));
    }
    ISC_TRY(launch_sum_steps(w.dg1, T, (long long)M * G4, (long long)M * G4, w.dpre_gates, s));  // hoisted pre_gates: sum over steps

    // ---- weight gradients: one contraction over all T*B rows per matrix
    ISC_TRY(planes_T(w.pX1, TM, 3 * H, w.X1T, s));
    ISC_TRY(planes_T(w.pX2, TM, 3 * H, w.X2T, s));
    ISC_TRY(planes_T(w.phL, TM, H, w.hLT, s));
    if (tm.att && tm.sw) ISC_TRY(planes_T(w.pcs, TM, 2 * H, w.csT, s));
    ISC_TRY(pgemm(precision, op_of(w.dlogT), op_of(w.hLT), g->classifier_w, H, (int)V, H, TMk, true, s));
    // language LSTM: W_ih [4H][2H] <- X2 rows [ctx | h_att], W_hh <- h_lang_prev
    ISC_TRY(pgemm(precision, op_of(w.dg2T), op_of(w.X2T), g->lang_lstm_w_ih, 2 * H, G4, 2 * H, TMk, true, s));
    ISC_TRY(pgemm(precision, op_of(w.dg2T), op_of(w.X2T, 2 * H), g->lang_lstm_w_hh, H, G4, H, TMk, true, s));
    // attention LSTM: W_ih[:, 0:H] <- h_lang_prev, W_ih[:, 2H:3H] <- xt, W_hh <- h_att_prev (fc slice: hoisted, below)
    ISC_TRY(pgemm(precision, op_of(w.dg1T), op_of(w.X1T), g->att_lstm_w_ih, 3 * H, G4, H, TMk, true, s));
    ISC_TRY(pgemm(precision, op_of(w.dg1T), op_of(w.X1T, H), g->att_lstm_w_ih + 2 * H, 3 * H, G4, H, TMk, true, s));
    ISC_TRY(pgemm(precision, op_of(w.dg1T), op_of(w.X1T, 2 * H), g->att_lstm_w_hh, H, G4, H, TMk, true, s));
    // h projections (h_att = X2 rows H..2H)
    if (tm.att) ISC_TRY(pgemm(precision, op_of(w.dhprojT), op_of(w.X2T, H), g->ca_h2att_w, H, H, H, TMk, true, s));
    if (tm.sw) ISC_TRY(pgemm(precision, op_of(w.dhprojT, H), op_of(w.X2T, H), g->sa_h2word_w, H, H, H, TMk, true, s));
    if (tm.att && tm.sw) {
      ISC_TRY(pgemm(precision, op_of(w.dhprojT, 2 * H), op_of(w.X2T, H), g->g_h2att_w, H, H, H, TMk, true, s));
      ISC_TRY(pgemm(precision, op_of(w.dhprojT, 2 * H), op_of(w.csT, 0), g->g_cont2att_w, H, H, H, TMk, true, s));
      ISC_TRY(pgemm(precision, op_of(w.dhprojT, 2 * H), op_of(w.csT, H), g->g_senti2att_w, H, H, H, TMk, true, s));
    }

    // ---- hoisted step-invariant terms: pre_gates = [fc | sl] Wpre^T (+ biases, already covered by colsum(dg1))
    ISC_TRY(split_planes(w.dpre_gates, G4, w.pB4.hi, w.pB4.lo, G4, M, G4, s));
    ISC_TRY(zero_pm(w.B4T, G4, s));
    ISC_TRY(to_T(w.dpre_gates, G4, M, G4, w.B4T, 0, s));
    ISC_TRY(pgemm(precision, op_of(w.pB4), op_of(w.WpreT), w.dfcsl, 2 * H, M, 2 * H, G4, false, s));  // d [fc | sl]
    ISC_TRY(zero_pm(w.fcT, H, s));
    ISC_TRY(zero_pm(w.slT, H, s));
    ISC_TRY(to_T(f.fc, H, M, H, w.fcT, 0, s));
    ISC_TRY(to_T(f.sl, H, M, H, w.slT, 0, s));
    ISC_TRY(pgemm(precision, op_of(w.B4T), op_of(w.fcT), g->att_lstm_w_ih + H, 3 * H, G4, H, Bk, true, s));
    ISC_TRY(pgemm(precision, op_of(w.B4T), op_of(w.slT), g->att_lstm_w_ih + 2 * H, 3 * H, G4, H, Bk, true, s));
    const float* dsl_extra = nullptr;
    if (tm.sw) {  // label2word(sl), the label term of the sentiment-attention query
      ISC_TRY(launch_colsum_add(w.dpre_word, H, M, H, g->sa_label2word_b, nullptr, s));
      ISC_TRY(zero_pm(w.tBT, H, s));
      ISC_TRY(to_T(w.dpre_word, H, M, H, w.tBT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.tBT), op_of(w.slT), g->sa_label2word_w, H, H, H, Bk, true, s));
      ISC_TRY(split_planes(w.dpre_word, H, w.ptB.hi, w.ptB.lo, H, M, H, s));
      ISC_TRY(pgemm(precision, op_of(w.ptB), op_of(w.Wl2wT), w.tB1, H, M, H, H, false, s));
      dsl_extra = w.tB1;
    }
    ISC_TRY(launch_relu_mask_bwd(w.dfcsl + H, 2 * H, dsl_extra, H, nullptr, 0, 0, nullptr, 0, 1.f, M, H, w.tB2, H, RowDest(), s));
    ISC_TRY(launch_embed_bwd(reinterpret_cast<const long long*>(senti_labels), 1, M, 1, 0, -1, 0, dims->n_labels, pk.lab_emb, w.tB2, H,
                             1, dropout ? dropout->sl : nullptr, dscale, 1.f, g->senti_label_embed, s));
    if (tm.att) {
      // fc = dropout(ReLU(fc_embed(fc_feats)))
      ISC_TRY(launch_relu_mask_bwd(w.dfcsl, 2 * H, nullptr, 0, f.fc, H, 0, dropout ? dropout->fc : nullptr, H, dscale, M, H, w.tB1,
                                   H, RowDest(), s));
      ISC_TRY(launch_colsum_add(w.tB1, H, M, H, g->fc_embed_b, nullptr, s));
      ISC_TRY(zero_pm(w.tBT, H, s));
      ISC_TRY(to_T(w.tB1, H, M, H, w.tBT, 0, s));
      ISC_TRY(zero_pm(w.rawfcT, D, s));
      ISC_TRY(to_T(fc_feats, D, M, (int)D, w.rawfcT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.tBT), op_of(w.rawfcT), g->fc_embed_w, D, H, (int)D, Bk, true, s));
      // region features: p_att = ReLU(att2att(att)), att = dropout(ReLU(att_embed(raw)))
      const long long BL = (long long)M * L;
      ISC_TRY(launch_relu_mask_bwd(w.dp_att, H, nullptr, 0, static_cast<const float*>(f.p_att), H, 1, nullptr, 0, 1.f, BL, H, w.dz,
                                   H, rd_of(nullptr, 0, w.pdz), s));
      ISC_TRY(launch_colsum_add(w.dz, H, BL, H, g->att2att_b, nullptr, s));
      ISC_TRY(zero_pm(w.dzT, H, s));
      ISC_TRY(zero_pm(w.attT, H, s));
      ISC_TRY(to_T(w.dz, H, BL, H, w.dzT, 0, s));
      ISC_TRY(to_T(static_cast<const float*>(f.att), H, BL, H, w.attT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.dzT), op_of(w.attT), g->att2att_w, H, H, H, (int)pad8(BL), true, s));
      ISC_TRY(pgemm(precision, op_of(w.pdz), op_of(w.Wa2aT), w.datt2, H, (int)BL, H, H, false, s));
      ISC_TRY(launch_relu_mask_bwd(w.datt, H, w.datt2, H, static_cast<const float*>(f.att), H, 0, dropout ? dropout->att : nullptr, H,
                                   dscale, BL, H, w.dz, H, RowDest(), s));
      ISC_TRY(launch_colsum_add(w.dz, H, BL, H, g->att_embed_b, nullptr, s));
      ISC_TRY(zero_pm(w.dzT, H, s));
      ISC_TRY(to_T(w.dz, H, BL, H, w.dzT, 0, s));
      ISC_TRY(zero_pm(w.rawT, D, s));
      ISC_TRY(to_T(att_feats, D, BL, (int)D, w.rawT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.dzT), op_of(w.rawT), g->att_embed_w, D, H, (int)D, (int)pad8(BL), true, s));
    }
    if (tm.sw) {
      // sentiment words: p_sw = ReLU(senti2att(sw)), sw = dropout(ReLU(word_embed([PAD | senti_words])))
      const long long BS = (long long)M * S;
      ISC_TRY(launch_relu_mask_bwd(w.dp_sw, H, nullptr, 0, f.p_sw, H, 1, nullptr, 0, 1.f, BS, H, w.dz, H, rd_of(nullptr, 0, w.pdz), s));
      ISC_TRY(launch_colsum_add(w.dz, H, BS, H, g->senti2att_b, nullptr, s));
      ISC_TRY(zero_pm(w.dzT, H, s));
      ISC_TRY(zero_pm(w.attT, H, s));
      ISC_TRY(to_T(w.dz, H, BS, H, w.dzT, 0, s));
      ISC_TRY(to_T(f.sw, H, BS, H, w.attT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.dzT), op_of(w.attT), g->senti2att_w, H, H, H, (int)pad8(BS), true, s));
      ISC_TRY(pgemm(precision, op_of(w.pdz), op_of(w.Ws2aT), w.datt2, H, (int)BS, H, H, false, s));
      ISC_TRY(launch_relu_mask_bwd(w.dsw, H, w.datt2, H, nullptr, 0, 0, nullptr, 0, 1.f, BS, H, w.dz, H, RowDest(), s));
      ISC_TRY(launch_embed_bwd(reinterpret_cast<const long long*>(senti_words), S - 1, M, S - 1, 1, dims->pad_id, 1, (int)V, pk.emb,
                               w.dz, H, S, dropout ? dropout->sw : nullptr, dscale, 1.f, g->word_embed, s));
    }
  }

  // ---- concept branch: cpt_feats = ReLU(cpt2fc(mean_j ReLU(word_embed(cpt_j)))) (captioner.py:201-204). Its gradient
  // comes from the domain-alignment loss (d_cpt_feats) and, in seq2seq mode, from the decoder (it stands in for fc).
  const bool dec_to_cpt = !tm.att && have_seq;
  if (cpt_words && (d_cpt_feats || dec_to_cpt)) {
    const float* a = d_cpt_feats;
    if (dec_to_cpt) {
      ISC_TRY(launch_relu_mask_bwd(w.dfcsl, 2 * H, d_cpt_feats, H, nullptr, 0, 0, nullptr, 0, 1.f, M, H, w.tB1, H, RowDest(), s));
      if (dropout && dropout->fc) {  // only the decoder's share went through the dropout
        ISC_TRY(launch_relu_mask_bwd(w.dfcsl, 2 * H, nullptr, 0, nullptr, 0, 0, dropout->fc, H, dscale, M, H, w.tB1, H, RowDest(), s));
        if (d_cpt_feats) ISC_TRY(launch_add2d(w.tB1, H, d_cpt_feats, H, M, H, s));
      }
      a = w.tB1;
    }
    ISC_TRY(launch_relu_mask_bwd(a, H, nullptr, 0, w.f.cpt_feats, H, 0, nullptr, 0, 1.f, M, H, w.tB2, H, rd_of(nullptr, 0, w.ptB), s));
    ISC_TRY(launch_colsum_add(w.tB2, H, M, H, g->cpt2fc_b, nullptr, s));
    ISC_TRY(zero_pm(w.tBT, H, s));
    ISC_TRY(zero_pm(w.cptT, H, s));
    ISC_TRY(to_T(w.tB2, H, M, H, w.tBT, 0, s));
    ISC_TRY(to_T(w.cpt_mean, H, M, H, w.cptT, 0, s));
    ISC_TRY(pgemm(precision, op_of(w.tBT), op_of(w.cptT), g->cpt2fc_w, H, H, H, Bk, true, s));
    ISC_TRY(pgemm(precision, op_of(w.ptB), op_of(w.WcptT), w.tB1, H, M, H, H, false, s));
    ISC_TRY(launch_embed_bwd(reinterpret_cast<const long long*>(cpt_words), n_cpt, M, n_cpt, 0, dims->pad_id, 1, (int)V, pk.emb, w.tB1,
                             H, 1, nullptr, 1.f, 1.0f / (float)n_cpt, g->word_embed, s));
  }
  return 0;
}

int isc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float clip, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_REQUIRE(params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad adam_step arguments");
  return launch_adam_clamp(params, grads, exp_avg, exp_avg_sq, n, clip, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                           static_cast<cudaStream_t>(stream));
}

}  // extern "C"
