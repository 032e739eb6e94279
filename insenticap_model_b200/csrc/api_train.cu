// C-ABI entry points of the training path (include/isc.h: isc_train_forward / isc_train_backward / isc_adam_step).
#include "engine.cuh"

using namespace isc;

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
  long long* it;  // [T + 1][M] tokens actually fed (ground truth, scheduled-sampling draws, or the sampled tokens of
                  // isc_train_forward_sample, whose last step writes row T)
  int* unfinished;   // [M]     isc_train_forward_sample: rows still decoding
  int* alive_count;  // [T_MAX] isc_train_forward_sample: unfinished rows after each step
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
  float *de_c, *de_s, *dctx_c, *dctx_s;  // per-step softmax gradients [T*M][L|S] and d ctx [T*M][H] (deferred attention backward)
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
  w.it = b.take<long long>(tm + m);
  w.unfinished = b.take<int>(m);
  w.alive_count = b.take<int>(T_MAX);
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
  w.de_c = b.take<float>(tm * L);
  w.de_s = b.take<float>(tm * S);
  w.dctx_c = b.take<float>(tm * H);
  w.dctx_s = b.take<float>(tm * H);
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
// Transposed operand planes are contracted over pad8(valid_cols) columns: only the padding columns need zeros, so a plane
// whose column count is a multiple of 8 (every one of them at the BASELINE sizes: T * B = 4096, B = 256, B * regions = 50176)
// is not cleared before it is overwritten — that was ~0.3 GB of memset per backward call.
int zero_pm_pad(const PM& p, size_t rows, long long valid_cols, cudaStream_t s) {
  return valid_cols % 8 == 0 ? 0 : zero_pm(p, rows, s);
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

// Optional events recorded inside isc_train_backward at the points where a group of parameter gradients is final
// (isc_train_backward_marks): lets the caller start the data-parallel all-reduce of those gradients while the rest of the
// backward (hoisted terms, prologue layers) still runs. Armed per call, per host thread.
thread_local cudaEvent_t g_marks[2] = {nullptr, nullptr};
int record_mark(int i, cudaStream_t s) {
  if (g_marks[i]) {
    ISC_CUDA(cudaEventRecord(g_marks[i], s));
    g_marks[i] = nullptr;
  }
  return 0;
}

}  // namespace
}  // namespace isc

extern "C" {

size_t isc_train_workspace_bytes(const isc_dims_t* dims, int precision, int B, int n_steps) {
  if (check_dims(dims) != 0 || check_precision(precision) != 0 || B <= 0 || n_steps <= 0) return 0;
  return carve_train(*dims, precision, B, n_steps, nullptr).total;
}

}  // extern "C"

namespace {
// isc_train_forward_sample: free-running sampling (Captioner.forward_rl :317-349) recorded on the tape
struct SampleOut {
  int sample_mode;            // 1 external noise, 2 counter-based Gumbel
  const float* noise;         // [n_steps, B, V] (mode 1)
  unsigned long long seed;
  long long* seq;             // [B, n_steps]
  float* seq_logprobs;        // [B, n_steps]
  float* seq_masks;           // [B, n_steps]
};

int train_forward_impl(const isc_dims_t* dims, const void* packed, int precision, int mode, const float* fc_feats,
                       const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                       const int64_t* senti_labels, int B, const int64_t* inputs, int64_t ld_inputs, int n_steps,
                       const isc_dropout_t* dropout, const isc_sched_sampling_t* ss, const SampleOut* so, float* logprobs,
                       float* fc_embedded, float* cpt_feats, void* workspace, size_t workspace_bytes, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_TRY(check_dims(dims));
  ISC_REQUIRE(precision == ISC_PREC_BF16X3, "training runs in ISC_PREC_BF16X3 only");
  ISC_REQUIRE(mode == ISC_MODE_XE || mode == ISC_MODE_SEQ2SEQ || mode == ISC_MODE_RL, "unknown mode %d", mode);
  ISC_REQUIRE(B > 0 && n_steps > 0 && n_steps <= T_MAX && logprobs && senti_labels && (so || (inputs && ld_inputs >= n_steps)),
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
  if (so) {
    ISC_CUDA(cudaMemsetAsync(so->seq, 0, (size_t)B * n_steps * sizeof(long long), c.s));
    ISC_CUDA(cudaMemsetAsync(so->seq_logprobs, 0, (size_t)B * n_steps * sizeof(float), c.s));
    ISC_CUDA(cudaMemsetAsync(so->seq_masks, 0, (size_t)B * n_steps * sizeof(float), c.s));
    ISC_CUDA(cudaMemsetAsync(w.alive_count, 0, T_MAX * sizeof(int), c.s));
    ISC_TRY(launch_greedy_init(w.it, w.unfinished, B, dims->sos_id, c.s));  // step 0 feeds <SOS>
  }
  for (int t = 0; t < n_steps; ++t) {
    long long* it_t = w.it + (size_t)t * B;
    if (so) {
      // fed token = what the previous step's selection wrote (it *= unfinished included)
    } else if (ss && ss->prob > 0.f && ss->uniform && t >= 1) {
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
    if (so) {  // the draw, its log-prob, the mask column and the next input, exactly as isc_decode_greedy forms them
      GreedyParams g;
      g.logits = io.logits;
      g.ld = io.ld_logits;
      g.B = B;
      g.V = (int)V;
      g.T = n_steps;
      g.t = t;
      g.sample_mode = so->sample_mode;
      g.noise = so->noise ? so->noise + (long long)t * B * V : nullptr;
      g.seed = so->seed;
      g.eos_id = dims->eos_id;
      g.it = w.it + (size_t)(t + 1) * B;
      g.unfinished = w.unfinished;
      g.alive_count = w.alive_count;
      g.seq = so->seq;
      g.seq_logprobs = so->seq_logprobs;
      g.seq_masks = so->seq_masks;
      ISC_TRY(launch_greedy_select(g, c.s));
    }
    ISC_TRY(launch_log_softmax(io.logits, io.ld_logits, B, (int)V, c.s));
  }
  return 0;
}
}  // namespace

extern "C" {

int isc_train_forward(const isc_dims_t* dims, const void* packed, int precision, int mode, const float* fc_feats,
                      const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                      const int64_t* senti_labels, int B, const int64_t* inputs, int64_t ld_inputs, int n_steps,
                      const isc_dropout_t* dropout, const isc_sched_sampling_t* ss, float* logprobs, float* fc_embedded,
                      float* cpt_feats, void* workspace, size_t workspace_bytes, isc_stream_t stream) {
  return train_forward_impl(dims, packed, precision, mode, fc_feats, att_feats, cpt_words, n_cpt, senti_words, senti_labels, B,
                            inputs, ld_inputs, n_steps, dropout, ss, nullptr, logprobs, fc_embedded, cpt_feats, workspace,
                            workspace_bytes, stream);
}

int isc_train_forward_sample(const isc_dims_t* dims, const void* packed, int precision, int mode, const float* fc_feats,
                             const float* att_feats, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                             const int64_t* senti_labels, int B, int n_steps, const isc_dropout_t* dropout, int sample_mode,
                             const float* noise, uint64_t seed, int64_t* seq, float* seq_logprobs, float* seq_masks,
                             float* logprobs, float* fc_embedded, float* cpt_feats, void* workspace, size_t workspace_bytes,
                             isc_stream_t stream) {
  ISC_REQUIRE(seq && seq_logprobs && seq_masks, "train_forward_sample: NULL output");
  ISC_REQUIRE(sample_mode == 1 || sample_mode == 2, "train_forward_sample: sample_mode must be 1 (noise) or 2 (seed)");
  ISC_REQUIRE(sample_mode != 1 || noise, "train_forward_sample: sample_mode 1 needs noise");
  SampleOut so;
  so.sample_mode = sample_mode;
  so.noise = noise;
  so.seed = seed;
  so.seq = reinterpret_cast<long long*>(seq);
  so.seq_logprobs = seq_logprobs;
  so.seq_masks = seq_masks;
  return train_forward_impl(dims, packed, precision, mode, fc_feats, att_feats, cpt_words, n_cpt, senti_words, senti_labels, B,
                            nullptr, 0, n_steps, dropout, nullptr, &so, logprobs, fc_embedded, cpt_feats, workspace,
                            workspace_bytes, stream);
}

int isc_train_backward_marks(void* event0, void* event1) {
  g_marks[0] = static_cast<cudaEvent_t>(event0);
  g_marks[1] = static_cast<cudaEvent_t>(event1);
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
  ISC_TRY(zero_pm_pad(w.W5T, H, V, s));
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
    ISC_CUDA(zf(w.dpre_word, (size_t)M * H));
    ISC_CUDA(zf(w.dpre_gates, (size_t)M * G4));
    ISC_CUDA(zf(w.dhproj, (size_t)TM * 3 * H));
    if (w.Vp != V) ISC_CUDA(zf(w.dlogits, (size_t)TM * w.Vp));  // pad columns must read as zero
    ISC_TRY(zero_pm(w.pdhproj, M, s));

    ISC_TRY(zero_pm_pad(w.dlogT, V, TM, s));
    ISC_TRY(zero_pm_pad(w.dg1T, G4, TM, s));
    ISC_TRY(zero_pm_pad(w.dg2T, G4, TM, s));
    ISC_TRY(zero_pm_pad(w.dhprojT, 3 * H, TM, s));
    ISC_TRY(zero_pm_pad(w.X1T, 3 * H, TM, s));
    ISC_TRY(zero_pm_pad(w.X2T, 3 * H, TM, s));
    ISC_TRY(zero_pm_pad(w.csT, 2 * H, TM, s));
    ISC_TRY(zero_pm_pad(w.hLT, H, TM, s));

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
        ab.de_c = w.de_c + r0 * L;      // deferred: d att / d p_att are built once, after the loop
        ab.dctx_c = w.dctx_c + r0 * H;
        ab.alpha_c = pk.alpha_c;
        ab.dalpha_c = g->ca_alpha_w;
      }
      if (tm.sw) {
        ab.sw = f.sw;
        ab.ea_sw = f.p_sw;
        ab.senti_w = w.senti_w + r0 * S;
        ab.de_s = w.de_s + r0 * S;
        ab.dctx_s_out = w.dctx_s + r0 * H;
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

    // ---- attention feature gradients for all steps in one pass per image
    if (tm.att)
      ISC_TRY(launch_attention_bwd_final(T, M, B, (int)L, static_cast<const float*>(f.p_att), w.cont_w, w.de_c, w.dctx_c, w.hproj,
                                         3 * H, 0, nullptr, pk.alpha_c, w.datt, w.dp_att, s));
    if (tm.sw)
      ISC_TRY(launch_attention_bwd_final(T, M, B, (int)S, f.p_sw, w.senti_w, w.de_s, w.dctx_s, w.hproj, 3 * H, H, f.pre_word,
                                         pk.alpha_s, w.dsw, w.dp_sw, s));

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
    // gradients of classifier.*, lang_lstm.* and attention.{h2att, cont2att, senti2att, att_alpha} are final from here
    // on: the data-parallel all-reduce of that (contiguous) tail of the flat gradient may start (isc_train_backward_marks)
    ISC_TRY(record_mark(0, s));

    // ---- hoisted step-invariant terms: pre_gates = [fc | sl] Wpre^T (+ biases, already covered by colsum(dg1))
    ISC_TRY(split_planes(w.dpre_gates, G4, w.pB4.hi, w.pB4.lo, G4, M, G4, s));
    ISC_TRY(zero_pm_pad(w.B4T, G4, M, s));
    ISC_TRY(to_T(w.dpre_gates, G4, M, G4, w.B4T, 0, s));
    ISC_TRY(pgemm(precision, op_of(w.pB4), op_of(w.WpreT), w.dfcsl, 2 * H, M, 2 * H, G4, false, s));  // d [fc | sl]
    ISC_TRY(zero_pm_pad(w.fcT, H, M, s));
    ISC_TRY(zero_pm_pad(w.slT, H, M, s));
    ISC_TRY(to_T(f.fc, H, M, H, w.fcT, 0, s));
    ISC_TRY(to_T(f.sl, H, M, H, w.slT, 0, s));
    ISC_TRY(pgemm(precision, op_of(w.B4T), op_of(w.fcT), g->att_lstm_w_ih + H, 3 * H, G4, H, Bk, true, s));
    ISC_TRY(pgemm(precision, op_of(w.B4T), op_of(w.slT), g->att_lstm_w_ih + 2 * H, 3 * H, G4, H, Bk, true, s));
    ISC_TRY(record_mark(1, s));  // att_lstm.{weight_ih, weight_hh, bias_ih, bias_hh} final
    const float* dsl_extra = nullptr;
    if (tm.sw) {  // label2word(sl), the label term of the sentiment-attention query
      ISC_TRY(launch_colsum_add(w.dpre_word, H, M, H, g->sa_label2word_b, nullptr, s));
      ISC_TRY(zero_pm_pad(w.tBT, H, M, s));
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
      ISC_TRY(zero_pm_pad(w.tBT, H, M, s));
      ISC_TRY(to_T(w.tB1, H, M, H, w.tBT, 0, s));
      ISC_TRY(zero_pm_pad(w.rawfcT, D, M, s));
      ISC_TRY(to_T(fc_feats, D, M, (int)D, w.rawfcT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.tBT), op_of(w.rawfcT), g->fc_embed_w, D, H, (int)D, Bk, true, s));
      // region features: p_att = ReLU(att2att(att)), att = dropout(ReLU(att_embed(raw)))
      const long long BL = (long long)M * L;
      ISC_TRY(launch_relu_mask_bwd(w.dp_att, H, nullptr, 0, static_cast<const float*>(f.p_att), H, 1, nullptr, 0, 1.f, BL, H, w.dz,
                                   H, rd_of(nullptr, 0, w.pdz), s));
      ISC_TRY(launch_colsum_add(w.dz, H, BL, H, g->att2att_b, nullptr, s));
      ISC_TRY(zero_pm_pad(w.dzT, H, BL, s));
      ISC_TRY(zero_pm_pad(w.attT, H, BL, s));
      ISC_TRY(to_T(w.dz, H, BL, H, w.dzT, 0, s));
      ISC_TRY(to_T(static_cast<const float*>(f.att), H, BL, H, w.attT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.dzT), op_of(w.attT), g->att2att_w, H, H, H, (int)pad8(BL), true, s));
      ISC_TRY(pgemm(precision, op_of(w.pdz), op_of(w.Wa2aT), w.datt2, H, (int)BL, H, H, false, s));
      ISC_TRY(launch_relu_mask_bwd(w.datt, H, w.datt2, H, static_cast<const float*>(f.att), H, 0, dropout ? dropout->att : nullptr, H,
                                   dscale, BL, H, w.dz, H, RowDest(), s));
      // tiled batch (dims->att_tile = R rows per image, att_feats holds M / R images): d att_embed_w = (sum of the R tiles'
      // gradients)^T . raw — the tiles share raw, so they are summed before the contraction (and the raw transposes and the
      // GEMM run over M / R images)
      const int Rt = dims->att_tile > 1 ? dims->att_tile : 1;
      ISC_REQUIRE(M % Rt == 0, "att_tile=%d does not divide B=%d", Rt, M);
      const float* dzs = w.dz;
      const long long BLi = BL / Rt;
      if (Rt > 1) {
        ISC_TRY(launch_sum_tiles(w.dz, Rt, L * H, M / Rt, w.datt2, s));  // datt2 was consumed by the pass above
        dzs = w.datt2;
      }
      ISC_TRY(launch_colsum_add(dzs, H, BLi, H, g->att_embed_b, nullptr, s));
      ISC_TRY(zero_pm_pad(w.dzT, H, BLi, s));
      ISC_TRY(to_T(dzs, H, BLi, H, w.dzT, 0, s));
      ISC_TRY(zero_pm_pad(w.rawT, D, BLi, s));
      ISC_TRY(to_T(att_feats, D, BLi, (int)D, w.rawT, 0, s));
      ISC_TRY(pgemm(precision, op_of(w.dzT), op_of(w.rawT), g->att_embed_w, D, H, (int)D, (int)pad8(BLi), true, s));
    }
    if (tm.sw) {
      // sentiment words: p_sw = ReLU(senti2att(sw)), sw = dropout(ReLU(word_embed([PAD | senti_words])))
      const long long BS = (long long)M * S;
      ISC_TRY(launch_relu_mask_bwd(w.dp_sw, H, nullptr, 0, f.p_sw, H, 1, nullptr, 0, 1.f, BS, H, w.dz, H, rd_of(nullptr, 0, w.pdz), s));
      ISC_TRY(launch_colsum_add(w.dz, H, BS, H, g->senti2att_b, nullptr, s));
      ISC_TRY(zero_pm_pad(w.dzT, H, BS, s));
      ISC_TRY(zero_pm_pad(w.attT, H, BS, s));
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
    ISC_TRY(zero_pm_pad(w.tBT, H, M, s));
    ISC_TRY(zero_pm_pad(w.cptT, H, M, s));
    ISC_TRY(to_T(w.tB2, H, M, H, w.tBT, 0, s));
    ISC_TRY(to_T(w.cpt_mean, H, M, H, w.cptT, 0, s));
    ISC_TRY(pgemm(precision, op_of(w.tBT), op_of(w.cptT), g->cpt2fc_w, H, H, H, Bk, true, s));
    ISC_TRY(pgemm(precision, op_of(w.ptB), op_of(w.WcptT), w.tB1, H, M, H, H, false, s));
    ISC_TRY(launch_embed_bwd(reinterpret_cast<const long long*>(cpt_words), n_cpt, M, n_cpt, 0, dims->pad_id, 1, (int)V, pk.emb, w.tB1,
                             H, 1, nullptr, 1.f, 1.0f / (float)n_cpt, g->word_embed, s));
  }
  ISC_TRY(record_mark(0, s));  // marks a call without sequence gradients never reached: everything is final here
  ISC_TRY(record_mark(1, s));
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
