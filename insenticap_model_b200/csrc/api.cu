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

#include "engine.cuh"

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
  if (precision != ISC_PREC_FP32) {
    // W1b = [h_lang_prev | h_att_prev] columns; xt_gates = ReLU(E) . W_ih[:, 2H:3H]^T through the tensor-core GEMM
    ISC_TRY(copy_block(p.W1b.f32, 2 * H, w->att_lstm_w_ih, 3 * H, G4, H, s));
    ISC_TRY(copy_block(p.W1b.f32 + H, 2 * H, w->att_lstm_w_hh, H, G4, H, s));
    ISC_TRY(split_planes_gate_interleaved(p.W1b.f32, 2 * H, p.W1b.hi, p.W1b.lo, 2 * H, H, 2 * H, s));
    ISC_TRY(split_planes_gate_interleaved(p.W4.f32, 3 * H, p.W4g.hi, p.W4g.lo, 3 * H, H, 3 * H, s));
    ISC_TRY(launch_embed_rows(nullptr, V, 1, 0, 0, V, p.emb, [&] {
      RowDest r;
      r.hi = p.erelu_hi;
      r.lo = p.erelu_lo;
      r.ldp = H;
      return r;
    }(), s));
    Operand a, wx;
    a.hi = p.erelu_hi;
    a.lo = p.erelu_lo;
    a.ldp = H;
    wx.hi = p.W1.hi + H;  // the xt columns of W1 = [h_lang | xt | h_att]
    wx.lo = p.W1.lo ? p.W1.lo + H : nullptr;
    wx.ldp = 3 * H;
    Dest d;
    d.f32 = p.xt_gates;
    d.ld = G4;
    ISC_TRY(gemm_tc(a, wx, d, V, G4, H, precision == ISC_PREC_BF16X3 ? 3 : 1, Epilogue(), s));
  }
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

int isc_prologue_bf16in(const isc_dims_t* dims, const void* packed, int precision, const void* fc_feats_bf16,
                        const void* att_feats_bf16, const int64_t* cpt_words, int n_cpt, const int64_t* senti_words,
                        const int64_t* senti_labels, int B, const isc_feats_t* out, void* workspace, size_t workspace_bytes,
                        isc_stream_t stream) {
  return run_prologue(dims, packed, precision, static_cast<const float*>(fc_feats_bf16), static_cast<const float*>(att_feats_bf16),
                      cpt_words, n_cpt, senti_words, senti_labels, B, 0, out, workspace, workspace_bytes, stream, nullptr,
                      nullptr, true);
}

}  // extern "C"

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
  // before the first word the K beams of an image are identical (captioner.py:378-381 starts from ONE state): step 0
  // runs on one row per image, and the merge kernel fans its candidates out to the K beams
  const bool compact0 = fused && K > 1;
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
    io.skip_pack = fused && t > 0;  // the merge kernel of step t - 1 has packed this step's recurrent operand rows
    if (compact0 && t == 0) {
      io.state_rows = M;
      ISC_TRY(run_step(c, w, B, 1, io));
    } else {
      ISC_TRY(run_step(c, w, M, K, io));
    }
    BeamParams bp;
    bp.compact0 = compact0 ? 1 : 0;
    if (fused && t + 1 < T) {
      bp.h_state = io.h_out;
      bp.state_rows = M;
      bp.x1 = rowdest(nullptr, 0, w.pX1, 3 * H);
      bp.x2 = rowdest(nullptr, 0, w.pX2, 3 * H);
    }
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

int isc_expand_f16(const void* src_f16, float* dst, int64_t n, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_REQUIRE(src_f16 && dst && n >= 0, "bad expand_f16 arguments");
  ISC_REQUIRE((reinterpret_cast<uintptr_t>(src_f16) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "expand_f16: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n8 = (n + 7) / 8;
  int blocks = (int)((n8 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps(ISC_K_POINTWISE, (double)n * 6.0, s);
  expand_f16_kernel<<<blocks, 256, 0, s>>>(static_cast<const __half*>(src_f16), dst, n);
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
