/*
 * isc.h — C ABI of libisc_b200.so: the B200 (sm_100a) caption-decode hot path of InSentiCap.
 *
 * The reference (ezeli/InSentiCap_model) is pure Python/PyTorch and has NO FFI, plugin or
 * operator interface; its boundary for this path is the Python
 * surface of `Captioner` and two reward helpers. Each entry point below names the reference
 * function it replaces (paths relative to the reference root); the Python binding a
 * maintainer would add is a ctypes stub, shown in INTEGRATION.md and implemented in
 * insenticap_model_b200/_lib.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every function returns int: 0 = ok, < 0 = invalid argument (message via
 *     isc_last_error(), thread-local), > 0 = cudaError_t of a failed launch/API call.
 *   - all pointers are DEVICE pointers unless the name ends in `_host`.
 *   - the caller owns every buffer, including the workspace whose size the
 *     `*_workspace_bytes` queries return; the library never allocates device memory,
 *     never synchronises, and enqueues all work on the given stream (CUDA-graph capturable).
 *   - there is no CPU fallback: on a device that is not sm_100 every compute entry point
 *     returns ISC_ERR_DEVICE.
 */
#ifndef ISC_H_
#define ISC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

typedef void* isc_stream_t; /* cudaStream_t */

#define ISC_OK 0
#define ISC_ERR_ARG (-1)
#define ISC_ERR_WORKSPACE (-2)
#define ISC_ERR_DEVICE (-3)
#define ISC_ERR_UNSUPPORTED (-4)

/* arithmetic of the dense contractions (SURVEY.md section 0 / 7 "hard parts" #1) */
#define ISC_PREC_FP32 0   /* fp32 FMA GEMMs on CUDA cores, fp32 features: debugging / strictest parity */
#define ISC_PREC_BF16X3 1 /* tcgen05, split-bf16 hi*hi + lo*hi + hi*lo, fp32 accumulate, fp32 features */
#define ISC_PREC_BF16 2   /* tcgen05, single bf16 pass, fp32 accumulate, bf16 projected features */

/* Model geometry. hidden covers word_emb = feat_emb = rnn_hid = att_hid (opts.py:80-95). */
typedef struct isc_dims {
  int32_t vocab;      /* V */
  int32_t hidden;     /* 512 (only value compiled in) */
  int32_t feat_dim;   /* 2048: fc_feat_dim == att_feat_dim */
  int32_t n_regions;  /* 196 */
  int32_t n_senti;    /* 11 = 10 sentiment words + the prepended PAD (captioner.py:307-309) */
  int32_t n_labels;   /* 3 sentiment categories */
  int32_t pad_id, sos_id, eos_id, unk_id; /* captioner.py:125-128 */
  /* Tiled batches (0 or 1: none). R > 1: every R consecutive batch rows are the same image — the SCST sampled pass with R
   * samples per image — and att_feats holds B / R images, [B/R, n_regions, feat_dim]: ReLU(att_embed(.)) is computed once
   * per image and expanded to the B rows under each row's own dropout mask (exactly what the reference computes on the
   * tiled tensor), and its weight gradient contracts over B/R images after the tiles' gradients have been summed. Every
   * other per-row input (fc_feats, words, labels) stays tiled. Honoured by isc_prologue, isc_train_forward(_sample) and
   * isc_train_backward in ISC_PREC_FP32 / ISC_PREC_BF16X3. */
  int32_t att_tile;
} isc_dims_t;

/* fp32 parameters, one pointer per tensor of Captioner.state_dict() (captioner.py:121-161). */
typedef struct isc_weights {
  const float* word_embed;        /* [V,H]  word_embed.0.weight */
  const float* senti_label_embed; /* [n_labels,H] */
  const float* fc_embed_w;  const float* fc_embed_b;   /* [H,D],[H] */
  const float* cpt2fc_w;    const float* cpt2fc_b;     /* [H,H],[H] */
  const float* att_embed_w; const float* att_embed_b;  /* [H,D],[H] */
  const float* att_lstm_w_ih; const float* att_lstm_w_hh; /* [4H,3H] (in: h_lang|fc|xt), [4H,H] */
  const float* att_lstm_b_ih; const float* att_lstm_b_hh; /* [4H] */
  const float* att2att_w;   const float* att2att_b;    /* [H,H],[H] */
  const float* senti2att_w; const float* senti2att_b;  /* captioner-level senti2att.0 */
  const float* ca_h2att_w;  const float* ca_h2att_b;   /* attention.cont_att.h2att */
  const float* ca_alpha_w;  const float* ca_alpha_b;   /* attention.cont_att.att_alpha [1,H],[1] */
  const float* sa_h2word_w; const float* sa_h2word_b;  /* attention.senti_att.h2word */
  const float* sa_label2word_w; const float* sa_label2word_b;
  const float* sa_alpha_w;  const float* sa_alpha_b;   /* attention.senti_att.word_alpha */
  const float* g_h2att_w;   const float* g_h2att_b;    /* attention.h2att */
  const float* g_cont2att_w; const float* g_cont2att_b;
  const float* g_senti2att_w; const float* g_senti2att_b;
  const float* g_alpha_w;   const float* g_alpha_b;    /* attention.att_alpha */
  const float* lang_lstm_w_ih; const float* lang_lstm_w_hh; /* [4H,2H] (in: att|h_att), [4H,H] */
  const float* lang_lstm_b_ih; const float* lang_lstm_b_hh;
  const float* classifier_w; const float* classifier_b; /* [V,H],[V] */
} isc_weights_t;

/* Step-invariant per-image features (output of isc_prologue, input of the decode calls).
 * NULL members select the reference's mode switches (captioner.py:98-103, :171-172):
 * att == NULL -> seq2seq (sentiment attention only); sw == NULL -> xe (content only). */
typedef struct isc_feats {
  float* fc;        /* [B,H]   ReLU(fc_embed(fc_feats)); in seq2seq mode = cpt_feats */
  void*  att;       /* [B,L,H] ReLU(att_embed(att_feats)); fp32, or bf16 when ISC_PREC_BF16 */
  void*  p_att;     /* [B,L,H] ReLU(att2att(att)); same dtype as att. ISC_PREC_BF16X3: exp(-2 * that) */
  float* sw;        /* [B,S,H] ReLU(word_embed([PAD|senti_words])) */
  float* p_sw;      /* [B,S,H] ReLU(senti2att(sw)). ISC_PREC_BF16X3: exp(-2 * that) */
  float* sl;        /* [B,H]   ReLU(senti_label_embed(labels)) */
  float* pre_gates; /* [B,4H]  W_ih[:,H:2H]·fc + W_ih[:,2H:3H]·sl + b_ih + b_hh (hoisted) */
  float* pre_word;  /* [B,H]   label2word(sl) (hoisted out of SentiAttention.forward) */
  float* cpt_feats; /* [B,H]   ReLU(cpt2fc(mean ReLU(word_embed(cpt_words)))) or NULL */
  /* Optional 16-bit copies for the attention kernel's fast path (decode loops; tensor-core precisions). The attention
   * streams both [B,L,H] tensors once per decode step (the HBM term of the step), so their width is the step's
   * HBM time. All three NULL -> the kernel reads att / p_att above. Written by isc_prologue when non-NULL:
   *   att16   fp16(att)                          (ISC_PREC_BF16X3; ISC_PREC_BF16 keeps reading its bf16 att)
   *   p_att16 fp16(exp(-2 * ReLU(att2att(att))) * 2^15)
   *   feat_flags[b] != 0: image b has a value outside the fast path's exact domain (projected feature > 10, i.e. a
   *           subnormal fp16, or att beyond fp16 range): its CTA reads the full-width tensors instead. */
  void*    att16;      /* [B,L,H] fp16 or NULL */
  void*    p_att16;    /* [B,L,H] fp16 or NULL */
  int32_t* feat_flags; /* [B] or NULL */
} isc_feats_t;

/* Training-mode dropout (nn.Dropout(p), captioner.py:132): uint8 KEEP masks (1 = keep) supplied by the caller so
 * that a run is reproducible and testable against the reference with the same masks; NULL = no dropout on
 * that tensor. Kept values are multiplied by scale = 1 / (1 - p). */
typedef struct isc_dropout {
  const uint8_t* fc;  /* [B,H]   on the fc embedding (:200/:296); seq2seq: on cpt_feats (:250) */
  const uint8_t* att; /* [B,L,H] on the region embedding, before att2att (:210/:304) */
  const uint8_t* sw;  /* [B,S,H] on the sentiment-word embedding, before senti2att (:258/:311) */
  const uint8_t* sl;  /* [B,H]   on the sentiment-label embedding (:214/:262/:315) */
  const uint8_t* out; /* [T,B,H] on h_lang before the classifier, one mask per step (:182) */
  float scale;
} isc_dropout_t;

const char* isc_version(void);
const char* isc_last_error(void);

/* 0 if the current CUDA device is sm_100 (B200), ISC_ERR_DEVICE otherwise. */
int isc_check_device(void);

/* ---- weights ---------------------------------------------------------------------------
 * Fuse / concatenate / split the fp32 parameters into the layouts the kernels read
 * (K-concatenated LSTM and projection matrices, summed biases, bf16 hi/lo planes).
 * Re-run after every optimiser step. Replaces nothing in the reference (it uses the
 * nn.Module tensors directly); it is the price of the fused step. */
size_t isc_packed_weights_bytes(const isc_dims_t* dims, int precision);
int isc_pack_weights(const isc_dims_t* dims, const isc_weights_t* w, int precision,
                     void* packed, size_t packed_bytes, isc_stream_t stream);

/* ---- prologue: Captioner.forward_rl :294-315, forward_xe :198-214, sample :357-376,
 *      forward_seq2seq :247-261 (eval-mode: dropout is the identity) -----------------------
 * fc_feats [B,D], att_feats [B,L,D] fp32 (both NULL with seq2seq != 0);
 * cpt_words int64 [B,n_cpt] or NULL; senti_words int64 [B,S-1] or NULL; senti_labels int64 [B]
 * or NULL. Writes every non-NULL member of *out. */
size_t isc_prologue_workspace_bytes(const isc_dims_t* dims, int precision, int B);
int isc_prologue(const isc_dims_t* dims, const void* packed, int precision,
                 const float* fc_feats, const float* att_feats,
                 const int64_t* cpt_words, int n_cpt,
                 const int64_t* senti_words, const int64_t* senti_labels,
                 int B, int seq2seq, const isc_feats_t* out,
                 void* workspace, size_t workspace_bytes, isc_stream_t stream,
                 const isc_dropout_t* dropout /* NULL: eval mode */);

/* The same with fc_feats / att_feats stored as bf16 (a bf16 feature shard, isc_shard_* below): ISC_PREC_BF16 only, eval
 * mode, not seq2seq. That mode rounds its fp32 inputs to bf16 before the GEMMs, so the result is bit-identical to
 * isc_prologue on the fp32 values the shard was written from, at half the host->device and HBM bytes. */
int isc_prologue_bf16in(const isc_dims_t* dims, const void* packed, int precision,
                        const void* fc_feats_bf16, const void* att_feats_bf16,
                        const int64_t* cpt_words, int n_cpt,
                        const int64_t* senti_words, const int64_t* senti_labels,
                        int B, const isc_feats_t* out,
                        void* workspace, size_t workspace_bytes, isc_stream_t stream);

/* Convert caller-supplied, already-embedded fp32 features (the att_feats / p_att_feats /
 * p_senti_word_feats arguments of Captioner.forward_step, captioner.py:168) into the
 * representation isc_feats_t holds for `precision`: bf16 for ISC_PREC_BF16 (att, p_att), and for
 * ISC_PREC_BF16X3 the PROJECTED tensors (p_att, p_sw; projected != 0) become exp(-2 x), which is
 * what the attention kernel's tanh reads there; everything else is a copy. n elements. */
int isc_convert_features(int precision, int projected, const float* src, void* dst, int64_t n,
                         isc_stream_t stream);
/* fp16 features (a fp16 feature shard, isc_shard_* below; half the host->device bytes of fp32) -> the fp32 tensors
 * isc_prologue reads. Exact: every fp16 value is an fp32 value. n elements, 16-byte aligned buffers. */
int isc_expand_f16(const void* src_f16, float* dst, int64_t n, isc_stream_t stream);

/* Recompute only the hoisted terms (pre_gates, pre_word) from feats->fc / feats->sl:
 * used when the caller supplies already-embedded features (Captioner.forward_step API). */
int isc_hoist(const isc_dims_t* dims, const void* packed, int precision, int B,
              const isc_feats_t* feats, void* workspace, size_t workspace_bytes,
              isc_stream_t stream);

/* ---- one decode step: Captioner.forward_step, captioner.py:168-186 -------------------------
 * M rows; row m uses image m / rows_per_image. it int64 [M]; h_in/c_in/h_out/c_out fp32
 * [2,M,H] (index 0 attention LSTM, 1 language LSTM). logprobs fp32 [M, ld_logprobs >= V]
 * receives log_softmax(classifier(h_lang)). Optional attention weights (NULL to skip):
 * cont_w [M,L], senti_w [M,S], gate_w [M] (Attention._get_weights, captioner.py:83-94). */
size_t isc_decode_workspace_bytes(const isc_dims_t* dims, int precision, int M);
int isc_decode_step(const isc_dims_t* dims, const void* packed, int precision,
                    const isc_feats_t* feats, int rows_per_image, int M,
                    const int64_t* it, const float* h_in, const float* c_in,
                    float* h_out, float* c_out, float* logprobs, int64_t ld_logprobs,
                    float* cont_w, float* senti_w, float* gate_w,
                    void* workspace, size_t workspace_bytes, isc_stream_t stream);

/* ---- batched greedy / sampled decode: Captioner.forward_rl loop, captioner.py:317-349 ------
 * sample_mode 0: argmax (sample_max=1). 1: Gumbel-max with caller noise fp32 [T,B,V]
 * (argmax(logprobs + noise), the reproducible stand-in for torch.multinomial).
 * 2: Gumbel-max with a counter-based generator keyed by (seed, t, row, word).
 * Outputs seq int64 [B,T], seq_logprobs fp32 [B,T], seq_masks fp32 [B,T]; semantics incl. the
 * whole-batch early stop (columns after it stay zero) follow :337-344.
 * Optional weights: cont_w [B,T,L], senti_w [B,T,S], gate_w [B,T] (zero after the stop). */
int isc_decode_greedy(const isc_dims_t* dims, const void* packed, int precision,
                      const isc_feats_t* feats, int B, int T, int sample_mode,
                      const float* noise, uint64_t seed,
                      int64_t* seq, float* seq_logprobs, float* seq_masks,
                      float* cont_w, float* senti_w, float* gate_w,
                      void* workspace, size_t workspace_bytes, isc_stream_t stream,
                      const uint8_t* out_mask /* [T,B,H] keep mask on h_lang (captioner.py:182) or NULL */,
                      float drop_scale);

/* ---- batched beam search: Captioner.sample, captioner.py:378-420, for B images at once -----
 * K = beam_size <= 8. tokens int64 [B,K,T] (EOS included, zero padded), scores fp64 [B,K]
 * (running fp64 sums of fp32 log-probs, :404-407), lengths int32 [B,K]. Beams sorted by score,
 * ties in pool order (parent, then rank) like the reference's stable sort (:409). */
int isc_decode_beam(const isc_dims_t* dims, const void* packed, int precision,
                    const isc_feats_t* feats, int B, int K, int T, int decoding_constraint,
                    int64_t* tokens, double* scores, int32_t* lengths,
                    void* workspace, size_t workspace_bytes, isc_stream_t stream);

/* ---- teacher forcing: forward_xe :216-240 / forward_seq2seq :263-288 with ss_prob = 0 -------
 * inputs int64 [B, ld_inputs]; feeds inputs[:, i] for i < n_steps; logprobs fp32
 * [B, n_steps, V] (the tensor the reference returns). Forward only. */
int isc_teacher_forced(const isc_dims_t* dims, const void* packed, int precision,
                       const isc_feats_t* feats, int B, int n_steps,
                       const int64_t* inputs, int64_t ld_inputs, float* logprobs,
                       void* workspace, size_t workspace_bytes, isc_stream_t stream);

/* ---- training: teacher-forced forward with a tape, hand-written backward, fused clamp + Adam ------------
 * Replaces autograd through Captioner.forward_xe (:194-240) / forward_seq2seq (:242-288), and through the
 * sampled forward_rl pass (models/decoder.py:86-88): REINFORCE re-scores the sampled tokens teacher-forced
 * in ISC_MODE_RL, which is the same computation. ISC_PREC_BF16X3 only.
 *   forward : prologue (+dropout masks) and n_steps decode steps feeding inputs[:, t]; writes
 *             logprobs fp32 [B, n_steps, V] (the reference's return value), the pre-dropout fc embedding
 *             (self.fc_feats) and cpt_feats (self.cpt_feats) when non-NULL, and the activation tape into
 *             the workspace, which must stay untouched until the matching backward.
 *   backward: incoming gradient either dense (dlogprobs [B, n_steps, V]) or as the fused masked-NLL /
 *             REINFORCE form loss = sum_{b,t} coef[b,t] * (-logprobs[b,t,targets[b,t]]) (XECriterion
 *             :427-440, RewardCriterion self_critical/utils.py:169-177; coef [B, n_steps], targets
 *             int64 [B, ld_targets]); d_cpt_feats [B,H] is the gradient wrt cpt_feats (MSE domain-
 *             alignment loss, train_xe.py:163) or NULL. Parameter gradients are ACCUMULATED (+=) into
 *             *grads, laid out like the reference's tensors. */
#define ISC_MODE_XE 0
#define ISC_MODE_SEQ2SEQ 1
#define ISC_MODE_RL 2
typedef struct isc_grads {
  float* word_embed; float* senti_label_embed;
  float* fc_embed_w; float* fc_embed_b; float* cpt2fc_w; float* cpt2fc_b; float* att_embed_w; float* att_embed_b;
  float* att_lstm_w_ih; float* att_lstm_w_hh; float* att_lstm_b_ih; float* att_lstm_b_hh;
  float* att2att_w; float* att2att_b; float* senti2att_w; float* senti2att_b;
  float* ca_h2att_w; float* ca_h2att_b; float* ca_alpha_w; float* ca_alpha_b;
  float* sa_h2word_w; float* sa_h2word_b; float* sa_label2word_w; float* sa_label2word_b;
  float* sa_alpha_w; float* sa_alpha_b;
  float* g_h2att_w; float* g_h2att_b; float* g_cont2att_w; float* g_cont2att_b;
  float* g_senti2att_w; float* g_senti2att_b; float* g_alpha_w; float* g_alpha_b;
  float* lang_lstm_w_ih; float* lang_lstm_w_hh; float* lang_lstm_b_ih; float* lang_lstm_b_hh;
  float* classifier_w; float* classifier_b;
} isc_grads_t; /* same order as isc_weights_t */
/* Scheduled sampling (captioner.py:219-228): at steps t >= 1 a row with uniform[t,b] < prob is fed a word drawn
 * from the previous step's distribution (Gumbel-max on the previous log-probs: noise [n_steps,B,V] if given, else
 * a counter-based generator keyed by seed) instead of inputs[b,t]. NULL / prob 0: plain teacher forcing. */
typedef struct isc_sched_sampling {
  float prob;
  const float* uniform; /* [n_steps, B] in [0,1) */
  const float* noise;   /* [n_steps, B, V] Gumbel noise or NULL */
  uint64_t seed;
} isc_sched_sampling_t;
size_t isc_train_workspace_bytes(const isc_dims_t* dims, int precision, int B, int n_steps);
int isc_train_forward(const isc_dims_t* dims, const void* packed, int precision, int mode,
                      const float* fc_feats, const float* att_feats, const int64_t* cpt_words, int n_cpt,
                      const int64_t* senti_words, const int64_t* senti_labels, int B,
                      const int64_t* inputs, int64_t ld_inputs, int n_steps, const isc_dropout_t* dropout,
                      const isc_sched_sampling_t* sched_sampling,
                      float* logprobs, float* fc_embedded, float* cpt_feats,
                      void* workspace, size_t workspace_bytes, isc_stream_t stream);
/* isc_train_forward with the decoder FREE-RUNNING instead of teacher-forced: the sampled pass of Captioner.forward_rl
 * (models/captioner.py:317-349, sample_max = 0) recorded on the tape, i.e. what models/decoder.py:86-88 runs under autograd.
 * Step 0 is fed <SOS>; every step draws its token by Gumbel-max on the step's logits (sample_mode 1: noise [n_steps,B,V]
 * given; 2: counter-based generator keyed by seed), writes seq / seq_logprobs / seq_masks [B,n_steps] with the bookkeeping
 * of isc_decode_greedy (mask column = unfinished, token *= unfinished, log-prob unmasked, columns after the step at which
 * every row has finished stay zero) and feeds the token to the next step. logprobs [B,n_steps,V] and the workspace then
 * hold the tape of exactly those tokens: isc_train_backward is called with inputs = [<SOS>, seq[:, :-1]], targets = seq.
 * One pass instead of a sampling decode followed by a teacher-forced re-scoring of its tokens. */
int isc_train_forward_sample(const isc_dims_t* dims, const void* packed, int precision, int mode,
                             const float* fc_feats, const float* att_feats, const int64_t* cpt_words, int n_cpt,
                             const int64_t* senti_words, const int64_t* senti_labels, int B, int n_steps,
                             const isc_dropout_t* dropout, int sample_mode, const float* noise, uint64_t seed,
                             int64_t* seq, float* seq_logprobs, float* seq_masks,
                             float* logprobs, float* fc_embedded, float* cpt_feats,
                             void* workspace, size_t workspace_bytes, isc_stream_t stream);
int isc_train_backward(const isc_dims_t* dims, const void* packed, int precision, int mode,
                       const float* fc_feats, const float* att_feats, const int64_t* cpt_words, int n_cpt,
                       const int64_t* senti_words, const int64_t* senti_labels, int B,
                       const int64_t* inputs, int64_t ld_inputs, int n_steps, const isc_dropout_t* dropout,
                       const float* logprobs, const float* dlogprobs, const int64_t* targets,
                       int64_t ld_targets, const float* coef, const float* d_cpt_feats,
                       const isc_grads_t* grads, void* workspace, size_t workspace_bytes, isc_stream_t stream);
/* Arms up to two CUDA events (cudaEvent_t, or NULL) for the NEXT isc_train_backward call of this host thread. The call
 * records event0 on its stream once the gradients of classifier.*, lang_lstm.* and attention.{h2att, cont2att,
 * senti2att, att_alpha} are final, and event1 once att_lstm.* is final — before the hoisted terms and the prologue
 * layers are differentiated — so that the caller can all-reduce those gradients (the path's one collective,
 * bucketed in reverse layer order) under the rest of the backward. Both are recorded at the latest when the call ends. */
int isc_train_backward_marks(void* event0, void* event1);
/* Element-wise clamp to +-clip (train_xe.py:19-23; clip <= 0 disables) then torch.optim.Adam's update, fused over a
 * flat fp32 buffer; grads are first multiplied by grad_scale (1 / world_size after a summing all-reduce). */
int isc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float clip, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int step, float grad_scale, isc_stream_t stream);

/* ---- image sentiment detector: SentimentDetector.forward / .sample (models/sentiment_detector.py:30-60) -----------
 * The step before the caption decode at inference (Detector.sample, models/decoder.py:186); SURVEY.md 8(f) row f2.
 * att_feats fp32 [B,14,14,feat_dim] -> output fp32 [B,n_cls] (the MLP's logits, forward()'s first result), maps fp32
 * [B,14,14] (forward()'s second result), labels int64 [B] (arg-max of softmax(output), replaced by neu_idx where the
 * winning probability is below threshold, :49-52) and scores fp32 [B]. eval() semantics (dropout = identity); exactly
 * two 3x3 convolutions (settings['sentiment_convs_num'] = 2). The convolution weights are passed ALREADY RESHAPED to
 * [C_out][9 * C_in] with K index ((ky * 3 + kx) * C_in + c_in), i.e. conv.weight.permute(0, 2, 3, 1).reshape(C_out, -1);
 * isc_senti_pack splits them into bf16 hi/lo planes. w1x1 [n_cls, feat_dim/4]; out_w [n_fc][n_cls][n_cls]. */
size_t isc_senti_packed_bytes(int feat_dim);
int isc_senti_pack(int feat_dim, const float* conv0_w, const float* conv1_w, void* packed, size_t packed_bytes,
                   isc_stream_t stream);
size_t isc_senti_workspace_bytes(int feat_dim, int B);
int isc_senti_detect(int feat_dim, int n_cls, const void* packed, const float* conv0_b, const float* conv1_b,
                     const float* w1x1, const float* b1x1, const float* out_w, const float* out_b, int n_fc,
                     const float* att_feats, int B, float threshold, int neu_idx, float* output, float* maps,
                     int64_t* labels, float* scores, void* workspace, size_t workspace_bytes, isc_stream_t stream);

/* ---- sentence sentiment classifier, inference: SentenceSentimentClassifier.forward (models/sent_senti_cls.py:38-56) --
 * The model behind get_cls_reward (self_critical/utils.py:120-151) and the XE pseudo-labels (train_xe.py:155-158);
 * SURVEY.md 8(f) row f3. seqs int64 [B, ld_seqs] (first T columns used), lengths int32 [B] (>= 1) -> pred fp32
 * [B, n_cls] and the per-word weights fp32 [B, T] (zero past each caption's length; the reference returns the first
 * max(lengths) columns). eval() semantics. Parameters: word_embed [V,H], the nn.LSTM's weight_ih_l0 / weight_hh_l0
 * [4H,H] and biases, excitation.0 / .2 and sent_senti_cls.0 [H,H], sent_senti_cls.3 [n_cls,H] with their biases. */
size_t isc_sentcls_packed_bytes(int vocab, int n_cls);
int isc_sentcls_pack(int vocab, int n_cls, const float* word_embed, const float* w_ih, const float* w_hh,
                     const float* b_ih, const float* b_hh, const float* exc0_w, const float* exc0_b,
                     const float* exc2_w, const float* exc2_b, const float* cls0_w, const float* cls0_b,
                     const float* cls3_w, const float* cls3_b, void* packed, size_t packed_bytes, isc_stream_t stream);
size_t isc_sentcls_workspace_bytes(int B, int T);
int isc_sentcls_forward(int vocab, int n_cls, const void* packed, const int64_t* seqs, int64_t ld_seqs,
                        const int32_t* lengths, int B, int T, float* pred, float* att_weights,
                        void* workspace, size_t workspace_bytes, isc_stream_t stream);

/* ---- feature shards: the input side of the path (dataloader.py:164-204 opens one HDF5 file per item) -----------------
 * One flat file of fixed-size records (fc_feats [D] then att_feats [L][D] per image; fp32, bit-exact with the
 * reference's arrays, or bf16), mmap'ed once. HOST-ONLY calls: no GPU is touched. isc_shard_gather copies the records
 * of indices[0..n) (any order, repeats allowed) into fc_dst [n][D] and att_dst [n][L][D] (either may be NULL) in the
 * shard's dtype with n_threads host threads; pass pinned buffers and follow with one cudaMemcpyAsync per tensor.
 * Layout: 64-byte header {"ISCFEAT1", u32 version = 1, u32 dtype, u32 D, u32 L, u64 n, u64 names_bytes, u64 data_offset,
 * u64 record_bytes}, n NUL-terminated names in record order, zero padding to data_offset, records. */
#define ISC_SHARD_F32 0
#define ISC_SHARD_BF16 1
#define ISC_SHARD_F16 2 /* IEEE half, round-to-nearest-even: 11 significant bits, exactly representable by the split-bf16 GEMM operands */
typedef void* isc_shard_t;
int isc_shard_write(const char* path, int dtype, int feat_dim, int n_regions, int64_t n_images,
                    const char* const* names, const float* fc_feats, const float* att_feats);
int isc_shard_open(const char* path, isc_shard_t* out);
int isc_shard_close(isc_shard_t shard);
int isc_shard_info(isc_shard_t shard, int64_t* n_images, int* feat_dim, int* n_regions, int* dtype);
int64_t isc_shard_find(isc_shard_t shard, const char* name);       /* record index, -1 if absent */
const char* isc_shard_name(isc_shard_t shard, int64_t index);      /* NULL if out of range */
int isc_shard_gather(isc_shard_t shard, const int64_t* indices, int64_t n, void* fc_dst, void* att_dst, int n_threads);
/* Zero-staging variant for shards that fit in host RAM: isc_shard_pin page-locks the records — the mapping in place
 * (cudaHostRegister, read-only) where the platform allows, else the file is read once into cudaHostAlloc'ed memory —
 * and isc_shard_copy_to_device then issues one async copy per record and tensor into DEVICE buffers fc_dst [n][D] /
 * att_dst [n][L][D] (shard dtype) on `stream`. */
int isc_shard_pin(isc_shard_t shard);
int isc_shard_copy_to_device(isc_shard_t shard, const int64_t* indices, int64_t n, void* fc_dst, void* att_dst,
                             isc_stream_t stream);

/* ---- dense contraction on its own (validation / profiling of the tensor-core kernel) -------
 * C[M,N] = act(A[M,K] · W[N,K]^T + bias[N]), fp32 in/out; act 0 none, 1 ReLU, 2 tanh.
 * With ISC_PREC_BF16X3 / ISC_PREC_BF16 the operands are split to bf16 planes in the
 * workspace and multiplied by the tcgen05 kernel. */
size_t isc_gemm_workspace_bytes(int precision, int M, int N, int K);
int isc_gemm_tn(int precision, const float* A, int64_t lda, const float* W, int64_t ldw,
                const float* bias, float* C, int64_t ldc, int M, int N, int K, int act,
                void* workspace, size_t workspace_bytes, isc_stream_t stream);

/* ---- CIDEr-D reward: self_critical/utils.py:56-83 + ciderD_scorer.py:13-192 ------------------
 * Captions are id arrays. A reference set is ref_tokens int32 [R, ref_ld] with ref_lens [R]
 * (words AFTER utils._array_to_str: SOS stripped, cut at EOS, EOS appended) and
 * img_offsets int32 [N+1] (image i owns refs img_offsets[i] .. img_offsets[i+1]-1).
 * The document-frequency table is an open-addressing hash of EXACT packed n-gram keys
 * (16 bit per token, so vocab <= 65534): table_slots a power of two >= 2 x distinct n-grams. */
size_t isc_cider_table_bytes(int64_t table_slots);
int isc_cider_build_df(const int32_t* ref_tokens, const int32_t* ref_lens, int32_t ref_ld,
                       const int32_t* img_offsets, int32_t n_images,
                       void* table, int64_t table_slots, int32_t* overflow_flag,
                       isc_stream_t stream);
/* hyps int64 [n_hyp, T] raw decoder output (leading SOS stripped, cut at first EOS, EOS
 * appended — utils._array_to_str); hyp_img int32 [n_hyp] indexes img_offsets. log_n_docs =
 * log(number of images the table was built from). scores fp64 [n_hyp] = CIDEr-D x 10. */
int isc_cider_score(const void* table, int64_t table_slots, double log_n_docs,
                    const int64_t* hyps, int32_t T, const int32_t* hyp_img, int32_t n_hyp,
                    const int32_t* ref_tokens, const int32_t* ref_lens, int32_t ref_ld,
                    const int32_t* img_offsets, int32_t sos_id, int32_t eos_id,
                    double* scores, isc_stream_t stream);
/* rewards fp64 [B,T] = (scores[b] - scores[B + b]) repeated over T (utils.py:81-82). */
int isc_self_critical_reward(const double* scores, int32_t B, int32_t T, double* rewards,
                             isc_stream_t stream);
/* n-gram term frequencies of one hypothesis, for parity tests: keys uint64 [<=64],
 * counts int32 [<=64], n_out int32 [1]. */
int isc_cider_ngram_counts(const int64_t* hyp, int32_t T, int32_t sos_id, int32_t eos_id,
                           uint64_t* keys, int32_t* counts, int32_t* n_out, isc_stream_t stream);

/* ---- measurement hooks (bench.py) -----------------------------------------------------------
 * isc_launch_count: kernels this library has launched in this process (monotonic).
 * Profiling: when enabled, every launch is bracketed by CUDA events on its own stream and its
 * algorithmic work is booked per kernel class; isc_profile_read synchronises the recorded events
 * and returns the class totals since the last isc_profile_reset. Do not enable inside a CUDA-graph
 * capture. work = flops (GEMM classes: 2*M*N*K*passes) or HBM bytes (all other classes). */
#define ISC_K_GEMM_TC 0
#define ISC_K_GEMM_SIMT 1
#define ISC_K_ATTENTION 2
#define ISC_K_LSTM 3
#define ISC_K_POINTWISE 4 /* embed/pack, gate mix, prologue gathers, plane split, fills */
#define ISC_K_SELECT 5    /* log_softmax, greedy pick, beam expansion + merge */
#define ISC_K_CIDER 6
#define ISC_K_TRAIN 7     /* backward-pass glue kernels, clamp + Adam */
#define ISC_K_NUM 8
uint64_t isc_launch_count(void);
int isc_profile_enable(int on);
int isc_profile_reset(void);
int isc_profile_read(int kernel_class, double* total_ms, double* total_work, int64_t* launches);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* ISC_H_ */
