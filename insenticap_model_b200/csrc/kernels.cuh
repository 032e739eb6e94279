// Internal launcher declarations shared by the kernel translation units and api.cu.
#pragma once
#include "common.cuh"

namespace isc {

// Row-major destination written by the element-wise kernels: fp32 and/or bf16 hi/lo planes
// (the planes are what the tcgen05 GEMM that consumes the row reads through TMA).
struct RowDest {
  float* f32 = nullptr;
  long long ld = 0;
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  long long ldp = 0;

#ifdef __CUDACC__
  __device__ __forceinline__ void store4(long long row, int col, float4 v) const {
    if (f32) *reinterpret_cast<float4*>(f32 + row * ld + col) = v;
    if (hi) {
      __nv_bfloat16 h[4], l[4];
      split_bf16(v.x, h[0], l[0]);
      split_bf16(v.y, h[1], l[1]);
      split_bf16(v.z, h[2], l[2]);
      split_bf16(v.w, h[3], l[3]);
      *reinterpret_cast<uint2*>(hi + row * ldp + col) = *reinterpret_cast<uint2*>(h);
      if (lo) *reinterpret_cast<uint2*>(lo + row * ldp + col) = *reinterpret_cast<uint2*>(l);
    }
  }
  __device__ __forceinline__ void store2(long long row, int col, float2 v) const {
    if (f32) *reinterpret_cast<float2*>(f32 + row * ld + col) = v;
    if (hi) {
      __nv_bfloat16 h[2], l[2];
      split_bf16(v.x, h[0], l[0]);
      split_bf16(v.y, h[1], l[1]);
      *reinterpret_cast<unsigned int*>(hi + row * ldp + col) = *reinterpret_cast<unsigned int*>(h);
      if (lo) *reinterpret_cast<unsigned int*>(lo + row * ldp + col) = *reinterpret_cast<unsigned int*>(l);
    }
  }
#endif
};

struct AttnParams {
  int R = 1, L = 0, S = 0;
  const float* hproj = nullptr;  // [M, ld_hproj]: [h2att(h) | h2word(h) | gate h2att(h)]
  long long ld_hproj = 0;
  const void* att = nullptr;    // [B,L,H] fp32 or bf16; null -> no content attention (seq2seq)
  const void* p_att = nullptr;
  // fast path (isc_feats_t::att16 / p_att16 / feat_flags): fp16 copies; p_att16 = exp(-2 p) * 2^15. A CTA whose image is
  // flagged (value outside the fp16 path's exact domain) reads att / p_att instead.
  const void* att16 = nullptr;
  const void* p_att16 = nullptr;
  const int* flags = nullptr;
  const float* sw = nullptr;    // [B,S,H]; null -> no sentiment attention (xe)
  const float* p_sw = nullptr;
  const float* pre_word = nullptr;  // [B,H] label2word(sl)
  const float* alpha_c = nullptr;   // [H]
  const float* alpha_s = nullptr;   // [H]
  RowDest cont_dst;
  int cont_col = 0;
  RowDest senti_dst;
  int senti_col = 0;
  float* cont_w = nullptr;  // [M, ld_cont_w] optional softmax weights
  long long ld_cont_w = 0;
  float* senti_w = nullptr;
  long long ld_senti_w = 0;
};

int launch_embed_pack(const long long* it, const int* parent, const float* h_in, int M, int V, const float* emb,
                      RowDest x1, RowDest x2, cudaStream_t stream);
int launch_lstm_pointwise(const float* gates, const int* parent, const float* c_prev, float* h_out, float* c_out,
                          RowDest extra, int extra_col, int M, cudaStream_t stream, const unsigned char* mask = nullptr,
                          float scale = 1.0f);
int launch_attention(const AttnParams& p, int B, int precision, cudaStream_t stream);
int launch_gate_mix(const float* g3, const float* cs, const float* alpha, const float* alpha_b, RowDest ctx,
                    float* gate_w, long long ld_gate_w, int M, cudaStream_t stream);
int launch_embed_rows(const long long* ids, long long groups, long long n_per_group, int prepend_pad, int pad_id, int V,
                      const float* emb, RowDest dst, cudaStream_t stream);
int launch_embed_mean(const long long* ids, int B, int n, int V, const float* emb, RowDest dst, cudaStream_t stream);
int launch_fill(float* p, long long n, float v, cudaStream_t stream);

// ---- selection (kernels_select.cu) ---------------------------------------------------------
// in-place log_softmax over rows of `x` [M, ld] (first V columns)
int launch_log_softmax(float* x, long long ld, int M, int V, cudaStream_t stream);

struct GreedyParams {
  const float* logits = nullptr;  // [B, ld]
  long long ld = 0;
  int B = 0, V = 0, T = 0, t = 0;
  int sample_mode = 0;            // 0 argmax, 1 external noise, 2 counter-based Gumbel
  const float* noise = nullptr;   // [B, V] for this step (mode 1)
  unsigned long long seed = 0;
  int eos_id = 2;
  long long* it = nullptr;        // [B] next input token (written)
  int* unfinished = nullptr;      // [B] (read/write)
  int* alive_count = nullptr;     // [T] number of unfinished rows after step t
  const float* rec = nullptr;     // fused argmax path (sample_mode 0): [B][np][sel_rec(4)], logits is null then
  int np = 0;
  long long* seq = nullptr;       // [B,T]
  float* seq_logprobs = nullptr;  // [B,T]
  float* seq_masks = nullptr;     // [B,T]
};
int launch_greedy_select(const GreedyParams& p, cudaStream_t stream);

struct BeamParams {
  const float* logits = nullptr;  // [B*K, ld]
  long long ld = 0;
  int B = 0, K = 0, V = 0, T = 0, t = 0;
  int constraint = 1;
  int pad_id = 0, sos_id = 1, eos_id = 2, unk_id = 3;
  // beam state, ping-pong by step parity: *_in read, *_out written
  const int* tok_in = nullptr;  // [B,K,T]
  int* tok_out = nullptr;
  const int* len_in = nullptr;  // [B,K]
  int* len_out = nullptr;
  const double* score_in = nullptr;  // [B,K]
  double* score_out = nullptr;
  const int* alive_in = nullptr;  // [B,K]
  int* alive_out = nullptr;
  long long* it = nullptr;  // [B*K] last word per beam: read as "last", written for the next step
  int* parent = nullptr;    // [B*K] absolute state row each new beam continues from (written)
  // scratch: per-row candidates published by the row CTAs, per-image ticket counter (zeroed once)
  // fused path: records written by the logits GEMM epilogue (common.cuh LogitsSelect); logits is null then
  const float* rec = nullptr;  // [B*K][np][sel_rec(k_sel)]
  int np = 0, k_sel = 4;
  float* cand_lp = nullptr;  // [B*K, 8]
  int* cand_word = nullptr;  // [B*K, 8]
  int* cand_count = nullptr; // [B*K]
  int* ticket = nullptr;     // [B]
  // step 0 ran on ONE row per image (all K beams of an image are identical before the first word): its records and
  // state rows are indexed by image, and the parents written for step 1 point at those rows
  int compact0 = 0;
  // fused path: the merge kernel also builds the NEXT step's recurrent operand rows — X1[m] = [h_lang | h_att] and
  // X2[m, 2H:3H] = h_lang of the state row parent[m] it has just chosen (what embed_pack does at step 0), which
  // saves a launch per step. h_state = this step's output state [2][state_rows][H]; null = leave it to embed_pack.
  const float* h_state = nullptr;
  long long state_rows = 0;
  RowDest x1, x2;
};
int launch_beam_select(const BeamParams& p, cudaStream_t stream);
int launch_beam_init(long long* it, int* alive, int* len, double* score, int* parent, int B, int K, int sos_id,
                     cudaStream_t stream);
int launch_beam_finalize(const int* tok, const int* len, const double* score, long long* tokens_out, double* scores_out,
                         int* lengths_out, int B, int K, int T, cudaStream_t stream);
int launch_greedy_init(long long* it, int* unfinished, int B, int sos_id, cudaStream_t stream);
int launch_ss_select(const float* logp_prev, long long ld_logp, const long long* truth, long long ld_truth,
                     const float* uniform, float prob, const float* noise, unsigned long long seed, int t, int B, int V,
                     long long* it, cudaStream_t stream);
// weights outputs must read as zero after the whole-batch early stop
int launch_zero_if_stopped(float* p, long long row_stride, int row_len, int B, const int* alive_count, int t,
                           cudaStream_t stream);

// ---- training / backward (train.cu) ---------------------------------------------------------------------
struct AttnBwdParams {
  int L = 0, S = 0;
  const float* dcs = nullptr; long long ld_dcs = 0; int cont_col = 0, senti_col = 0;
  const float* hproj = nullptr; long long ld_hproj = 0;
  const float* pre_word = nullptr;
  const float* att = nullptr; const float* ea_att = nullptr;
  const float* sw = nullptr; const float* ea_sw = nullptr;
  const float* cont_w = nullptr; const float* senti_w = nullptr;
  const float* alpha_c = nullptr; const float* alpha_s = nullptr;
  float* datt = nullptr; float* dp_att = nullptr; float* dsw = nullptr; float* dp_sw = nullptr;
  RowDest dhproj;
  float* dpre_word = nullptr;
  float* dalpha_c = nullptr; float* dalpha_s = nullptr;
  // deferred accumulation (datt == null): this step's softmax gradients and d ctx are kept for launch_attention_bwd_final
  float* de_c = nullptr; float* de_s = nullptr;            // [M,L], [M,S]
  float* dctx_c = nullptr; float* dctx_s_out = nullptr;    // [M,H]
};
int launch_attention_bwd_final(int T, int M, int B, int n_items, const float* ea, const float* w_all, const float* de_all,
                               const float* dctx_all, const float* hproj, long long ld_hproj, int q_col, const float* pre_word,
                               const float* alpha, float* dfeat, float* dp, cudaStream_t s);
int launch_logsoftmax_bwd(const float* logp, const float* dlogp, const long long* target, long long ld_target,
                          const float* coef, int T, int M, int V, float* dlogits, long long ld_out, cudaStream_t s);
int launch_lstm_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh_a, long long ld_a,
                    const unsigned char* mask, float scale, const float* dh_b, long long ld_b, const float* dh_c,
                    long long ld_c, float* dc_carry, float* dgates, RowDest planes, int M, cudaStream_t s);
int launch_gate_bwd(const float* dctx, long long ld_dctx, const float* cs, const float* g3, const float* gate_w,
                    const float* alpha, float* dcs, RowDest dpre3, int dpre3_col, float* dalpha, float* dalpha_b, int M,
                    cudaStream_t s);
int launch_attention_bwd(const AttnBwdParams& a, int M, cudaStream_t s);
int launch_embed_bwd(const long long* ids, long long ld_ids, long long groups, long long ids_per_group, int prepend_pad,
                     int pad_id, int skip_pad, int V, const float* emb, const float* g, long long ld_g,
                     long long g_rows_per_group, const unsigned char* mask, float scale, float gscale, float* demb,
                     cudaStream_t s);
int launch_relu_mask_bwd(const float* a, long long ld_a, const float* b, long long ld_b, const float* gate, long long ld_gate,
                         int gate_is_exp, const unsigned char* mask, long long ld_mask, float scale, long long rows, int cols,
                         float* out, long long ld_out, RowDest planes, cudaStream_t s);
int launch_apply_mask(float* x, long long ld, const unsigned char* mask, float scale, long long rows, int cols, RowDest planes,
                      cudaStream_t s);
int launch_expand_mask(const float* src, int R, long long per, long long n_dst_images, const unsigned char* mask, float scale,
                       float* dst, int cols, RowDest planes, cudaStream_t s);
int launch_sum_tiles(const float* src, int R, long long per, long long n_images, float* dst, cudaStream_t s);
int launch_colsum_add(const float* src, long long ld, long long rows, int cols, float* dst, float* dst2, cudaStream_t s);
int launch_sum_steps(const float* src, int T, long long stride_t, long long n, float* dst, cudaStream_t s);
int launch_add2d(float* dst, long long ld_dst, const float* src, long long ld_src, long long rows, int cols, cudaStream_t s);
int launch_split_transpose(const float* src, long long ld_src, long long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo,
                           long long ld_dst, long long col0, cudaStream_t s);
int launch_transpose_bf16(const __nv_bfloat16* src, long long ld_src, long long rows, int cols, __nv_bfloat16* dst,
                          long long ld_dst, long long col0, cudaStream_t s);
int launch_adam_clamp(float* p, const float* g, float* m, float* v, long long n, float clip, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int step, float grad_scale, cudaStream_t s);

}  // namespace isc
