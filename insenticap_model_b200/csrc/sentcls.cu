// Sentence sentiment classifier, inference (SentenceSentimentClassifier.forward, /root/reference/models/sent_senti_cls.py:38-56):
// the model behind the classifier reward of the RL step (self_critical/utils.py:120-151) and the XE pseudo-labels
// (train_xe.py:155-158, models/decoder.py:131-134); SURVEY.md section 8(f) row f3. ReLU word embeddings -> one-layer LSTM
// over the caption (packed-sequence semantics: outputs past a caption's length are zero) -> squeeze-excitation word weights
// -> weighted sentence feature -> two-layer classifier. The LSTM runs on the decode path's fused gate GEMM (tcgen05, LSTM
// cell in the epilogue), the excitation / classifier layers on the plain tensor-core GEMM; split-bf16 throughout.
#include "engine.cuh"

namespace isc {
namespace {

// outm[b*T + t] = t < len[b] ? h_all[t][b] : 0   (pad_packed_sequence), fp32 + operand planes
__global__ void __launch_bounds__(128) mask_rows_kernel(const float* __restrict__ h_all, const int* __restrict__ lens, int B, int T,
                                                        float* __restrict__ outm, RowDest planes) {
  const long long row = blockIdx.x;  // b * T + t
  const int b = (int)(row / T), t = (int)(row - (long long)b * T);
  const int c = threadIdx.x * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < lens[b]) v = *reinterpret_cast<const float4*>(h_all + ((long long)t * B + b) * H + c);
  *reinterpret_cast<float4*>(outm + row * H + c) = v;
  planes.store4(row, c, v);
}

// squeeze[b][t] = t < len[b] ? mean_j e2[b*T + t][j] : 0   (the second pack / pad round trip + AdaptiveAvgPool1d)
__global__ void __launch_bounds__(128) squeeze_kernel(const float* __restrict__ e2, const int* __restrict__ lens, int T,
                                                      float* __restrict__ squeeze) {
  __shared__ float red[4];
  const long long row = blockIdx.x;
  const int b = (int)(row / T), t = (int)(row - (long long)b * T);
  const float4 v = *reinterpret_cast<const float4*>(e2 + row * H + threadIdx.x * 4);
  float s = warp_sum(v.x + v.y + v.z + v.w);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) squeeze[row] = t < lens[b] ? (red[0] + red[1] + red[2] + red[3]) / (float)H : 0.f;
}

// sent_feats[b] = sum_t squeeze[b][t] * outm[b*T + t]   (squeeze_res.bmm(out), sent_senti_cls.py:53)
__global__ void __launch_bounds__(128) sent_feat_kernel(const float* __restrict__ outm, const float* __restrict__ squeeze, int T,
                                                        RowDest dst) {
  const long long b = blockIdx.x;
  const int c = threadIdx.x * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < T; ++t) {
    const float w = squeeze[b * T + t];
    const float4 v = *reinterpret_cast<const float4*>(outm + (b * T + t) * H + c);
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
  }
  dst.store4(b, c, acc);
}

struct SentClsPacked {
  Mat Wl;      // [4H][2H] = [W_ih | W_hh]
  Mat We0, We2, Wc0;  // [H][H]
  Mat Wc3;     // [n_cls][H]
  float *bl, *be0, *be2, *bc0, *bc3, *emb;
  size_t total;
};
SentClsPacked carve_sentcls_packed(int V, int n_cls, void* base) {
  Bump b(base);
  SentClsPacked p;
  auto mat = [&](Mat& m, int rows, int cols) {
    m.rows = rows;
    m.cols = cols;
    m.f32 = b.take<float>((size_t)rows * cols);
    m.hi = b.take<bf16>((size_t)rows * cols);
    m.lo = b.take<bf16>((size_t)rows * cols);
  };
  mat(p.Wl, G4, 2 * H);
  mat(p.We0, H, H);
  mat(p.We2, H, H);
  mat(p.Wc0, H, H);
  mat(p.Wc3, n_cls, H);
  p.bl = b.take<float>(G4);
  p.be0 = b.take<float>(H);
  p.be2 = b.take<float>(H);
  p.bc0 = b.take<float>(H);
  p.bc3 = b.take<float>(n_cls);
  p.emb = b.take<float>((size_t)V * H);
  p.total = (b.off + 255) & ~size_t(255);
  return p;
}
struct SentClsWs {
  long long* it;   // [B]
  Planes pX[2];    // [B][2H] LSTM operand [x_t | h_{t-1}], ping-pong: step t reads one while its epilogue fills the other
  float *h_all;    // [T][B][H]
  float *c[2];     // [B][H] ping-pong
  float *outm, *e2, *squeeze, *tmpB;  // [B*T][H], [B*T][H], [B*T], [B][H]
  Planes poutm, pe1, pfeat, pc1;      // planes of the GEMM operands
  size_t total;
};
SentClsWs carve_sentcls_ws(int B, int T, void* base) {
  Bump b(base);
  SentClsWs w;
  const size_t m = (size_t)B, bt = m * T;
  auto planes = [&](Planes& p, size_t n) {
    p.hi = b.take<bf16>(n);
    p.lo = b.take<bf16>(n);
  };
  w.it = b.take<long long>(m);
  planes(w.pX[0], m * 2 * H);
  planes(w.pX[1], m * 2 * H);
  w.h_all = b.take<float>(bt * H);
  w.c[0] = b.take<float>(m * H);
  w.c[1] = b.take<float>(m * H);
  w.outm = b.take<float>(bt * H);
  w.e2 = b.take<float>(bt * H);
  w.squeeze = b.take<float>(bt);
  w.tmpB = b.take<float>(m * H);
  planes(w.poutm, bt * H);
  planes(w.pe1, bt * H);
  planes(w.pfeat, m * H);
  planes(w.pc1, m * H);
  w.total = (b.off + 255) & ~size_t(255);
  return w;
}

}  // namespace
}  // namespace isc

using namespace isc;

extern "C" {

size_t isc_sentcls_packed_bytes(int vocab, int n_cls) {
  if (vocab <= 0 || n_cls <= 0) return 0;
  return carve_sentcls_packed(vocab, n_cls, nullptr).total;
}

int isc_sentcls_pack(int vocab, int n_cls, const float* word_embed, const float* w_ih, const float* w_hh, const float* b_ih,
                     const float* b_hh, const float* exc0_w, const float* exc0_b, const float* exc2_w, const float* exc2_b,
                     const float* cls0_w, const float* cls0_b, const float* cls3_w, const float* cls3_b, void* packed,
                     size_t packed_bytes, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_REQUIRE(vocab > 0 && n_cls > 0 && word_embed && w_ih && w_hh && b_ih && b_hh && exc0_w && exc0_b && exc2_w && exc2_b &&
                  cls0_w && cls0_b && cls3_w && cls3_b && packed,
              "bad sentcls_pack arguments");
  SentClsPacked p = carve_sentcls_packed(vocab, n_cls, packed);
  if (packed_bytes < p.total) {
    set_error("sentcls packed buffer too small: %zu < %zu", packed_bytes, p.total);
    return ISC_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ISC_TRY(copy_block(p.Wl.f32, 2 * H, w_ih, H, G4, H, s));
  ISC_TRY(copy_block(p.Wl.f32 + H, 2 * H, w_hh, H, G4, H, s));
  ISC_TRY(copy_block(p.We0.f32, H, exc0_w, H, H, H, s));
  ISC_TRY(copy_block(p.We2.f32, H, exc2_w, H, H, H, s));
  ISC_TRY(copy_block(p.Wc0.f32, H, cls0_w, H, H, H, s));
  ISC_TRY(copy_block(p.Wc3.f32, H, cls3_w, H, n_cls, H, s));
  ISC_TRY(add_vec(p.bl, b_ih, b_hh, G4, s));
  ISC_TRY(add_vec(p.be0, exc0_b, nullptr, H, s));
  ISC_TRY(add_vec(p.be2, exc2_b, nullptr, H, s));
  ISC_TRY(add_vec(p.bc0, cls0_b, nullptr, H, s));
  ISC_TRY(add_vec(p.bc3, cls3_b, nullptr, n_cls, s));
  ISC_TRY(copy_block(p.emb, H, word_embed, H, vocab, H, s));
  ISC_TRY(split_planes_gate_interleaved(p.Wl.f32, 2 * H, p.Wl.hi, p.Wl.lo, 2 * H, H, 2 * H, s));  // fused LSTM GEMM layout
  const Mat* mats[] = {&p.We0, &p.We2, &p.Wc0, &p.Wc3};
  for (const Mat* m : mats) ISC_TRY(finish_mat(*m, ISC_PREC_BF16X3, s));
  return 0;
}

size_t isc_sentcls_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  return carve_sentcls_ws(B, T, nullptr).total;
}

int isc_sentcls_forward(int vocab, int n_cls, const void* packed, const int64_t* seqs, int64_t ld_seqs, const int32_t* lengths,
                        int B, int T, float* pred, float* att_weights, void* workspace, size_t workspace_bytes,
                        isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_REQUIRE(vocab > 0 && n_cls > 0 && packed && seqs && lengths && B > 0 && T > 0 && ld_seqs >= T && pred && att_weights,
              "bad sentcls_forward arguments");
  SentClsPacked p = carve_sentcls_packed(vocab, n_cls, const_cast<void*>(packed));
  SentClsWs w = carve_sentcls_ws(B, T, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("sentcls workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long BT = (long long)B * T;
  // h_0 = c_0 = 0
  ISC_CUDA(cudaMemsetAsync(w.pX[0].hi, 0, (size_t)B * 2 * H * sizeof(bf16), s));
  ISC_CUDA(cudaMemsetAsync(w.pX[0].lo, 0, (size_t)B * 2 * H * sizeof(bf16), s));
  ISC_CUDA(cudaMemsetAsync(w.c[0], 0, (size_t)B * H * sizeof(float), s));
  for (int t = 0; t < T; ++t) {
    ISC_CUDA(cudaMemcpy2DAsync(w.it, sizeof(long long), seqs + t, ld_seqs * sizeof(long long), sizeof(long long), B,
                               cudaMemcpyDeviceToDevice, s));
    // x_t = ReLU(E[w_t]) into the first half of the operand; the fused epilogue writes h_t into the second half
    const Planes& cur = w.pX[t & 1];
    const Planes& nxt = w.pX[(t + 1) & 1];
    ISC_TRY(launch_embed_rows(w.it, B, 1, 0, 0, vocab, p.emb, rowdest(nullptr, 0, cur, 2 * H), s));
    LstmEpilogue le;
    le.c_prev = w.c[t & 1];
    le.c_out = w.c[(t + 1) & 1];
    le.h_out = w.h_all + (size_t)t * B * H;
    le.x_hi = nxt.hi;  // h_t becomes the second half of the NEXT step's operand
    le.x_lo = nxt.lo;
    le.ldx = 2 * H;
    le.x_col = H;
    ISC_TRY(gemm_tc_lstm(operand(nullptr, 0, cur, 2 * H), p.Wl.op(), B, 2 * H, 3, p.bl, nullptr, 0, 1, le, s));
  }
  {
    ProfScope ps(ISC_K_POINTWISE, (double)BT * H * 4.0 * 3, s);
    mask_rows_kernel<<<(unsigned)BT, 128, 0, s>>>(w.h_all, lengths, B, T, w.outm, rowdest(nullptr, 0, w.poutm, H));
    ISC_LAUNCH_CHECK();
  }
  // excitation: Linear -> ReLU -> Linear -> Sigmoid (sent_senti_cls.py:24-29)
  {
    Epilogue e;
    e.bias = p.be0;
    e.act = ACT_RELU;
    Dest d;
    d.hi = w.pe1.hi;
    d.lo = w.pe1.lo;
    d.ldp = H;
    ISC_TRY(gemm_tc(operand(nullptr, 0, w.poutm, H), p.We0.op(), d, (int)BT, H, H, 3, e, s));
    Epilogue e2;
    e2.bias = p.be2;
    e2.act = ACT_SIGMOID;
    Dest d2;
    d2.f32 = w.e2;
    d2.ld = H;
    ISC_TRY(gemm_tc(operand(nullptr, 0, w.pe1, H), p.We2.op(), d2, (int)BT, H, H, 3, e2, s));
  }
  {
    ProfScope ps(ISC_K_POINTWISE, (double)BT * H * 4.0 * 2, s);
    squeeze_kernel<<<(unsigned)BT, 128, 0, s>>>(w.e2, lengths, T, att_weights);
    ISC_LAUNCH_CHECK();
    sent_feat_kernel<<<B, 128, 0, s>>>(w.outm, att_weights, T, rowdest(nullptr, 0, w.pfeat, H));
    ISC_LAUNCH_CHECK();
  }
  // classifier: Linear -> ReLU -> (dropout) -> Linear (:31-36)
  {
    Epilogue e;
    e.bias = p.bc0;
    e.act = ACT_RELU;
    Dest d;
    d.hi = w.pc1.hi;
    d.lo = w.pc1.lo;
    d.ldp = H;
    ISC_TRY(gemm_tc(operand(nullptr, 0, w.pfeat, H), p.Wc0.op(), d, B, H, H, 3, e, s));
    Epilogue e3;
    e3.bias = p.bc3;
    Dest d3;
    d3.f32 = pred;
    d3.ld = n_cls;
    ISC_TRY(gemm_tc(operand(nullptr, 0, w.pc1, H), p.Wc3.op(), d3, B, n_cls, H, 3, e3, s));
  }
  return 0;
}

}  // extern "C"
