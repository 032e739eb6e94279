// Image sentiment detector (SentimentDetector.forward / .sample, /root/reference/models/sentiment_detector.py:30-60),
// the step right before the caption decode at inference (models/decoder.py:186) and SURVEY.md section 8(f) row f2:
// two 3x3 convolutions (2048 -> 1024 -> 512 channels over the 14x14 region grid, 9.2 GFLOP per image), ReLU, a 1x1
// convolution to the sentiment classes, global average pooling, a small MLP, softmax, thresholded arg-max.
// The convolutions run as tcgen05 GEMMs (gemm_tc_conv3x3: one GEMM per convolution over a zero-bordered 16x16 grid, the
// nine taps are row-shifted K segments) in split-bf16 (bf16x3); the raw fp32 features are split inside the first GEMM.
#include "engine.cuh"

namespace isc {
namespace {

// att_feats fp32 [B,14,14,C] -> zero-bordered grid [B,16,16,C]; one block per grid row of C floats
__global__ void __launch_bounds__(128) pad16_kernel(const float* __restrict__ src, float* __restrict__ dst, int C) {
  const long long row = blockIdx.x;  // image * 256 + gy * 16 + gx
  const int g = (int)(row & 255), gy = g >> 4, gx = g & 15;
  const long long img = row >> 8;
  float4* d = reinterpret_cast<float4*>(dst + row * C);
  if (gy == 0 || gy == 15 || gx == 0 || gx == 15) {
    for (int i = threadIdx.x; i < C / 4; i += 128) d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* s = reinterpret_cast<const float4*>(src + ((img * 14 + (gy - 1)) * 14 + (gx - 1)) * C);
  for (int i = threadIdx.x; i < C / 4; i += 128) d[i] = __ldg(s + i);
}

constexpr int MAX_CLS = 8;
// Per image: 1x1 convolution on every pixel (a warp per pixel), average pooling, the output MLP, softmax, the
// class-weighted sentiment map and the thresholded label. feat: ReLU features on the 16x16 grid [B*256][F].
__global__ void __launch_bounds__(256) senti_head_kernel(const float* __restrict__ feat, int F, const float* __restrict__ w1x1,
                                                         const float* __restrict__ b1x1, const float* __restrict__ out_w,
                                                         const float* __restrict__ out_b, int n_fc, int n_cls, float threshold,
                                                         int neu_idx, float* __restrict__ output, float* __restrict__ maps,
                                                         long long* __restrict__ labels, float* __restrict__ scores) {
  __shared__ float sf[MAX_CLS][196];
  __shared__ float pooled[MAX_CLS], prob[MAX_CLS];
  const long long img = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pos = warp; pos < 196; pos += 8) {
    const int y = pos / 14, x = pos - y * 14;
    const float* f = feat + ((img << 8) + (y + 1) * 16 + (x + 1)) * F;
    float acc[MAX_CLS];
#pragma unroll
    for (int k = 0; k < MAX_CLS; ++k) acc[k] = 0.f;
    for (int c = lane * 4; c < F; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(f + c);
      for (int k = 0; k < n_cls; ++k) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w1x1 + (long long)k * F + c));
        acc[k] += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
      }
    }
    for (int k = 0; k < n_cls; ++k) {
      const float t = warp_sum(acc[k]);
      if (lane == 0) sf[k][pos] = t + b1x1[k];
    }
  }
  __syncthreads();
  if (warp < n_cls) {  // global average pooling, a warp per class
    float t = 0.f;
    for (int pos = lane; pos < 196; pos += 32) t += sf[warp][pos];
    t = warp_sum(t);
    if (lane == 0) pooled[warp] = t / 196.0f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x[MAX_CLS], y[MAX_CLS];
    for (int k = 0; k < n_cls; ++k) x[k] = pooled[k];
    for (int l = 0; l < n_fc; ++l) {  // nn.Sequential of Linear(n_cls, n_cls) layers, no activation in between
      for (int o = 0; o < n_cls; ++o) {
        float t = out_b[l * n_cls + o];
        for (int i = 0; i < n_cls; ++i) t += out_w[(l * n_cls + o) * n_cls + i] * x[i];
        y[o] = t;
      }
      for (int k = 0; k < n_cls; ++k) x[k] = y[k];
    }
    float mx = -INFINITY;
    for (int k = 0; k < n_cls; ++k) {
      output[img * n_cls + k] = x[k];
      mx = fmaxf(mx, x[k]);
    }
    float sum = 0.f;
    for (int k = 0; k < n_cls; ++k) {
      prob[k] = expf(x[k] - mx);
      sum += prob[k];
    }
    int best = 0;
    for (int k = 0; k < n_cls; ++k) {
      prob[k] /= sum;
      if (prob[k] > prob[best]) best = k;
    }
    scores[img] = prob[best];
    labels[img] = prob[best] < threshold ? neu_idx : best;  // sentiment_detector.py:49-52
  }
  __syncthreads();
  for (int pos = threadIdx.x; pos < 196; pos += 256) {
    float t = 0.f;
    for (int k = 0; k < n_cls; ++k) t += prob[k] * sf[k][pos];
    maps[img * 196 + pos] = t;
  }
}

struct SentiPacked {
  bf16 *w0_hi, *w0_lo, *w1_hi, *w1_lo;  // [C/2][9C], [C/4][9C/2]
  size_t total;
};
SentiPacked carve_senti_packed(int C, void* base) {
  Bump b(base);
  SentiPacked p;
  p.w0_hi = b.take<bf16>((size_t)(C / 2) * 9 * C);
  p.w0_lo = b.take<bf16>((size_t)(C / 2) * 9 * C);
  p.w1_hi = b.take<bf16>((size_t)(C / 4) * 9 * (C / 2));
  p.w1_lo = b.take<bf16>((size_t)(C / 4) * 9 * (C / 2));
  p.total = (b.off + 255) & ~size_t(255);
  return p;
}
constexpr int SENTI_CHUNK = 74;  // images per pass: 148 row tiles x 4 / x 2 wide column tiles = whole waves on 148 SMs
struct SentiWs {
  float* pad;          // [chunk*256][C] zero-bordered fp32 input
  bf16 *a1_hi, *a1_lo; // [chunk*256][C/2] first convolution's output planes (border rows zero)
  float* feat;         // [chunk*256][C/4] ReLU features
  int chunk;
  size_t total;
};
SentiWs carve_senti_ws(int C, int B, void* base) {
  Bump b(base);
  SentiWs w;
  w.chunk = B < SENTI_CHUNK ? B : SENTI_CHUNK;
  const size_t rows = (size_t)w.chunk * 256;
  w.pad = b.take<float>(rows * C);
  w.a1_hi = b.take<bf16>(rows * (C / 2));
  w.a1_lo = b.take<bf16>(rows * (C / 2));
  w.feat = b.take<float>(rows * (C / 4));
  w.total = (b.off + 255) & ~size_t(255);
  return w;
}

}  // namespace
}  // namespace isc

using namespace isc;

extern "C" {

size_t isc_senti_packed_bytes(int feat_dim) {
  if (feat_dim <= 0 || feat_dim % 256 != 0) return 0;
  return carve_senti_packed(feat_dim, nullptr).total;
}

int isc_senti_pack(int feat_dim, const float* conv0_w, const float* conv1_w, void* packed, size_t packed_bytes,
                   isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_REQUIRE(feat_dim > 0 && feat_dim % 256 == 0 && conv0_w && conv1_w && packed, "bad senti_pack arguments");
  SentiPacked p = carve_senti_packed(feat_dim, packed);
  if (packed_bytes < p.total) {
    set_error("senti packed buffer too small: %zu < %zu", packed_bytes, p.total);
    return ISC_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = feat_dim;
  ISC_TRY(split_planes(conv0_w, 9LL * C, p.w0_hi, p.w0_lo, 9LL * C, C / 2, 9 * C, s));
  ISC_TRY(split_planes(conv1_w, 9LL * (C / 2), p.w1_hi, p.w1_lo, 9LL * (C / 2), C / 4, 9 * (C / 2), s));
  return 0;
}

size_t isc_senti_workspace_bytes(int feat_dim, int B) {
  if (feat_dim <= 0 || feat_dim % 256 != 0 || B <= 0) return 0;
  return carve_senti_ws(feat_dim, B, nullptr).total;
}

int isc_senti_detect(int feat_dim, int n_cls, const void* packed, const float* conv0_b, const float* conv1_b,
                     const float* w1x1, const float* b1x1, const float* out_w, const float* out_b, int n_fc,
                     const float* att_feats, int B, float threshold, int neu_idx, float* output, float* maps,
                     int64_t* labels, float* scores, void* workspace, size_t workspace_bytes, isc_stream_t stream) {
  ISC_TRY(check_device());
  ISC_REQUIRE(feat_dim > 0 && feat_dim % 256 == 0 && n_cls >= 1 && n_cls <= MAX_CLS && n_fc >= 0 && B > 0,
              "bad senti_detect geometry (feat_dim %% 256, 1 <= classes <= %d)", MAX_CLS);
  ISC_REQUIRE(packed && conv0_b && conv1_b && w1x1 && b1x1 && (n_fc == 0 || (out_w && out_b)) && att_feats && output && maps &&
                  labels && scores,
              "NULL senti_detect argument");
  const int C = feat_dim;
  SentiPacked p = carve_senti_packed(C, const_cast<void*>(packed));
  SentiWs w = carve_senti_ws(C, B, workspace);
  if (!workspace || workspace_bytes < w.total) {
    set_error("senti workspace too small: %zu < %zu", workspace_bytes, w.total);
    return ISC_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Operand w0, w1, a1;
  w0.hi = p.w0_hi;
  w0.lo = p.w0_lo;
  w0.ldp = 9LL * C;
  w1.hi = p.w1_hi;
  w1.lo = p.w1_lo;
  w1.ldp = 9LL * (C / 2);
  a1.hi = w.a1_hi;
  a1.lo = w.a1_lo;
  a1.ldp = C / 2;
  for (int b0 = 0; b0 < B; b0 += w.chunk) {
    const int nb = B - b0 < w.chunk ? B - b0 : w.chunk;
    const int rows = nb * 256;
    {
      ProfScope ps(ISC_K_POINTWISE, (double)nb * (196.0 + 256.0) * C * 4.0, s);
      pad16_kernel<<<rows, 128, 0, s>>>(att_feats + (long long)b0 * 196 * C, w.pad, C);
      ISC_LAUNCH_CHECK();
    }
    // conv_0: 2048 -> 1024, no activation (sentiment_detector.py:13-16); border rows zeroed: they are conv_1's padding
    Epilogue e0;
    e0.bias = conv0_b;
    Dest d0;
    d0.hi = w.a1_hi;
    d0.lo = w.a1_lo;
    d0.ldp = C / 2;
    ISC_TRY(gemm_tc_conv3x3(w.pad, C, Operand(), w0, d0, rows, C / 2, C, 3, e0, 1, s));
    // conv_1: 1024 -> 512, then (dropout: identity in eval) ReLU (:17-18)
    Epilogue e1;
    e1.bias = conv1_b;
    e1.act = ACT_RELU;
    Dest d1;
    d1.f32 = w.feat;
    d1.ld = C / 4;
    ISC_TRY(gemm_tc_conv3x3(nullptr, 0, a1, w1, d1, rows, C / 4, C / 2, 3, e1, 0, s));
    {
      ProfScope ps(ISC_K_POINTWISE, (double)nb * 196.0 * (C / 4) * 4.0, s);
      senti_head_kernel<<<nb, 256, 0, s>>>(w.feat, C / 4, w1x1, b1x1, out_w, out_b, n_fc, n_cls, threshold, neu_idx,
                                           output + (long long)b0 * n_cls, maps + (long long)b0 * 196, reinterpret_cast<long long*>(labels) + b0, scores + b0);
      ISC_LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // extern "C"
