// fp32 CUDA-core GEMM (ISC_PREC_FP32): C = act(A · W^T + bias + rowadd + addmat).
// Plain fp32 FMA accumulation — the strictest-parity arithmetic for the dense contractions of
// /root/reference/models/captioner.py (nn.Linear / nn.LSTMCell); the tensor-core kernel in
// gemm_tc.cu is the throughput path. Also holds the fp32 -> bf16 hi/lo plane splitter.
#include "common.cuh"

namespace isc {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, long long lda,
                                                        const float* __restrict__ W, long long ldw, int M, int N, int K,
                                                        const float* __restrict__ bias, const float* __restrict__ rowadd,
                                                        long long ld_rowadd, int rows_per_group,
                                                        const float* __restrict__ addmat, long long ld_addmat, int act,
                                                        float* __restrict__ C, long long ldc, __nv_bfloat16* hi,
                                                        __nv_bfloat16* lo, long long ldp) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    // 64 rows x 16 k: 1024 elements per operand, 4 per thread, k fastest for coalescing
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = threadIdx.x + i * 256;
      int r = e >> 4, k = e & 15;
      float a = 0.f, w = 0.f;
      if (k0 + k < K) {
        if (m0 + r < M) a = A[(long long)(m0 + r) * lda + k0 + k];
        if (n0 + r < N) w = W[(long long)(n0 + r) * ldw + k0 + k];
      }
      As[k][r] = a;
      Ws[k][r] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (rowadd) v += rowadd[(long long)(m / rows_per_group) * ld_rowadd + n];
      if (addmat) v += addmat[(long long)m * ld_addmat + n];
      v = apply_act(v, act);
      if (C) C[(long long)m * ldc + n] = v;
      if (hi) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        hi[(long long)m * ldp + n] = h;
        if (lo) lo[(long long)m * ldp + n] = l;
      }
    }
  }
}

// 4 independent 16-byte loads in flight per thread: the kernel also runs with a deliberately small grid
// (beside a persistent GEMM, see run_prologue) and must still cover the HBM latency
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ src, long long ld_src,
                                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                    long long ld_dst, long long rows, int cols, int gate_hh) {
  // gate_hh > 0: the source is an LSTM weight [4 * gate_hh][cols] in nn.LSTMCell's gate-major row order; the planes are
  // written GATE-INTERLEAVED — destination row r = (16-unit group) * 64 + gate * 16 + unit — so that the rows of all four
  // gates of a group of hidden units are contiguous and one TMA box fetches a fused-LSTM tile (gemm_tc_lstm)
  const int cols4 = cols >> 2;  // cols % 4 == 0 enforced by the host
  const long long total = rows * cols4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 v[4];
    long long r[4];
    int c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        r[u] = i / cols4;
        c[u] = (int)(i - r[u] * cols4) * 4;
        const long long rs = gate_hh > 0 ? ((r[u] >> 4) & 3) * gate_hh + (r[u] >> 6) * 16 + (r[u] & 15) : r[u];
        v[u] = __ldg(reinterpret_cast<const float4*>(src + rs * ld_src + c[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * stride < total) {
        __nv_bfloat16 h[4], l[4];
        split_bf16(v[u].x, h[0], l[0]);
        split_bf16(v[u].y, h[1], l[1]);
        split_bf16(v[u].z, h[2], l[2]);
        split_bf16(v[u].w, h[3], l[3]);
        *reinterpret_cast<uint2*>(hi + r[u] * ld_dst + c[u]) = *reinterpret_cast<uint2*>(h);
        if (lo) *reinterpret_cast<uint2*>(lo + r[u] * ld_dst + c[u]) = *reinterpret_cast<uint2*>(l);
      }
    }
  }
}

}  // namespace

int gemm_simt(const Operand& A, const Operand& W, const Dest& C, int M, int N, int K, const Epilogue& ep,
              cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  ISC_REQUIRE(A.f32 && W.f32, "gemm_simt: fp32 operands missing");
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
  ProfScope ps(ISC_K_GEMM_SIMT, 2.0 * M * N * K, stream);
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(A.f32, A.ld, W.f32, W.ld, M, N, K, ep.bias, ep.rowadd, ep.ld_rowadd,
                                             ep.rows_per_group > 0 ? ep.rows_per_group : 1, ep.addmat, ep.ld_addmat,
                                             ep.act, C.f32, C.ld, C.hi, C.lo, C.ldp);
  ISC_LAUNCH_CHECK();
  return 0;
}

int split_planes(const float* src, int64_t ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t ld_dst, int64_t rows,
                 int cols, cudaStream_t stream, int max_blocks) {
  if (rows <= 0 || cols <= 0) return 0;
  ISC_REQUIRE(cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
              "split_planes: cols/ld must be multiples of 4 and src 16-byte aligned");
  long long total = rows * (cols / 4);
  int blocks = (int)((total + 1023) / 1024);
  if (blocks < 1) blocks = 1;
  const int cap = max_blocks > 0 ? max_blocks : 148 * 8;
  if (blocks > cap) blocks = cap;
  ProfScope ps(ISC_K_POINTWISE, (double)rows * cols * (4.0 + 2.0 + (lo ? 2.0 : 0.0)), stream);
  split_kernel<<<blocks, 256, 0, stream>>>(src, ld_src, hi, lo, ld_dst, rows, cols, 0);
  ISC_LAUNCH_CHECK();
  return 0;
}

int split_planes_gate_interleaved(const float* src, long long ld_src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long ld_dst,
                                  int hidden, int cols, cudaStream_t stream) {
  ISC_REQUIRE(src && hi && hidden > 0 && hidden % 16 == 0 && cols > 0 && cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0,
              "split_planes_gate_interleaved: bad arguments");
  const long long rows = 4LL * hidden, total = rows * (cols >> 2);
  long long blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope ps(ISC_K_POINTWISE, (double)rows * cols * (4.0 + 2.0 + (lo ? 2.0 : 0.0)), stream);
  split_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, ld_src, hi, lo, ld_dst, rows, cols, hidden);
  ISC_LAUNCH_CHECK();
  return 0;
}

}  // namespace isc
