// Token selection on the vocabulary logits: log_softmax (captioner.py:183), the greedy / sampled
// pick of forward_rl (captioner.py:328-344) and the beam expansion + pooled stable top-K of
// Captioner.sample (captioner.py:394-409). One CTA per row (per image for the beam), 128-bit
// coalesced reads of the logits row, warp-shuffle + shared-memory block reductions.
#include <math_constants.h>

#include "kernels.cuh"

namespace isc {
namespace {

constexpr int NT = 256;
constexpr int KMAX = 8;

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) r += red[i];
  return r;
}
// (value desc, index asc) argmax over the block; every thread gets the winner
__device__ __forceinline__ void block_argmax(float& v, int& idx, float* redv, int* redi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) {
      v = ov;
      idx = oi;
    }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    redv[threadIdx.x >> 5] = v;
    redi[threadIdx.x >> 5] = idx;
  }
  __syncthreads();
  v = redv[0];
  idx = redi[0];
#pragma unroll
  for (int i = 1; i < NT / 32; ++i) {
    float ov = redv[i];
    int oi = redi[i];
    if (ov > v || (ov == v && oi < idx)) {
      v = ov;
      idx = oi;
    }
  }
}

// row statistics with 128-bit loads: max over the first V entries
__device__ __forceinline__ float row_max(const float* row, int V, float* red) {
  float mx = -CUDART_INF_F;
  const int v4 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? (V >> 2) : 0;
  for (int i = threadIdx.x; i < v4; i += NT) {
    float4 x = reinterpret_cast<const float4*>(row)[i];  // plain load: the row may be rewritten in place
    mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
  }
  for (int i = v4 * 4 + threadIdx.x; i < V; i += NT) mx = fmaxf(mx, row[i]);
  return block_max(mx, red);
}
__device__ __forceinline__ float row_sumexp(const float* row, int V, float mx, float* red) {
  float s = 0.f;
  const int v4 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? (V >> 2) : 0;
  for (int i = threadIdx.x; i < v4; i += NT) {
    float4 x = reinterpret_cast<const float4*>(row)[i];  // plain load: the row may be rewritten in place
    s += expf(x.x - mx) + expf(x.y - mx) + expf(x.z - mx) + expf(x.w - mx);
  }
  for (int i = v4 * 4 + threadIdx.x; i < V; i += NT) s += expf(row[i] - mx);
  return block_sum(s, red);
}

__global__ void __launch_bounds__(NT) log_softmax_kernel(float* __restrict__ x, long long ld, int V) {
  __shared__ float red[NT / 32];
  float* row = x + (long long)blockIdx.x * ld;
  const float mx = row_max(row, V, red);
  const float ls = logf(row_sumexp(row, V, mx, red));
  for (int i = threadIdx.x; i < V; i += NT) row[i] = (row[i] - mx) - ls;
}

__device__ __forceinline__ float gumbel_counter(unsigned long long seed, unsigned long long ctr) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (ctr + 1ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  float u = ((float)(z >> 41) + 0.5f) * (1.0f / 8388608.0f);  // (0,1), 23 random bits
  return -logf(-logf(u));
}

__global__ void __launch_bounds__(NT) greedy_select_kernel(GreedyParams p) {
  __shared__ float red[NT / 32];
  __shared__ float redv[NT / 32];
  __shared__ int redi[NT / 32];
  const int b = blockIdx.x;
  const int t = p.t;
  if (t > 0 && p.alive_count[t - 1] == 0) return;  // whole-batch early stop (captioner.py:343-344)
  const float* row = p.logits + (long long)b * p.ld;
  const int V = p.V;
  // pass 1: max and (for plain argmax) its first index
  float bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += NT) {
    float x = row[i];
    if (x > bv) {
      bv = x;
      bi = i;
    }
  }
  block_argmax(bv, bi, redv, redi);
  const float mx = bv;
  const float ls = logf(row_sumexp(row, V, mx, red));
  int pick = bi;
  float lp = (mx - mx) - ls;
  if (p.sample_mode != 0) {
    float sv = -CUDART_INF_F;
    int si = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += NT) {
      float g = (p.sample_mode == 1)
                    ? p.noise[(long long)b * V + i]
                    : gumbel_counter(p.seed, ((unsigned long long)t * p.B + b) * (unsigned long long)V + i);
      float s = ((row[i] - mx) - ls) + g;
      if (s > sv) {
        sv = s;
        si = i;
      }
    }
    block_argmax(sv, si, redv, redi);
    pick = si;
    lp = (row[pick] - mx) - ls;
  }
  if (threadIdx.x == 0) {
    const int unf = p.unfinished[b];
    const long long tok = unf ? pick : 0;  // finished rows emit PAD (captioner.py:338)
    p.seq[(long long)b * p.T + t] = tok;
    p.seq_logprobs[(long long)b * p.T + t] = lp;  // written unmasked (captioner.py:340)
    p.seq_masks[(long long)b * p.T + t] = unf ? 1.f : 0.f;
    const int unf2 = unf && (tok != p.eos_id);
    p.unfinished[b] = unf2;
    p.it[b] = tok;
    if (unf2) atomicAdd(p.alive_count + t, 1);
  }
}

// Pool the candidates of an image's rows with its carried finished beams and keep the K best by fp64 score,
// stably (captioner.py:404-409); then write the new beam state. Called by every thread of ONE block per image,
// after the rows' candidates (cand_lp / cand_word / cand_count) are visible.
// s_lp / s_word / s_cnt: the rows' candidates in SHARED memory (fused merge kernel: [K][KMAX] / [K]); null = read the
// global scratch the unfused row CTAs published them to.
__device__ __forceinline__ void beam_pool(const BeamParams& p, int b, const float* s_lp = nullptr, const int* s_word = nullptr,
                                          const int* s_cnt = nullptr) {
  __shared__ int sel_parent[KMAX], sel_word[KMAX], sel_n;
  const int K = p.K, t = p.t, T = p.T;
  if (threadIdx.x == 0) {
    double pool_score[KMAX * KMAX + KMAX];
    int pool_parent[KMAX * KMAX + KMAX], pool_word[KMAX * KMAX + KMAX];
    long long lasts[KMAX];
    int n = 0;
    for (int kk = 0; kk < K; ++kk) {
      const int mm = b * K + kk;
      lasts[kk] = p.it[mm];
      if (!p.alive_in[mm]) continue;
      const double base = p.score_in[mm];
      if (t > 0 && lasts[kk] == p.eos_id) {
        pool_score[n] = base;
        pool_parent[n] = kk;
        pool_word[n] = -1;
        ++n;
      } else {
        const int cnt = s_cnt ? s_cnt[kk] : __ldcg(p.cand_count + mm);
        for (int r = 0; r < cnt; ++r) {
          // python-float running sum of fp32 log-probs (captioner.py:404-407)
          pool_score[n] = base + (double)(s_lp ? s_lp[kk * KMAX + r] : __ldcg(p.cand_lp + mm * KMAX + r));
          pool_parent[n] = kk;
          pool_word[n] = s_word ? s_word[kk * KMAX + r] : __ldcg(p.cand_word + mm * KMAX + r);
          ++n;
        }
      }
    }
    // stable top-K of the pool by score (python sorted(reverse=True) keeps pool order on ties)
    unsigned long long taken = 0ULL;
    int j = 0;
    for (; j < K && j < n; ++j) {
      int best = -1;
      for (int i = 0; i < n; ++i) {
        if ((taken >> i) & 1ULL) continue;
        if (best < 0 || pool_score[i] > pool_score[best]) best = i;
      }
      taken |= 1ULL << best;
      const int kk = pool_parent[best], w = pool_word[best];
      sel_parent[j] = kk;
      sel_word[j] = w;
      p.score_out[b * K + j] = pool_score[best];
      p.alive_out[b * K + j] = 1;
      p.len_out[b * K + j] = p.len_in[b * K + kk] + (w >= 0 ? 1 : 0);
      p.parent[b * K + j] = (p.compact0 && t == 0) ? b : b * K + kk;
      p.it[b * K + j] = (w >= 0) ? (long long)w : lasts[kk];
    }
    sel_n = j;
    for (; j < K; ++j) {
      p.score_out[b * K + j] = 0.0;
      p.alive_out[b * K + j] = 0;
      p.len_out[b * K + j] = 0;
      p.parent[b * K + j] = (p.compact0 && t == 0) ? b : b * K + j;
      p.it[b * K + j] = p.sos_id;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * T; i += blockDim.x) {
    const int j = i / T, tt = i - j * T;
    int v = 0;
    if (j < sel_n) {
      const int kk = sel_parent[j];
      v = p.tok_in[(b * K + kk) * T + tt];
      if (sel_word[j] >= 0 && tt == p.len_in[b * K + kk]) v = sel_word[j];
    }
    p.tok_out[(b * K + j) * T + tt] = v;
  }
}

// ---- merge of the LogitsSelect records written by the logits GEMM epilogue (common.cuh) --------------------
// One warp reduces the np records of a row to the row max, log(sum exp) and the row's K <= SEL_K best unmasked
// logits (value desc, column asc). Results are broadcast to every lane.
template <int SEL_K>
__device__ __forceinline__ void merge_row(const float* __restrict__ rec, int np, int K, int lane, float& mx_out,
                                          float& lse_out, float (&out_v)[SEL_K], int (&out_i)[SEL_K]) {
  constexpr int SEL_REC = sel_rec(SEL_K);
  float mx = -CUDART_INF_F;
  for (int pi = lane; pi < np; pi += 32) mx = fmaxf(mx, __ldcg(rec + (long long)pi * SEL_REC));
  mx = warp_max(mx);
  float sum = 0.f;
  float cv[SEL_K];
  int ci[SEL_K];
#pragma unroll
  for (int k = 0; k < SEL_K; ++k) {
    cv[k] = -CUDART_INF_F;
    ci[k] = 0x7fffffff;
  }
  for (int pi = lane; pi < np; pi += 32) {
    const float4* r4 = reinterpret_cast<const float4*>(rec + (long long)pi * SEL_REC);
    const float4 h = __ldcg(r4);
    if (h.x > -CUDART_INF_F) sum += h.y * __expf(h.x - mx);
#pragma unroll
    for (int q = 0; q < SEL_K / 4; ++q) {
      const float4 v4 = __ldcg(r4 + 1 + q), i4 = __ldcg(r4 + 1 + SEL_K / 4 + q);
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
      const int ii[4] = {__float_as_int(i4.x), __float_as_int(i4.y), __float_as_int(i4.z), __float_as_int(i4.w)};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (vv[e] > cv[SEL_K - 1]) topk_insert<SEL_K>(cv, ci, vv[e], ii[e]);
    }
  }
  sum = warp_sum(sum);
  mx_out = mx;
  lse_out = logf(sum);
#pragma unroll
  for (int r = 0; r < SEL_K; ++r) {
    out_v[r] = -CUDART_INF_F;
    out_i[r] = 0x7fffffff;
    if (r < K) {
      float bv = cv[0];
      int bi = ci[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      out_v[r] = bv;
      out_i[r] = bi;
      if (ci[0] == bi) {  // the winning lane pops its head
#pragma unroll
        for (int k = 0; k + 1 < SEL_K; ++k) {
          cv[k] = cv[k + 1];
          ci[k] = ci[k + 1];
        }
        cv[SEL_K - 1] = -CUDART_INF_F;
        ci[SEL_K - 1] = 0x7fffffff;
      }
    }
  }
}

// Beam expansion + pooled stable top-K. One CTA per beam ROW scans its logits:
//   pass 1 = thread-local maxima with 128-bit loads; the (K+4)-th largest thread maximum is a threshold
//   tau that at least K unmasked entries reach (<= 4 entries are masked);
//   pass 2 = sum of exp for the log-softmax normaliser + collection of the few entries >= tau;
//   thread 0 ranks them (value desc, index asc) and publishes the row's K candidates.
// The LAST CTA of an image to finish (atomic ticket) pools the candidates of the image's rows with the
// carried finished beams and ranks them stably by fp64 score (captioner.py:409).
constexpr int CAND_CAP = 512;

__global__ void __launch_bounds__(NT) beam_select_kernel(BeamParams p) {
  __shared__ float red[NT / 32];
  __shared__ float cand_v[CAND_CAP];
  __shared__ int cand_i[CAND_CAP];
  __shared__ int cand_n;
  __shared__ int is_last;
  __shared__ float gmax[32];
  __shared__ float stat[2];
  const int m = blockIdx.x, K = p.K, V = p.V, t = p.t;
  const int b = m / K;
  const bool mask_special = (p.pad_id != p.eos_id);
  const int last = (int)p.it[m];
  const bool alive = p.alive_in[m] != 0;
  const bool finished = alive && t > 0 && last == p.eos_id;  // carried unchanged (captioner.py:385-386)

  if (alive && !finished) {
    const float* row = p.logits + (long long)m * p.ld;
    const int v4 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? (V >> 2) : 0;
    // ---- pass 1: thread-local maximum
    float tmax = -CUDART_INF_F;
    for (int i = threadIdx.x; i < v4; i += NT) {
      float4 x = reinterpret_cast<const float4*>(row)[i];
      tmax = fmaxf(fmaxf(tmax, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
    }
    for (int i = v4 * 4 + threadIdx.x; i < V; i += NT) tmax = fmaxf(tmax, row[i]);
    // 32 group maxima (8 lanes each); warp 0 ranks them: row max and the (K+4)-th largest as threshold
    {
      float g = tmax;
      g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, 1));
      g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, 2));
      g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, 4));
      if ((threadIdx.x & 7) == 0) gmax[threadIdx.x >> 3] = g;
      __syncthreads();
      if (threadIdx.x < 32) {
        const float v = gmax[threadIdx.x];
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float o = __shfl_sync(0xffffffffu, v, j);
          rank += (o > v || (o == v && j < (int)threadIdx.x)) ? 1 : 0;
        }
        const int need = (K + 4 < 32) ? K + 4 : 32;
        if (rank == need - 1) stat[1] = v;
        if (rank == 0) stat[0] = v;
      }
      __syncthreads();
    }
    const float mx = stat[0];
    const float tau = stat[1];
    if (threadIdx.x == 0) cand_n = 0;
    __syncthreads();
    // ---- pass 2: normaliser + candidates
    float s = 0.f;
    auto visit = [&](float x, int i) {
      s += __expf(x - mx);
      if (x >= tau) {
        const bool masked = (mask_special && (i == p.pad_id || i == p.sos_id || i == p.unk_id)) || (p.constraint && i == last);
        if (!masked) {
          int slot = atomicAdd(&cand_n, 1);
          if (slot < CAND_CAP) {
            cand_v[slot] = x;
            cand_i[slot] = i;
          }
        }
      }
    };
    for (int i = threadIdx.x; i < v4; i += NT) {
      float4 x = reinterpret_cast<const float4*>(row)[i];
      visit(x.x, 4 * i);
      visit(x.y, 4 * i + 1);
      visit(x.z, 4 * i + 2);
      visit(x.w, 4 * i + 3);
    }
    for (int i = v4 * 4 + threadIdx.x; i < V; i += NT) visit(row[i], i);
    const float ls = logf(block_sum(s, red));
    __syncthreads();
    if (threadIdx.x == 0) {
      const int n = cand_n < CAND_CAP ? cand_n : CAND_CAP;
      int r = 0;
      for (; r < K; ++r) {
        int best = -1;
        for (int i = 0; i < n; ++i) {
          if (cand_i[i] < 0) continue;
          if (best < 0 || cand_v[i] > cand_v[best] || (cand_v[i] == cand_v[best] && cand_i[i] < cand_i[best])) best = i;
        }
        if (best < 0) break;
        p.cand_lp[m * KMAX + r] = (cand_v[best] - mx) - ls;  // log_softmax value, fp32 like the reference
        p.cand_word[m * KMAX + r] = cand_i[best];
        cand_i[best] = -1;
      }
      p.cand_count[m] = r;
    }
  }
  // ---- ticket: the last CTA of this image merges
  if (threadIdx.x == 0) {
    __threadfence();
    const int prev = atomicAdd(p.ticket + b, 1);
    is_last = (prev == K - 1);
    if (is_last) {
      p.ticket[b] = 0;  // re-armed for the next step
      __threadfence();
    }
  }
  __syncthreads();
  if (!is_last) return;

  beam_pool(p, b);
}

// Fused path: the logits GEMM already produced per-slice partials; one CTA per image (4 warps, a warp per row)
// merges them and pools the image's beams.
template <int SEL_K>
__global__ void __launch_bounds__(128) beam_merge_kernel(BeamParams p) {
  constexpr int SEL_REC = sel_rec(SEL_K);
  pdl_trigger();
  pdl_wait();
  // the rows' candidates stay in shared memory (the pooling thread used to read them back from global scratch)
  __shared__ float s_lp[KMAX * KMAX];
  __shared__ int s_word[KMAX * KMAX], s_cnt[KMAX];
  const int b = blockIdx.x, K = p.K, t = p.t;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int kk = warp; kk < K; kk += 4) {
    const int m = b * K + kk;
    const bool alive = p.alive_in[m] != 0;
    const bool finished = alive && t > 0 && (int)p.it[m] == p.eos_id;
    if (lane == 0) s_cnt[kk] = 0;
    if (alive && !finished) {
      float mx, lse, v[SEL_K];
      int w[SEL_K];
      const long long mrec = (p.compact0 && t == 0) ? b : m;  // step 0 computed one row per image
      merge_row<SEL_K>(p.rec + mrec * p.np * SEL_REC, p.np, K, lane, mx, lse, v, w);
      if (lane == 0) {
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < SEL_K; ++r) {
          if (r < K && v[r] > -CUDART_INF_F) {
            s_lp[kk * KMAX + r] = (v[r] - mx) - lse;  // log_softmax value, fp32 like the reference
            s_word[kk * KMAX + r] = w[r];
            cnt = r + 1;
          }
        }
        s_cnt[kk] = cnt;
      }
    }
  }
  __syncthreads();
  beam_pool(p, b, s_lp, s_word, s_cnt);
  if (p.h_state) {  // the next step's operand rows of this image's beams, read through the parents chosen above
    __syncthreads();
    const int c = threadIdx.x * 4;
    for (int j = 0; j < K; ++j) {
      const long long m = (long long)b * K + j;
      const long long src = p.parent[m];
      const float4 ha = *reinterpret_cast<const float4*>(p.h_state + src * H + c);
      const float4 hl = *reinterpret_cast<const float4*>(p.h_state + (p.state_rows + src) * H + c);
      p.x1.store4(m, c, hl);
      p.x1.store4(m, H + c, ha);
      p.x2.store4(m, 2 * H + c, hl);
    }
  }
}

// Fused greedy pick (sample_max = 1): a warp per row takes the argmax and the normaliser from the records.
__global__ void __launch_bounds__(256) greedy_merge_kernel(GreedyParams p) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int t = p.t;
  if (b >= p.B) return;
  if (t > 0 && p.alive_count[t - 1] == 0) return;  // whole-batch early stop (captioner.py:343-344)
  constexpr int SEL_K = 4, SEL_REC = sel_rec(4);
  float mx, lse, v[SEL_K];
  int w[SEL_K];
  merge_row<SEL_K>(p.rec + (long long)b * p.np * SEL_REC, p.np, 1, lane, mx, lse, v, w);
  if (lane == 0) {
    const int unf = p.unfinished[b];
    const long long tok = unf ? w[0] : 0;  // finished rows emit PAD (captioner.py:338)
    p.seq[(long long)b * p.T + t] = tok;
    p.seq_logprobs[(long long)b * p.T + t] = (v[0] - mx) - lse;  // written unmasked (captioner.py:340)
    p.seq_masks[(long long)b * p.T + t] = unf ? 1.f : 0.f;
    const int unf2 = unf && (tok != p.eos_id);
    p.unfinished[b] = unf2;
    p.it[b] = tok;
    if (unf2) atomicAdd(p.alive_count + t, 1);
  }
}

// Scheduled sampling (captioner.py:219-228): with probability ss_prob a row is fed a word drawn from the PREVIOUS
// step's distribution instead of the ground-truth word. The draw is Gumbel-max on the previous log-probs (noise
// injected or counter-based, like the sampled decode). One CTA per row.
__global__ void __launch_bounds__(NT) ss_select_kernel(const float* __restrict__ logp_prev, long long ld_logp,
                                                       const long long* __restrict__ truth, long long ld_truth,
                                                       const float* __restrict__ uniform, float prob,
                                                       const float* __restrict__ noise, unsigned long long seed, int t, int B,
                                                       int V, long long* __restrict__ it) {
  __shared__ float redv[NT / 32];
  __shared__ int redi[NT / 32];
  const int b = blockIdx.x;
  if (!(uniform[b] < prob)) {
    if (threadIdx.x == 0) it[b] = truth[(long long)b * ld_truth];
    return;
  }
  const float* row = logp_prev + (long long)b * ld_logp;
  float sv = -CUDART_INF_F;
  int si = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += NT) {
    const float g = noise ? noise[(long long)b * V + i]
                          : gumbel_counter(seed, ((unsigned long long)t * B + b) * (unsigned long long)V + i);
    const float s = row[i] + g;
    if (s > sv) {
      sv = s;
      si = i;
    }
  }
  block_argmax(sv, si, redv, redi);
  if (threadIdx.x == 0) it[b] = si;
}

__global__ void beam_init_kernel(long long* it, int* alive, int* len, double* score, int* parent, int M, int K, int sos_id) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  it[m] = sos_id;
  alive[m] = (m % K) == 0;  // a single <SOS> candidate per image at t = 0 (captioner.py:379)
  len[m] = 0;
  score[m] = 0.0;
  parent[m] = m;
}
__global__ void beam_finalize_kernel(const int* tok, const int* len, const double* score, long long* tokens_out,
                                     double* scores_out, int* lengths_out, int M, int T) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M * T) tokens_out[i] = tok[i];
  if (i < M) {
    scores_out[i] = score[i];
    lengths_out[i] = len[i];
  }
}
__global__ void greedy_init_kernel(long long* it, int* unfinished, int B, int sos_id) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  it[b] = sos_id;
  unfinished[b] = 1;
}

}  // namespace

int launch_log_softmax(float* x, long long ld, int M, int V, cudaStream_t stream) {
  if (M <= 0) return 0;
  ProfScope ps(ISC_K_SELECT, (double)M * V * 4.0 * 2, stream);
  log_softmax_kernel<<<M, NT, 0, stream>>>(x, ld, V);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_greedy_select(const GreedyParams& p, cudaStream_t stream) {
  if (p.rec) {
    ISC_REQUIRE(p.sample_mode == 0, "fused greedy pick only serves sample_max = 1");
    ProfScope ps(ISC_K_SELECT, (double)p.B * p.np * sel_rec(4) * 4.0, stream);
    ISC_CUDA(launch_pdl(greedy_merge_kernel, dim3((p.B + 7) / 8), dim3(256), 0, stream, p));
    ISC_LAUNCH_CHECK();
    return 0;
  }
  ProfScope ps(ISC_K_SELECT, (double)p.B * p.V * 4.0 * (p.sample_mode == 1 ? 2 : 1), stream);
  greedy_select_kernel<<<p.B, NT, 0, stream>>>(p);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_beam_select(const BeamParams& p, cudaStream_t stream) {
  ISC_REQUIRE(p.K >= 1 && p.K <= KMAX, "beam size %d not in 1..%d", p.K, KMAX);
  ISC_REQUIRE(p.cand_lp && p.cand_word && p.cand_count && p.ticket, "beam_select: scratch buffers missing");
  if (p.rec) {
    ISC_REQUIRE(p.K <= p.k_sel && (p.k_sel == 4 || p.k_sel == 8), "fused beam merge: beam size %d exceeds the %d candidates per slice",
                p.K, p.k_sel);
    ProfScope ps(ISC_K_SELECT, (double)p.B * p.K * p.np * sel_rec(p.k_sel) * 4.0, stream);
    if (p.k_sel == 8) ISC_CUDA(launch_pdl(beam_merge_kernel<8>, dim3(p.B), dim3(128), 0, stream, p));
    else ISC_CUDA(launch_pdl(beam_merge_kernel<4>, dim3(p.B), dim3(128), 0, stream, p));
    ISC_LAUNCH_CHECK();
    return 0;
  }
  ProfScope ps(ISC_K_SELECT, (double)p.B * p.K * p.V * 4.0, stream);
  beam_select_kernel<<<p.B * p.K, NT, 0, stream>>>(p);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_beam_init(long long* it, int* alive, int* len, double* score, int* parent, int B, int K, int sos_id,
                     cudaStream_t stream) {
  int M = B * K;
  ProfScope ps(ISC_K_SELECT, (double)M * 32.0, stream);
  beam_init_kernel<<<(M + 255) / 256, 256, 0, stream>>>(it, alive, len, score, parent, M, K, sos_id);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_beam_finalize(const int* tok, const int* len, const double* score, long long* tokens_out, double* scores_out,
                         int* lengths_out, int B, int K, int T, cudaStream_t stream) {
  int M = B * K;
  ProfScope ps(ISC_K_SELECT, (double)M * T * 12.0, stream);
  beam_finalize_kernel<<<(M * T + 255) / 256, 256, 0, stream>>>(tok, len, score, tokens_out, scores_out, lengths_out, M, T);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_ss_select(const float* logp_prev, long long ld_logp, const long long* truth, long long ld_truth,
                     const float* uniform, float prob, const float* noise, unsigned long long seed, int t, int B, int V,
                     long long* it, cudaStream_t stream) {
  ProfScope ps(ISC_K_SELECT, (double)B * V * 4.0 * (noise ? 2 : 1), stream);
  ss_select_kernel<<<B, NT, 0, stream>>>(logp_prev, ld_logp, truth, ld_truth, uniform, prob, noise, seed, t, B, V, it);
  ISC_LAUNCH_CHECK();
  return 0;
}
int launch_greedy_init(long long* it, int* unfinished, int B, int sos_id, cudaStream_t stream) {
  ProfScope ps(ISC_K_SELECT, (double)B * 12.0, stream);
  greedy_init_kernel<<<(B + 255) / 256, 256, 0, stream>>>(it, unfinished, B, sos_id);
  ISC_LAUNCH_CHECK();
  return 0;
}

}  // namespace isc
