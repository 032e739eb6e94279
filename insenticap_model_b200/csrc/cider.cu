// GPU CIDEr-D reward (self_critical/utils.py:56-83 + cider/pyciderevalcap/ciderD/ciderD_scorer.py).
//
// n-grams (n = 1..4) are EXACT packed keys: 16 bits per token holding token+1 (so id 0 is
// representable and the order n is the number of non-zero fields) — never a lossy hash. The
// document-frequency table is an open-addressing hash map key -> count; "hashing" only picks the
// probe start. All tf-idf arithmetic is fp64 like the reference (numpy scalars / python floats).
//   build : one CTA per image, its references' n-gram SET is inserted once (compute_doc_freq :52-64)
//   score : one CTA per hypothesis; thread g owns n-gram position g of the hypothesis / reference
//           (counts2vec :121-145, sim :147-173, length = number of bigram positions :142-143).
#include "kernels.cuh"

namespace isc {
namespace {

constexpr int MAXW = 32;            // words per caption after _array_to_str (incl. the appended EOS)
constexpr int MAXG = 4 * MAXW;      // n-gram positions per caption (upper bound)
constexpr int MAXWB = 64;           // words per reference caption in the DF corpus (untruncated captions)
constexpr int MAXSET = 2048;        // n-gram positions over all references of one image (DF build)
constexpr int NTH = 128;

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// position g -> (n, start) for a caption of W words: n-grams of order n occupy W-n+1 slots
__device__ __forceinline__ bool gram_at(int g, int W, int& n, int& start) {
  for (n = 1; n <= 4; ++n) {
    int cnt = W - n + 1;
    if (cnt <= 0) return false;
    if (g < cnt) {
      start = g;
      return true;
    }
    g -= cnt;
  }
  return false;
}
__device__ __forceinline__ int num_grams(int W) {
  int t = 0;
  for (int n = 1; n <= 4; ++n) t += (W - n + 1 > 0) ? (W - n + 1) : 0;
  return t;
}
__device__ __forceinline__ unsigned long long pack_key(const int* words, int start, int n) {
  unsigned long long k = 0;
  for (int i = 0; i < n; ++i) k |= (unsigned long long)(words[start + i] + 1) << (16 * i);
  return k;
}

__device__ __forceinline__ double df_lookup(const unsigned long long* keys, const int* counts, long long slots,
                                            unsigned long long key) {
  long long h = (long long)(mix64(key) & (unsigned long long)(slots - 1));
  for (long long probe = 0; probe < slots; ++probe) {
    unsigned long long k = keys[h];
    if (k == key) return (double)counts[h];
    if (k == 0ULL) return 0.0;
    h = (h + 1) & (slots - 1);
  }
  return 0.0;
}

__global__ void __launch_bounds__(NTH) cider_build_kernel(const int* __restrict__ ref_tokens, const int* __restrict__ ref_lens,
                                                          int ref_ld, const int* __restrict__ img_offsets,
                                                          unsigned long long* keys, int* counts, long long slots,
                                                          int* overflow) {
  __shared__ unsigned long long set[MAXSET];
  __shared__ int words[MAXWB];
  __shared__ int n_set;
  const int img = blockIdx.x;
  if (threadIdx.x == 0) n_set = 0;
  __syncthreads();
  for (int r = img_offsets[img]; r < img_offsets[img + 1]; ++r) {
    int W = ref_lens[r];
    if (W > MAXWB) {
      if (threadIdx.x == 0) atomicExch(overflow, 1);
      W = MAXWB;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += NTH) words[i] = ref_tokens[(long long)r * ref_ld + i];
    __syncthreads();
    const int G = num_grams(W);
    const int base = n_set;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += NTH) {
      int n, st;
      if (gram_at(g, W, n, st) && base + g < MAXSET) set[base + g] = pack_key(words, st, n);
    }
    if (threadIdx.x == 0) {
      if (base + G > MAXSET) {
        atomicExch(overflow, 1);
        n_set = MAXSET;
      } else {
        n_set = base + G;
      }
    }
    __syncthreads();
  }
  const int N = n_set;
  for (int i = threadIdx.x; i < N; i += NTH) {
    const unsigned long long key = set[i];
    bool first = true;
    for (int j = 0; j < i; ++j)
      if (set[j] == key) {
        first = false;
        break;
      }
    if (!first) continue;
    long long h = (long long)(mix64(key) & (unsigned long long)(slots - 1));
    bool done = false;
    for (long long probe = 0; probe < slots && !done; ++probe) {
      unsigned long long prev = atomicCAS(keys + h, 0ULL, key);
      if (prev == 0ULL || prev == key) {
        atomicAdd(counts + h, 1);
        done = true;
      } else {
        h = (h + 1) & (slots - 1);
      }
    }
    if (!done) atomicExch(overflow, 1);
  }
}

// Vectorise one caption held in `words` (W words): per position g the packed key, whether it is the
// first occurrence, its term frequency and tf-idf weight; norm2[n] accumulates the squared norms.
__device__ __forceinline__ void vectorize(const int* words, int W, unsigned long long* keys, double* vec, int* order,
                                          double* norm2, const unsigned long long* tkeys, const int* tcounts,
                                          long long slots, double log_n_docs, int* tf_out) {
  const int G = num_grams(W);
  const int g = threadIdx.x;
  unsigned long long key = 0ULL;
  int n = 0, st = 0;
  if (g < G && gram_at(g, W, n, st)) key = pack_key(words, st, n);
  if (g < MAXG) {
    keys[g] = key;
    vec[g] = 0.0;
    order[g] = 0;
  }
  if (g < 4) norm2[g] = 0.0;
  __syncthreads();
  if (g < G) {
    int tf = 0;
    bool first = true;
    for (int j = 0; j < G; ++j) {
      if (keys[j] == key) {
        ++tf;
        if (j < g) first = false;
      }
    }
    if (first) {
      const double df = log(fmax(1.0, df_lookup(tkeys, tcounts, slots, key)));
      const double v = (double)tf * (log_n_docs - df);
      vec[g] = v;
      order[g] = n;  // 1..4 marks a unique entry
      atomicAdd(&norm2[n - 1], v * v);
      if (tf_out) tf_out[g] = tf;
    } else if (tf_out) {
      tf_out[g] = 0;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int hyp_to_words(const long long* hyp, int T, int sos, int eos, int* words) {
  // utils._array_to_str: drop a leading SOS, cut at the first EOS, always append EOS
  int W = 0;
  int i = (T > 0 && hyp[0] == sos) ? 1 : 0;
  for (; i < T && W < MAXW - 1; ++i) {
    long long w = hyp[i];
    if (w == eos) break;
    words[W++] = (int)w;
  }
  words[W++] = eos;
  return W;
}

__global__ void __launch_bounds__(NTH) cider_score_kernel(const unsigned long long* __restrict__ tkeys,
                                                          const int* __restrict__ tcounts, long long slots, double log_n_docs,
                                                          const long long* __restrict__ hyps, int T, const int* __restrict__ hyp_img,
                                                          const int* __restrict__ ref_tokens, const int* __restrict__ ref_lens,
                                                          int ref_ld, const int* __restrict__ img_offsets, int sos, int eos,
                                                          double* __restrict__ scores) {
  __shared__ int words[MAXW];
  __shared__ int Wsh;
  __shared__ unsigned long long hk[MAXG], rk[MAXG];
  __shared__ double hv[MAXG], rv[MAXG];
  __shared__ int ho[MAXG], ro[MAXG];
  __shared__ double hn2[4], rn2[4], val[4], total[4];
  const int h = blockIdx.x;
  if (threadIdx.x == 0) Wsh = hyp_to_words(hyps + (long long)h * T, T, sos, eos, words);
  if (threadIdx.x < 4) total[threadIdx.x] = 0.0;
  __syncthreads();
  const int Wh = Wsh;
  vectorize(words, Wh, hk, hv, ho, hn2, tkeys, tcounts, slots, log_n_docs, nullptr);
  const int Gh = num_grams(Wh);
  const int len_h = Wh - 1 > 0 ? Wh - 1 : 0;  // "length" = number of bigrams (ciderD_scorer.py:142-143)
  const int img = hyp_img[h];
  const int r0 = img_offsets[img], r1 = img_offsets[img + 1];
  for (int r = r0; r < r1; ++r) {
    __syncthreads();
    int Wr = ref_lens[r];
    Wr = Wr > MAXW ? MAXW : Wr;
    for (int i = threadIdx.x; i < Wr; i += NTH) words[i] = ref_tokens[(long long)r * ref_ld + i];
    if (threadIdx.x < 4) val[threadIdx.x] = 0.0;
    __syncthreads();
    vectorize(words, Wr, rk, rv, ro, rn2, tkeys, tcounts, slots, log_n_docs, nullptr);
    const int Gr = num_grams(Wr);
    const int g = threadIdx.x;
    if (g < Gh && ho[g] > 0) {
      const unsigned long long key = hk[g];
      double refv = 0.0;
      for (int j = 0; j < Gr; ++j)
        if (ro[j] > 0 && rk[j] == key) {
          refv = rv[j];
          break;
        }
      atomicAdd(&val[ho[g] - 1], fmin(hv[g], refv) * refv);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      const int n = threadIdx.x;
      double v = val[n];
      const double nh = sqrt(hn2[n]), nr = sqrt(rn2[n]);
      if (nh != 0.0 && nr != 0.0) v /= (nh * nr);
      const int len_r = Wr - 1 > 0 ? Wr - 1 : 0;
      const double delta = (double)(len_h - len_r);
      v *= exp(-(delta * delta) / (2.0 * 6.0 * 6.0));
      total[n] += v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nref = r1 - r0;
    double s = (total[0] + total[1] + total[2] + total[3]) / 4.0;
    scores[h] = nref > 0 ? s / (double)nref * 10.0 : 0.0;
  }
}

__global__ void __launch_bounds__(NTH) cider_counts_kernel(const long long* hyp, int T, int sos, int eos,
                                                           unsigned long long* keys_out, int* counts_out, int* n_out) {
  __shared__ int words[MAXW];
  __shared__ int Wsh;
  __shared__ unsigned long long hk[MAXG];
  __shared__ double hv[MAXG];
  __shared__ int ho[MAXG], tf[MAXG];
  __shared__ double hn2[4];
  if (threadIdx.x == 0) Wsh = hyp_to_words(hyp, T, sos, eos, words);
  __syncthreads();
  // an empty table (slots = 1, key 0) is enough: only the term frequencies are wanted
  __shared__ unsigned long long zero_key;
  __shared__ int zero_cnt;
  if (threadIdx.x == 0) {
    zero_key = 0ULL;
    zero_cnt = 0;
  }
  __syncthreads();
  vectorize(words, Wsh, hk, hv, ho, hn2, &zero_key, &zero_cnt, 1, 0.0, tf);
  if (threadIdx.x == 0) {
    int n = 0;
    const int G = num_grams(Wsh);
    for (int g = 0; g < G; ++g)
      if (ho[g] > 0) {
        keys_out[n] = hk[g];
        counts_out[n] = tf[g];
        ++n;
      }
    *n_out = n;
  }
}

__global__ void reward_kernel(const double* scores, int B, int T, double* rewards) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * T) {
    int b = i / T;
    rewards[i] = scores[b] - scores[B + b];
  }
}

}  // namespace
}  // namespace isc

using namespace isc;

extern "C" {

size_t isc_cider_table_bytes(int64_t table_slots) {
  if (table_slots <= 0 || (table_slots & (table_slots - 1)) != 0) return 0;
  return (size_t)table_slots * (sizeof(unsigned long long) + sizeof(int));
}

int isc_cider_build_df(const int32_t* ref_tokens, const int32_t* ref_lens, int32_t ref_ld, const int32_t* img_offsets,
                       int32_t n_images, void* table, int64_t table_slots, int32_t* overflow_flag, isc_stream_t stream) {
  ISC_REQUIRE(ref_tokens && ref_lens && img_offsets && table && overflow_flag, "NULL pointer");
  ISC_REQUIRE(table_slots > 0 && (table_slots & (table_slots - 1)) == 0, "table_slots must be a power of two");
  ISC_REQUIRE(n_images > 0 && ref_ld > 0, "n_images / ref_ld must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ISC_CUDA(cudaMemsetAsync(table, 0, isc_cider_table_bytes(table_slots), s));
  ISC_CUDA(cudaMemsetAsync(overflow_flag, 0, sizeof(int), s));
  unsigned long long* keys = static_cast<unsigned long long*>(table);
  int* counts = reinterpret_cast<int*>(keys + table_slots);
  ProfScope ps(ISC_K_CIDER, (double)n_images * 5 * 62 * 32.0, s);
  cider_build_kernel<<<n_images, NTH, 0, s>>>(ref_tokens, ref_lens, ref_ld, img_offsets, keys, counts, table_slots,
                                              overflow_flag);
  ISC_LAUNCH_CHECK();
  return 0;
}

int isc_cider_score(const void* table, int64_t table_slots, double log_n_docs, const int64_t* hyps, int32_t T,
                    const int32_t* hyp_img, int32_t n_hyp, const int32_t* ref_tokens, const int32_t* ref_lens,
                    int32_t ref_ld, const int32_t* img_offsets, int32_t sos_id, int32_t eos_id, double* scores,
                    isc_stream_t stream) {
  ISC_REQUIRE(table && hyps && hyp_img && ref_tokens && ref_lens && img_offsets && scores, "NULL pointer");
  ISC_REQUIRE(table_slots > 0 && (table_slots & (table_slots - 1)) == 0, "table_slots must be a power of two");
  ISC_REQUIRE(T > 0 && T < MAXW, "T=%d must be in 1..%d", T, MAXW - 1);
  if (n_hyp <= 0) return 0;
  const unsigned long long* keys = static_cast<const unsigned long long*>(table);
  const int* counts = reinterpret_cast<const int*>(keys + table_slots);
  // ~372 n-gram probes x 32 B sectors per scored hypothesis incl. its references
  ProfScope ps(ISC_K_CIDER, (double)n_hyp * 372 * 32.0, static_cast<cudaStream_t>(stream));
  cider_score_kernel<<<n_hyp, NTH, 0, static_cast<cudaStream_t>(stream)>>>(
      keys, counts, table_slots, log_n_docs, reinterpret_cast<const long long*>(hyps), T, hyp_img, ref_tokens, ref_lens,
      ref_ld, img_offsets, sos_id, eos_id, scores);
  ISC_LAUNCH_CHECK();
  return 0;
}

int isc_self_critical_reward(const double* scores, int32_t B, int32_t T, double* rewards, isc_stream_t stream) {
  ISC_REQUIRE(scores && rewards && B > 0 && T > 0, "bad reward arguments");
  ProfScope ps(ISC_K_CIDER, (double)B * T * 8.0, static_cast<cudaStream_t>(stream));
  reward_kernel<<<(B * T + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, B, T, rewards);
  ISC_LAUNCH_CHECK();
  return 0;
}

int isc_cider_ngram_counts(const int64_t* hyp, int32_t T, int32_t sos_id, int32_t eos_id, uint64_t* keys,
                           int32_t* counts, int32_t* n_out, isc_stream_t stream) {
  ISC_REQUIRE(hyp && keys && counts && n_out && T > 0 && T < MAXW, "bad ngram_counts arguments");
  ProfScope ps(ISC_K_CIDER, 1024.0, static_cast<cudaStream_t>(stream));
  cider_counts_kernel<<<1, NTH, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(hyp), T, sos_id, eos_id, reinterpret_cast<unsigned long long*>(keys), counts,
      n_out);
  ISC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
