// tcgen05 (5th-gen tensor core) GEMM for sm_100a:  C = act(A · W^T + bias + rowadd + addmat)
//
//   A [M,K] and W [N,K] are K-major bf16 "planes". With passes == 3 every operand has a hi and a
//   lo plane (x ~= hi + lo) and the kernel accumulates hi·hi + lo·hi + hi·lo in one fp32 TMEM
//   accumulator ("bf16x3": 16 mantissa bits per operand on the bf16 tensor pipe, SURVEY.md
//   section 0); with passes == 1 only the hi planes are read.
//
//   Persistent: one CTA per SM walks 128x128 output tiles (n fastest, so concurrent CTAs share an
//   A panel in L2). Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread
//   tcgen05.mma issuer, warps 2..5 = epilogue. Operand tiles are 128 rows x 64 bf16 (128 B rows) in
//   the 128B-swizzled K-major layout that both TMA (CU_TENSOR_MAP_SWIZZLE_128B) and the UMMA
//   shared-memory descriptor (layout type 2, SBO = 1024 B) understand. Three mbarrier pipelines:
//   smem stages full/empty (TMA <-> MMA), and two TMEM accumulator stages full/empty
//   (MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1. The 8 epilogue warps
//   move 32x32 sub-tiles TMEM -> registers -> XOR-swizzled smem -> registers so that every global access
//   (bias / additive terms in, fp32 and bf16 planes out) is a coalesced 128-bit access.
//
// This is the only dense-contraction kernel of the decode path in the tensor-core precisions:
// LSTM gate GEMMs, attention projections, vocabulary logits and the prologue feature embeddings
// (reference: nn.LSTMCell / nn.Linear calls in /root/reference/models/captioner.py:138-161).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

// LSTM cell in the fused gate-GEMM epilogue: 1 = raw-MUFU sigmoid / tanh (default), 0 = libdevice expf / tanhf / division
#ifndef ISC_CELL_FAST
#define ISC_CELL_FAST 1
#endif

namespace isc {
namespace tc {

constexpr int BM = 128;
constexpr int EPI_WARPS = 8;                // two per TMEM lane quarter, each owns half of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int ACC_STAGES = 2;               // TMEM accumulator double buffer
constexpr int CONV_WARPS = 4;               // AF kernels: warps that split the fp32 A tile into bf16 planes (4: a row per thread; 8: half a row,
                                            // measured no faster and the 576-thread kernel spills)
constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;  // one XOR-swizzled 32x32 fp32 staging tile per epilogue warp

// BN = 128 for the step GEMMs (M = 3072 rows: more, smaller tiles fill 148 SMs better), BN = 256 where there are many
// tiles (vocabulary logits, prologue): 25 % less L2 -> smem traffic per flop, which is what bounds this kernel.
// CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x BN tile. Each CTA loads its 128 rows of A and
// HALF of the B rows, and the pair's MMAs read both halves — per flop, half the B bytes cross L2 -> smem and the
// smem port; the accumulator of each CTA (its 128 rows x BN columns) stays in its own TMEM.
// AF = 1: the A operand arrives as fp32 (TMA into a staging ring) and four extra warps split it into the bf16 hi/lo
// operand tiles in shared memory — no separate plane-split pass over HBM (used for the raw region features).
template <int PASSES, int BN, int CG = 1, int AF = 0>
struct Cfg {
  static constexpr int kPlanes = PASSES == 3 ? 2 : 1;
  // K extent of one pipeline stage. 64 bf16 = one 128-byte swizzle row; the wide split-bf16 tile uses 32 (64-byte
  // swizzle) so that four 48 KiB stages fit instead of two 96 KiB ones — two stages cannot keep enough bytes in
  // flight to cover the L2 latency.
  static constexpr int kBK = (AF || (PASSES == 3 && BN == 256 && CG == 1)) ? 32 : 64;
  static constexpr int kBRows = BN / CG;  // B rows this CTA loads
  static constexpr int kATileBytes = BM * kBK * 2;
  static constexpr int kBTileBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kPlanes * (kATileBytes + kBTileBytes);  // A_hi,(A_lo),B_hi,(B_lo)
  static constexpr int kStgBytes = AF ? BM * kBK * 4 : 0;  // one fp32 A tile (128 rows x 128 B)
  static constexpr int kStgSlots = 2;
  static constexpr int kStages = (192 * 1024 - kStgSlots * kStgBytes) / kStageBytes;  // x3: 3 / 4 stages, bf16: 6 / 4
  static constexpr int kTmemCols = ACC_STAGES * BN;  // 256 / 512 columns (power of two)
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStgSlots * kStgBytes + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kThreads = NUM_THREADS + (AF ? CONV_WARPS * 32 : 0);
};

enum { EPI_STD = 0, EPI_LOGITS = 1, EPI_LSTM = 2, EPI_LOGITS8 = 3, EPI_GATE = 4 };  // EPI_LOGITS8: 8 candidates per slice

struct EpiParams {
  const float* bias;
  const float* rowadd;
  long long ld_rowadd;
  int rows_per_group;
  const float* addmat;
  long long ld_addmat;
  int act;
  float* c;
  long long ldc;
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  long long ldp;
  Half16Out h16;  // optional fp16 copy of the result (EPI_STD)
  int M, N, K;
  // EPI_LOGITS: per (row, 128-column slice) online-softmax partials + top candidates instead of the logits
  LogitsSelect sel;
  // EPI_LSTM: the LSTM cell applied to the four gate pre-activations of each hidden unit
  LstmEpilogue lstm;
  // EPI_GATE: gate weight and mixed context of Attention.forward from the four column tiles of a row block (cluster of 4)
  GateEpilogue gate;
  int lstm_wide;  // EPI_LSTM with BN = 256: 256-column tiles per 128-row block, the rest of the row in 128-column tiles
  // 3x3 convolution as ONE GEMM over a zero-bordered 16x16 grid per image (rows = image * 256 + y * 16 + x): the K loop
  // runs over 9 segments of seg_kb k-blocks; segment s reads the A rows shifted by seg_off[s] = dy * 16 + dx (TMA
  // zero-fills rows outside the tensor), W is [N][9 * C] with K index s * C + c. seg_kb == 0: plain GEMM.
  int seg_kb;
  int seg_off[9];
  int zero_border;  // force the output rows on the 16x16 grid's border to zero (they feed the next convolution as padding)
  // split-K (template SK): the kernel is launched in clusters of split_k CTAs; CTA r of a cluster runs slice r of its tile's
  // K loop and the cluster sums the partial tiles through distributed shared memory (see the SK epilogue)
  int split_k;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trap (an error code on the host), never a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("isc gemm_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 cache-policy operands (the encodings createpolicy.fractional.L2::evict_* 1.0 produces). The weight operand of the
// step GEMMs is re-read sixteen times per call but the attention kernel streams 0.8 GB through L2 between two uses:
// loaded evict_last, the ~50 MB of weight planes a step touches survive that stream and the GEMM pipelines turn around
// in L2 latency instead of HBM latency.
constexpr unsigned long long kEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                      unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// cta_group::2 variants: the load signals the LEADER CTA's mbarrier (peer bit of the shared::cluster address cleared),
// the MMA spans both CTAs of the pair, the commit arrives on the same barrier offset in both CTAs.
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {  // arrive on `bar` of CTA `cta` of the cluster
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes from the shared memory of CTA `cta` of the cluster, at the address `local` has in this CTA
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local, uint32_t cta) {
  float4 v;
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %4, %5;\n\t"
      "ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [ra];\n\t}"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "r"(local), "r"(cta)
      : "memory");
  return v;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local, uint32_t cta) {
  float v;
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %1, %2;\n\t"
      "ld.shared::cluster.f32 %0, [ra];\n\t}"
      : "=f"(v)
      : "r"(local), "r"(cta)
      : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major tile, 128-byte swizzle, rows of 128 B, 8-row groups
// 1024 B apart (cute::UMMA::SmemDescriptor: start [0,14), LBO [16,30), SBO [32,46),
// version [46,48) = 1, layout_type [61,64) = 2 for SWIZZLE_128B).
// ROW_BYTES = 128: SWIZZLE_128B (layout type 2), 64: SWIZZLE_64B (layout type 4); SBO = 8 rows.
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                     // LBO: unused for swizzled K-major
  d |= static_cast<uint64_t>((8 * ROW_BYTES) >> 4) << 32;  // SBO
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(ROW_BYTES == 128 ? 2 : 4) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both
// K-major, N at [17,23) in units of 8, M at [24,29) in units of 16.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ kernel
__device__ __forceinline__ float4 load4_guarded(const float* p, int n, int N) {
  // p points at column n of a row; columns >= N do not exist
  if (n + 3 < N && (reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) r.x = __ldg(p);
  if (n + 1 < N) r.y = __ldg(p + 1);
  if (n + 2 < N) r.z = __ldg(p + 2);
  if (n + 3 < N) r.w = __ldg(p + 3);
  return r;
}

__device__ __forceinline__ unsigned pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (unsigned)__bfloat16_as_ushort(a) | ((unsigned)__bfloat16_as_ushort(b) << 16);
}

template <int ACT>
__device__ __forceinline__ float act_ct(float v) {
  if (ACT == ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == ACT_TANH) return tanh_ex2(v);  // two raw MUFU ops, |err| ~ 2e-7 (the gate GEMM, whose output only feeds a dot product + sigmoid)
  if (ACT == ACT_EXPNEG2_RELU) return exp_neg2(fmaxf(v, 0.0f));
  if (ACT == ACT_SIGMOID) return sigmoid_accurate(v);
  return v;
}

// EPI_LSTM tile list. A tile holds all four gates of its hidden units (columns = [i f g o] x 16 units per 64-column
// group). BN = 128: uniform 32-unit tiles. BN = 256: MIXED widths — per 128-row block `wide` tiles of 64 units (256
// columns) first, then the remaining units in 32-unit (128-column) tiles, enumerated wide tiles first: with
// M = 3072 (24 row blocks) and wide = 6 that is 144 wide tiles + 96 narrow ones, i.e. one wide and (for 92 CTAs) one
// narrow tile per SM instead of 1.3 waves of wide or 2.6 waves of narrow tiles.
template <int BN, int ROWS /* rows of a tile: BM, or 2 * BM for a CTA pair */>
__device__ __forceinline__ void lstm_tile(int tile, int tiles_m, int tiles_n, int wide, int& m0, int& u0, int& width) {
  if (BN == 128) {
    m0 = (tile / tiles_n) * ROWS;
    u0 = (tile % tiles_n) * 32;
    width = 128;
    return;
  }
  const int n_wide = tiles_m * wide;
  if (tile < n_wide) {
    m0 = (tile / wide) * ROWS;
    u0 = (tile % wide) * 64;
    width = 256;
  } else {
    const int nn = (tiles_n - wide) * 2, t = tile - n_wide;
    m0 = (t / nn) * ROWS;
    u0 = wide * 64 + (t % nn) * 32;
    width = 128;
  }
}

#ifdef ISC_GEMM_TRACE
// Debug build only (profiles/gemm_trace.py): per-CTA globaltimer stamps of the fused-LSTM GEMM's phases.
__device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ void trace_stamp(int slot) {
  if (g_trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[blockIdx.x * 16 + slot] = t;
  }
}
// which instantiation stamps: the fused-LSTM GEMM by default, -DISC_TRACE_AF the fp32-A prologue GEMM, or any
// expression over the template parameters, e.g. -DISC_TRACE_SEL="(EPI==EPI_STD&&BN==256&&ACT==ACT_NONE&&AF==0&&CG==1)"
#ifndef ISC_TRACE_SEL
#ifdef ISC_TRACE_AF
#define ISC_TRACE_SEL (AF == 1)
#else
#define ISC_TRACE_SEL (EPI == EPI_LSTM)
#endif
#endif
#define ISC_TRACE(cond, slot) do { if (ISC_TRACE_SEL && (cond)) trace_stamp(slot); } while (0)
#else
#define ISC_TRACE(cond, slot) do { } while (0)
#endif

// H16: the standard epilogue also writes the fp16 copy described by EpiParams::h16 (a compile-time switch: as a run-time
// test inside the store loop it cost every plain GEMM of the decode step ~4 us)
// SK: split-K across the CTAs of a thread-block cluster, for GEMMs with fewer tiles than SMs and a long K loop
// (EpiParams::split_k = cluster size; EPI_STD, ACT_NONE, CG = 1, BN = 128 only).
template <int PASSES, int BN, int ACT, int EPI, int CG, int AF, int H16 = 0, int SK = 0>
__global__ void __launch_bounds__(NUM_THREADS + (AF ? CONV_WARPS * 32 : 0), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const EpiParams ep) {
  using C = Cfg<PASSES, BN, CG, AF>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + C::kStages * C::kStageBytes;  // AF: fp32 A staging ring (map_a_hi is the fp32 map then)
  float* epi_smem = reinterpret_cast<float*>(stg + C::kStgSlots * C::kStgBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(epi_smem) + EPI_BYTES);
  uint64_t* full_bar = bars;                           // [kStages] TMA (+ converters) -> MMA
  uint64_t* empty_bar = bars + C::kStages;             // [kStages] MMA -> TMA (+ converters)
  uint64_t* acc_full = bars + 2 * C::kStages;          // [ACC_STAGES] MMA -> epilogue
  uint64_t* acc_empty = acc_full + ACC_STAGES;         // [ACC_STAGES] epilogue -> MMA
  uint64_t* stg_full = acc_empty + ACC_STAGES;         // [kStgSlots] TMA -> converters
  uint64_t* stg_empty = stg_full + C::kStgSlots;       // [kStgSlots] converters -> TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_empty + C::kStgSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  ISC_TRACE(threadIdx.x == 0, 0);
  constexpr int BK = C::kBK;
  const int num_kb = ep.seg_kb > 0 ? 9 * ep.seg_kb : (ep.K + BK - 1) / BK;
  const int tiles_n = (ep.N + BN - 1) / BN;
  const int tiles_m = (ep.M + CG * BM - 1) / (CG * BM);  // CG = 2: a tile is 256 rows, 128 per CTA of the pair
  const int num_tiles = (EPI == EPI_LSTM && BN == 256) ? tiles_m * (ep.lstm_wide + (tiles_n - ep.lstm_wide) * 2) : tiles_m * tiles_n;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  // SK: one cluster per tile (grid = tiles * split_k, no walking); this CTA's K slice is its rank in the cluster
  const int split_k = SK ? ep.split_k : 1;
  const int ks = SK ? (int)cluster_ctarank() : 0;
  const int walker = CG == 2 ? (int)(blockIdx.x >> 1) : (SK ? (int)blockIdx.x / split_k : (int)blockIdx.x);  // the pair walks the tile list together
  const int walkers = CG == 2 ? (int)(gridDim.x >> 1) : (SK ? num_tiles : (int)gridDim.x);
  const int kb_begin = SK ? (int)((long long)ks * num_kb / split_k) : 0;
  const int kb_end = SK ? (int)((long long)(ks + 1) * num_kb / split_k) : num_kb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_hi)) : "memory");
    if (PASSES == 3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_lo)) : "memory");
    }
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], AF ? 1 + CG * CONV_WARPS : 1);  // AF: the producer's arrive (B bytes) + one arrive per converter warp (A tiles; of both CTAs of a pair)
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < C::kStgSlots; ++s) {
      mbar_init(&stg_full[s], 1);
      mbar_init(&stg_empty[s], CONV_WARPS);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], CG * EPI_WARPS);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is warp-wide; the same warp frees it at the end
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything is signalled into them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only this CTA's shared memory, TMEM and kernel parameters: it may run while the previous
  // kernel of the stream is still draining (programmatic dependent launch); global memory is read from here on
  pdl_trigger();
  pdl_wait();
  ISC_TRACE(threadIdx.x == 0, 1);

  if (warp == 0) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      int it = 0;
      const unsigned long long wpol = kEvictLast;  // weight (B) operand: keep in L2 across the decode steps
      for (int tile = walker; tile < num_tiles; tile += walkers) {
        int m0 = (tile / tiles_n) * (CG * BM) + cta_rank * BM;
        int n0 = (tile % tiles_n) * BN + cta_rank * C::kBRows;  // CG = 2: this CTA's half of the B rows
        int lw = BN;                                            // EPI_LSTM: n0 = first hidden unit, lw = tile width
        if (EPI == EPI_LSTM) {
          lstm_tile<BN, CG * BM>(tile, tiles_m, tiles_n, ep.lstm_wide, m0, n0, lw);
          m0 += cta_rank * BM;
        }
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          ISC_TRACE(it == 8, 12);
          if (!AF) mbar_wait(&empty_bar[s], ph ^ 1);  // first round passes immediately (AF: see below)
          ISC_TRACE(it == 8, 13);
          uint8_t* st = smem + s * C::kStageBytes;
          uint8_t* stb = st + C::kPlanes * C::kATileBytes;
          const int k0 = kb * BK;
          // A coordinates: plain GEMM (k0, m0); convolution segments shift the rows and wrap the column
          int ak = k0, am = m0;
          if (ep.seg_kb > 0) {
            const int sgm = kb / ep.seg_kb;
            ak = (kb - sgm * ep.seg_kb) * BK;
            am = m0 + ep.seg_off[sgm];
          }
          if (AF) {
            // The fp32 A tile goes into the staging ring for the converter warps FIRST: it only needs a free staging
            // slot, so its load latency (~1.1 us) runs while the operand stage is still being read by the MMAs. (Issued
            // after the wait for the operand stage, the stage's round trip was free -> issue -> A lands -> converted =
            // 2.1 us, i.e. 0.7 us per k-block with three stages against 0.39 us of MMAs: phase trace, DESIGN.md.)
            const int slot = it & 1;
            mbar_wait(&stg_empty[slot], ((it >> 1) & 1) ^ 1);
            mbar_expect_tx(&stg_full[slot], C::kStgBytes);
            tma_load_2d(stg + slot * C::kStgBytes, &map_a_hi, &stg_full[slot], ak, am);
            // B planes straight into the stage
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (CG == 2) {
              // CTA pair: each CTA loads HALF of the tile's B rows; every byte of both halves completes on the leader's barrier
              if (cta_rank == 0) mbar_expect_tx(&full_bar[s], 2 * C::kPlanes * C::kBTileBytes);
              tma_load_2d_pair_hint(stb, &map_b_hi, &full_bar[s], k0, n0, wpol);
              if (PASSES == 3) tma_load_2d_pair_hint(stb + C::kBTileBytes, &map_b_lo, &full_bar[s], k0, n0, wpol);
            } else {
              mbar_expect_tx(&full_bar[s], C::kPlanes * C::kBTileBytes);
              tma_load_2d_hint(stb, &map_b_hi, &full_bar[s], k0, n0, wpol);
              if (PASSES == 3) tma_load_2d_hint(stb + C::kBTileBytes, &map_b_lo, &full_bar[s], k0, n0, wpol);
            }
          } else if (CG == 2 && EPI == EPI_LSTM) {
            // CTA pair: each CTA loads its 128 rows of A and HALF of the tile's (gate-interleaved, contiguous) B rows,
            // 64-row boxes; every byte of both CTAs completes on the leader's barrier
            const int half_rows = lw >> 1;
            if (cta_rank == 0) mbar_expect_tx(&full_bar[s], 2 * C::kPlanes * (C::kATileBytes + half_rows * BK * 2));
            tma_load_2d_pair(st, &map_a_hi, &full_bar[s], k0, m0);
            if (PASSES == 3) tma_load_2d_pair(st + C::kATileBytes, &map_a_lo, &full_bar[s], k0, m0);
#pragma unroll
            for (int q = 0; q < BN / 128; ++q) {
              if (q * 64 < half_rows) {
                const int brow = 4 * n0 + cta_rank * half_rows + q * 64;
                tma_load_2d_pair_hint(stb + q * 64 * BK * 2, &map_b_hi, &full_bar[s], k0, brow, wpol);
                if (PASSES == 3) tma_load_2d_pair_hint(stb + C::kBTileBytes + q * 64 * BK * 2, &map_b_lo, &full_bar[s], k0, brow, wpol);
              }
            }
          } else if (CG == 2) {
            // both CTAs' bytes complete on the leader's barrier, which alone is armed and waited on
            if (cta_rank == 0) mbar_expect_tx(&full_bar[s], 2 * C::kStageBytes);
            tma_load_2d_pair(st, &map_a_hi, &full_bar[s], k0, m0);
            if (PASSES == 3) tma_load_2d_pair(st + C::kATileBytes, &map_a_lo, &full_bar[s], k0, m0);
            tma_load_2d_pair_hint(stb, &map_b_hi, &full_bar[s], k0, n0, wpol);
            if (PASSES == 3) tma_load_2d_pair_hint(stb + C::kBTileBytes, &map_b_lo, &full_bar[s], k0, n0, wpol);
          } else if (EPI == EPI_LSTM) {
            // the weight planes are gate-interleaved: the tile's B rows start at 4 * (first unit), one 128-row box per
            // 128 columns
            mbar_expect_tx(&full_bar[s], C::kPlanes * (C::kATileBytes + lw * BK * 2));
            tma_load_2d(st, &map_a_hi, &full_bar[s], k0, m0);
            if (PASSES == 3) tma_load_2d(st + C::kATileBytes, &map_a_lo, &full_bar[s], k0, m0);
#pragma unroll
            for (int q = 0; q < BN / 128; ++q) {
              if (q * 128 < lw) {
                tma_load_2d_hint(stb + q * 128 * BK * 2, &map_b_hi, &full_bar[s], k0, 4 * n0 + q * 128, wpol);
                if (PASSES == 3) tma_load_2d_hint(stb + C::kBTileBytes + q * 128 * BK * 2, &map_b_lo, &full_bar[s], k0, 4 * n0 + q * 128, wpol);
              }
            }
          } else {
            mbar_expect_tx(&full_bar[s], C::kStageBytes);
            tma_load_2d(st, &map_a_hi, &full_bar[s], ak, am);
            if (PASSES == 3) tma_load_2d(st + C::kATileBytes, &map_a_lo, &full_bar[s], ak, am);
            tma_load_2d_hint(stb, &map_b_hi, &full_bar[s], k0, n0, wpol);
            if (PASSES == 3) tma_load_2d_hint(stb + C::kBTileBytes, &map_b_lo, &full_bar[s], k0, n0, wpol);
          }
          ISC_TRACE(it == 0, 5);
          ISC_TRACE(it == 8, 6);
        }
      }
    }
    if (SK || EPI == EPI_GATE) {  // the cluster-wide barrier of the SK / gate epilogues: all threads of the cluster take part
      __syncwarp();
      cluster_sync_all();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0 && cta_rank == 0) {  // CG = 2: the leader issues for the pair
      constexpr uint32_t idesc = umma_idesc_bf16(CG * BM, BN);
      int it = 0, j = 0;
      for (int tile = walker; tile < num_tiles; tile += walkers, ++j) {
        const int as = j & 1;
        mbar_wait(&acc_empty[as], ((j >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        uint32_t accumulate = 0;
        uint32_t idesc_t = idesc;
        if (EPI == EPI_LSTM && BN == 256 && tile >= tiles_m * ep.lstm_wide) idesc_t = umma_idesc_bf16(CG * BM, 128);  // narrow tile
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          ISC_TRACE(it == 0, 2);
          ISC_TRACE(it == 8, 7);
          const uint32_t st = smem_u32(smem + s * C::kStageBytes);
          const uint32_t a_hi = st, a_lo = st + C::kATileBytes;
          const uint32_t b_hi = st + C::kPlanes * C::kATileBytes, b_lo = b_hi + C::kBTileBytes;
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
            const uint32_t a = (p == 1) ? a_lo : a_hi;  // hi·hi, lo·hi, hi·lo
            const uint32_t b = (p == 2) ? b_lo : b_hi;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // +32 B per 16-element K step inside the 128 B swizzle atom
              if (CG == 2)
                umma_bf16_pair(tmem_d, umma_desc_sw<2 * BK>(a + k * 32), umma_desc_sw<2 * BK>(b + k * 32), idesc_t, accumulate);
              else
                umma_bf16(tmem_d, umma_desc_sw<2 * BK>(a + k * 32), umma_desc_sw<2 * BK>(b + k * 32), idesc_t, accumulate);
              accumulate = 1;
            }
          }
          // frees the smem stage (in both CTAs of a pair) once the MMAs above have read it
          if (CG == 2) umma_commit_pair(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
        }
        // accumulator of this tile complete (signalled to the epilogue warps of both CTAs of a pair)
        if (CG == 2) umma_commit_pair(&acc_full[as]);
        else umma_commit(&acc_full[as]);
        ISC_TRACE(j < 2, 3 + j);
      }
    }
    if (SK || EPI == EPI_GATE) {
      __syncwarp();
      cluster_sync_all();
    }
  } else if (AF && warp >= 2 + EPI_WARPS) {
    // ===================== converters: fp32 A tile (staging) -> bf16 hi/lo operand tiles of the stage =====================
    // thread t owns 1 / PARTS of row r = t & 127: PER floats of the 128B-swizzled staging row -> PER bf16 in each of the
    // 64B-swizzled operand tiles
    constexpr int PARTS = CONV_WARPS / 4, PER = 32 / PARTS;
    const int t = threadIdx.x - (2 + EPI_WARPS) * 32;
    const int r = t & 127, hf = t >> 7;
    int it = 0;
    for (int tile = walker; tile < num_tiles; tile += walkers) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1;
        const int slot = it & 1;
        mbar_wait(&stg_full[slot], (it >> 1) & 1);
        ISC_TRACE(t == 0 && it == 8, 14);
        const uint8_t* src = stg + slot * C::kStgBytes + r * 128;
        float v[PER];
#pragma unroll
        for (int c = 0; c < PER / 4; ++c) {
          const float4 x = *reinterpret_cast<const float4*>(src + (((hf * (PER / 4) + c) ^ (r & 7)) << 4));
          v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        }
        // the fp32 values are in registers: hand the staging slot back to the producer now, not after the writes
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&stg_empty[slot])) : "memory");
        __align__(16) __nv_bfloat16 hh[PER], ll[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) split_bf16(v[i], hh[i], ll[i]);
        mbar_wait(&empty_bar[s], ph ^ 1);  // the MMAs that read this stage's previous contents are done
        uint8_t* dh = smem + s * C::kStageBytes + r * 64;
#pragma unroll
        for (int c = 0; c < PER / 8; ++c) {
          const int pos = ((hf * (PER / 8) + c) ^ ((r >> 1) & 3)) << 4;
          *reinterpret_cast<uint4*>(dh + pos) = reinterpret_cast<const uint4*>(hh)[c];
          if (PASSES == 3) *reinterpret_cast<uint4*>(dh + C::kATileBytes + pos) = reinterpret_cast<const uint4*>(ll)[c];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
          // a pair's MMAs are issued by the leader and read BOTH CTAs' A tiles: the peer's converters signal the leader's barrier
          if (CG == 2) mbar_arrive_cta(&full_bar[s], 0);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
        }
      }
    }
  } else if (EPI == EPI_LOGITS || EPI == EPI_LOGITS8) {
    constexpr int SEL_K = EPI == EPI_LOGITS8 ? 8 : 4;
    constexpr int SEL_REC = sel_rec(SEL_K);
    // ===================== epilogue: softmax partials + top candidates per (row, 128-column slice) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes this warp may touch: [32*quarter, +32)
    const int half = ew >> 2;      // which half of the tile's columns this warp scans
    constexpr int SLICE = BN / 2;
    float* bias_s = epi_smem + ew * SLICE;  // this warp's bias slice
    const LogitsSelect& sel = ep.sel;
    constexpr float kLog2e = 1.4426950408889634f;
    int j = 0;
    for (int tile = walker; tile < num_tiles; tile += walkers, ++j) {
      const int tn = tile % tiles_n;
      const int m0 = (tile / tiles_n) * (CG * BM) + cta_rank * BM, n_base = tn * BN + half * SLICE;
      const int as = j & 1;
      const int row = m0 + quarter * 32 + lane;
      for (int i = lane; i < SLICE; i += 32) bias_s[i] = (ep.bias && n_base + i < ep.N) ? __ldg(ep.bias + n_base + i) : 0.f;
      const int last_w = (sel.constraint && sel.last && row < ep.M) ? (int)__ldg(sel.last + row) : -1;
      __syncwarp();
      float mx = -INFINITY, sum = 0.f;
      float cv[SEL_K];
      int ci[SEL_K];
#pragma unroll
      for (int k = 0; k < SEL_K; ++k) {
        cv[k] = -INFINITY;
        ci[k] = 0x7fffffff;
      }
      mbar_wait(&acc_full[as], (j >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < SLICE / 32; ++cc) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * SLICE + cc * 32, v);
        const int nb = n_base + cc * 32;
        if (nb >= ep.N) break;  // warp-uniform: the rest of the slice is past the vocabulary
        float cm = -INFINITY;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          v[q] = (nb + q < ep.N) ? v[q] + bias_s[cc * 32 + q] : -INFINITY;
          cm = fmaxf(cm, v[q]);
        }
        const float mn = fmaxf(mx, cm);  // finite: column nb is inside the vocabulary
        float part = 0.f;
        const float mn2 = mn * kLog2e;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[q], kLog2e, -mn2)));
          part += e;
        }
        float scale;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(scale) : "f"((mx - mn) * kLog2e));  // mx = -inf -> 0
        sum = fmaf(sum, scale, part);
        mx = mn;
        if (SEL_K == 4) {
          // Branch-free: every lane owns a different row, so a per-column "does it enter the list" branch diverges on
          // nearly every column and the warp pays branch + reconvergence 128 times per slice (ncu: more than half of
          // this epilogue's stall samples sat on FSETP / BRA / BSYNC, and the kernel was epilogue-bound). The
          // select-based insertion of a value that does not qualify (or of -inf for a masked column) changes nothing,
          // so it simply runs for all 32 columns. The special ids only occur in the chunk that holds them.
          const int max_special = max(sel.pad_id, max(sel.sos_id, sel.unk_id));
          if (sel.mask_special && nb <= max_special) {  // warp-uniform
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const int n = nb + q;
              if (n == sel.pad_id || n == sel.sos_id || n == sel.unk_id) v[q] = -INFINITY;
            }
          }
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int n = nb + q;
            topk_insert<SEL_K>(cv, ci, n == last_w ? -INFINITY : v[q], n);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            if (v[q] > cv[SEL_K - 1]) {
              const int n = nb + q;
              const bool masked = (sel.mask_special && (n == sel.pad_id || n == sel.sos_id || n == sel.unk_id)) || n == last_w;
              if (!masked) topk_insert<SEL_K>(cv, ci, v[q], n);
            }
          }
        }
      }
      // the accumulator stage is drained: release it to the (leader's) MMA warp before the stores
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cta(&acc_empty[as], 0);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
      }
      if (row < ep.M) {
        float4* r4 = reinterpret_cast<float4*>(sel.rec + ((long long)row * sel.np + (tn * 2 + half)) * SEL_REC);
        r4[0] = make_float4(mx, sum, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < SEL_K; k += 4) {
          r4[1 + k / 4] = make_float4(cv[k], cv[k + 1], cv[k + 2], cv[k + 3]);
          r4[1 + SEL_K / 4 + k / 4] = make_float4(__int_as_float(ci[k]), __int_as_float(ci[k + 1]), __int_as_float(ci[k + 2]),
                                                  __int_as_float(ci[k + 3]));
        }
      }
      __syncwarp();  // bias slice is rewritten for the next tile
    }
  } else if (EPI == EPI_GATE) {
    // ===================== epilogue: gate weight + mixed context (cluster of 4 = the four column tiles of a row block) =====
    const GateEpilogue& G = ep.gate;
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes this warp may touch: [32*quarter, +32)
    const int half = ew >> 2;      // which half of the tile's columns this warp drains
    float4* sc = reinterpret_cast<float4*>(epi_smem) + ew * 32 * 8;  // 32 rows x 8 float4, slot ^= row & 7
    float* mine = reinterpret_cast<float*>(sc);  // after the chunks: [0, 32) this warp's row sums (read by the cluster), [32, 64) w
    const int r_off = lane >> 3, c4 = lane & 7;
    const int tile = walker;  // one tile per CTA: grid = tiles, cluster rank = column tile
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const int row0 = m0 + quarter * 32 + r_off;  // this lane's rows: row0 + 4*i, i = 0..7
    int nvalid = (ep.M - row0 + 3) >> 2;
    nvalid = nvalid < 0 ? 0 : (nvalid > 8 ? 8 : nvalid);
    float psum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) psum[i] = 0.f;
    mbar_wait(&acc_full[0], 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < BN / 64; ++cc) {
      const int c = half * (BN / 64) + cc;
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, v);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        sc[lane * 8 + (q ^ (lane & 7))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      __syncwarp();
      const int n = n0 + c * 32 + c4 * 4;  // N = 512: every column is valid
      const float4 b4 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(G.alpha + n));
      float4 add4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        add4[i] = (i < nvalid && ep.addmat) ? __ldg(reinterpret_cast<const float4*>(ep.addmat + (long long)(row0 + 4 * i) * ep.ld_addmat + n))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + r_off;
        float4 x = sc[r * 8 + (c4 ^ (r & 7))];
        x.x = act_ct<ACT_TANH>(x.x + b4.x + add4[i].x);
        x.y = act_ct<ACT_TANH>(x.y + b4.y + add4[i].y);
        x.z = act_ct<ACT_TANH>(x.z + b4.z + add4[i].z);
        x.w = act_ct<ACT_TANH>(x.w + b4.w + add4[i].w);
        psum[i] += a4.x * x.x + a4.y * x.y + a4.z * x.z + a4.w * x.w;
      }
      __syncwarp();  // staging tile is reused by the next chunk
    }
    tcgen05_fence_before();
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // the eight lanes that share a row
      psum[i] += __shfl_xor_sync(0xffffffffu, psum[i], 1);
      psum[i] += __shfl_xor_sync(0xffffffffu, psum[i], 2);
      psum[i] += __shfl_xor_sync(0xffffffffu, psum[i], 4);
    }
    if (c4 == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[r_off + 4 * i] = psum[i];
    }
    __syncwarp();
    cluster_sync_all();
    // gate weight of row quarter * 32 + lane: the eight partial sums (4 column tiles x 2 halves) in a fixed order, so
    // that the four CTAs (and both half-warps) of a row block arrive at the same bits
    float e = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
      for (int h = 0; h < 2; ++h) e += ld_dsmem_f32(smem_u32(epi_smem + (h * 4 + (ew & 3)) * 1024 + lane), d);  // warp (half h, this quarter)
    const float wl = sigmoid_accurate(e + __ldg(G.alpha_b));
    mine[32 + lane] = wl;
    if (G.gate_w && n0 == 0 && half == 0 && m0 + quarter * 32 + lane < ep.M)
      G.gate_w[(long long)(m0 + quarter * 32 + lane) * G.ld_gate_w] = wl;
    __syncwarp();
#pragma unroll 1
    for (int cc = 0; cc < BN / 64; ++cc) {
      const int n = n0 + (half * (BN / 64) + cc) * 32 + c4 * 4;
      float4 cv[8], sv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < nvalid) {
          const float* p = G.cs + (long long)(row0 + 4 * i) * G.ld_cs + n;
          cv[i] = __ldg(reinterpret_cast<const float4*>(p));
          sv[i] = __ldg(reinterpret_cast<const float4*>(p + H));
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < nvalid) {
          const long long row = row0 + 4 * i;
          const float w = mine[32 + r_off + 4 * i];
          float4 o;
          o.x = w * cv[i].x + (1.f - w) * sv[i].x;
          o.y = w * cv[i].y + (1.f - w) * sv[i].y;
          o.z = w * cv[i].z + (1.f - w) * sv[i].z;
          o.w = w * cv[i].w + (1.f - w) * sv[i].w;
          if (ep.c) *reinterpret_cast<float4*>(ep.c + row * ep.ldc + n) = o;
          if (ep.hi) {
            __nv_bfloat16 h[4], l[4];
            split_bf16(o.x, h[0], l[0]);
            split_bf16(o.y, h[1], l[1]);
            split_bf16(o.z, h[2], l[2]);
            split_bf16(o.w, h[3], l[3]);
            *reinterpret_cast<uint2*>(ep.hi + row * ep.ldp + n) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
            if (ep.lo) *reinterpret_cast<uint2*>(ep.lo + row * ep.ldp + n) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
          }
        }
      }
    }
  } else if (EPI == EPI_LSTM) {

    // ===================== epilogue: LSTM cell on the gate pre-activations (tile columns = [i f g o] x 16 units per group) =====
    // (Measured and dropped in round 2: routing c_prev, the addend rows and the results through a per-warp shared-memory
    // tile so that the warp's global accesses are coalesced — 4x fewer L1 wavefronts, but the shuffles, warp barriers and
    // smem round trips cost more: attention-LSTM GEMM 48 -> 59 us, language-LSTM GEMM 52 -> 56 us.)
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes this warp may touch: [32*quarter, +32)
    const int half = ew >> 2;      // which half of the tile's 16-unit groups
    const LstmEpilogue& L = ep.lstm;
    const int HH = ep.N >> 2;      // hidden size
    int j = 0;
    for (int tile = walker; tile < num_tiles; tile += walkers, ++j) {
      int m0, ut, width;
      lstm_tile<BN, CG * BM>(tile, tiles_m, tiles_n, ep.lstm_wide, m0, ut, width);
      m0 += cta_rank * BM;
      const int nsub = width >> 7;  // 16-unit groups per warp: 1 (128-column tile) or 2 (256-column tile)
      const int as = j & 1;
      const long long row = m0 + quarter * 32 + lane;
      const bool live = row < ep.M;
      const long long src = (live && L.parent) ? L.parent[row] : row;
      const float* radd = ep.rowadd ? ep.rowadd + (long long)((unsigned)row / (unsigned)ep.rows_per_group) * ep.ld_rowadd : nullptr;
      const float* gat = nullptr;
      if (live && L.gather_tab) {
        long long tok = L.gather_idx[row];
        tok = tok < 0 ? 0 : (tok >= L.gather_rows ? L.gather_rows - 1 : tok);
        gat = L.gather_tab + tok * ep.N;
      }
      for (int ss = 0; ss < nsub; ++ss) {
        const int grp = half * nsub + ss;
        const int u0 = ut + grp * 16;  // first hidden unit of this thread in this pass
        float cp[16];
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(L.c_prev + src * HH + u0) + q);
            cp[4 * q] = t.x; cp[4 * q + 1] = t.y; cp[4 * q + 2] = t.z; cp[4 * q + 3] = t.w;
          }
        }
        if (ss == 0) {
          mbar_wait(&acc_full[as], (j >> 1) & 1);
          tcgen05_fence_after();
          ISC_TRACE(threadIdx.x == 64 && j < 2, 8 + 2 * j);
        }
        float vif[32], vgo[32];  // [i(16) | f(16)], [g(16) | o(16)]
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + grp * 64;
        tmem_ld32(tcol, vif);
        tmem_ld32(tcol + 32, vgo);
        if (ss == nsub - 1) {
          // the accumulator stage is drained: release it to the MMA warp before the math and the stores
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cta(&acc_empty[as], 0);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
          }
        }
        if (live) {
          // same association as the unfused path: (acc + bias) + rowadd, then nn.LSTMCell's pointwise math
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float* v = (g < 2 ? vif : vgo) + (g & 1) * 16;
            const int col = g * HH + u0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (ep.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col) + q);
                v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
              }
              if (radd) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(radd + col) + q);
                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
              }
              if (gat) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(gat + col) + q);
                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
              }
            }
          }
          float hn[16], cn[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
#if ISC_CELL_FAST
            // MUFU-based gates (sigmoid_fast, tanh_ex2: ~3e-7 / 2e-7 error, far inside the split-bf16 GEMM's own 8e-6)
            cn[u] = sigmoid_fast(vif[16 + u]) * cp[u] + sigmoid_fast(vif[u]) * tanh_ex2(vgo[u]);
            hn[u] = sigmoid_fast(vgo[16 + u]) * tanh_ex2(cn[u]);
#else
            cn[u] = sigmoid_accurate(vif[16 + u]) * cp[u] + sigmoid_accurate(vif[u]) * tanhf(vgo[u]);
            hn[u] = sigmoid_accurate(vgo[16 + u]) * tanhf(cn[u]);
#endif
          }
          float4* co = reinterpret_cast<float4*>(L.c_out + row * HH + u0);
          float4* ho = reinterpret_cast<float4*>(L.h_out + row * HH + u0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            co[q] = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
            ho[q] = make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
          }
          if (L.x_hi) {
            if (L.mask) {
              const uint4 k = __ldg(reinterpret_cast<const uint4*>(L.mask + row * HH + u0));
              const unsigned char* kb = reinterpret_cast<const unsigned char*>(&k);
#pragma unroll
              for (int u = 0; u < 16; ++u) hn[u] *= kb[u] ? L.scale : 0.f;
            }
            __align__(16) __nv_bfloat16 hh[16], ll[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) split_bf16(hn[u], hh[u], ll[u]);
            uint4* dh = reinterpret_cast<uint4*>(L.x_hi + row * L.ldx + L.x_col + u0);
            dh[0] = reinterpret_cast<const uint4*>(hh)[0];
            dh[1] = reinterpret_cast<const uint4*>(hh)[1];
            if (L.x_lo) {
              uint4* dl = reinterpret_cast<uint4*>(L.x_lo + row * L.ldx + L.x_col + u0);
              dl[0] = reinterpret_cast<const uint4*>(ll)[0];
              dl[1] = reinterpret_cast<const uint4*>(ll)[1];
            }
          }
        }
      }
      ISC_TRACE(threadIdx.x == 64 && j < 2, 9 + 2 * j);
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> swizzled smem -> global =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes this warp may touch: [32*quarter, +32)
    const int half = ew >> 2;      // which half of the tile's columns this warp drains
    float4* sc = reinterpret_cast<float4*>(epi_smem) + ew * 32 * 8;  // 32 rows x 8 float4, slot ^= row & 7
    const int r_off = lane >> 3, c4 = lane & 7;  // transposed view: 4 rows x 8 float4 per warp access
    const bool c_vec = ep.c && ((reinterpret_cast<uintptr_t>(ep.c) & 15) == 0) && ((ep.ldc & 3) == 0);
    const bool p_vec = ep.hi && ((reinterpret_cast<uintptr_t>(ep.hi) & 7) == 0) && ((ep.ldp & 3) == 0) &&
                       (!ep.lo || (reinterpret_cast<uintptr_t>(ep.lo) & 7) == 0);
    const bool radd_vec = ep.rowadd && ((reinterpret_cast<uintptr_t>(ep.rowadd) & 15) == 0) && ((ep.ld_rowadd & 3) == 0);
    const bool madd_vec = ep.addmat && ((reinterpret_cast<uintptr_t>(ep.addmat) & 15) == 0) && ((ep.ld_addmat & 3) == 0);
    const bool bias_vec = ep.bias && ((reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0);
    int j = 0;
    if (SK) {
      // ===================== split-K epilogue (one tile per cluster, one K slice per CTA) =====================
      // 1. every CTA parks its raw fp32 partial tile in its own shared memory — the operand stages are free: each stage
      //    was consumed by the MMAs the accumulator barrier has just reported complete, and a CTA runs one K slice only;
      // 2. cluster barrier; 3. CTA r sums rows [r, r + 1) * BM / split_k of all the partial tiles through distributed
      // shared memory, in slice order, and finishes them (bias, addends, fp32 / bf16-plane stores).
      const int tile = walker;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      float4* part = reinterpret_cast<float4*>(smem);  // [BM][BN / 4] float4, slot ^= row & 31
      mbar_wait(&acc_full[0], 0);
      tcgen05_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN / 64; ++cc) {
        const int c = half * (BN / 64) + cc;
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, v);
        const int r = quarter * 32 + lane;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          part[r * (BN / 4) + ((c * 8 + q) ^ (r & 31))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      tcgen05_fence_before();
      __syncwarp();
      cluster_sync_all();
      const int rows_cta = BM / split_k, rows_warp = rows_cta / EPI_WARPS;  // split_k in {2, 4, 8}: 8 / 4 / 2 rows per warp
      const int n = n0 + lane * 4;
      const bool full4 = n + 3 < ep.N;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ep.bias && n < ep.N) b4 = (full4 && bias_vec) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : load4_guarded(ep.bias + n, n, ep.N);
#pragma unroll 1
      for (int k = 0; k < rows_warp; k += 2) {
        float4 x[2], add[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = ks * rows_cta + ew * rows_warp + k + u;
          const long long row = m0 + r;
          const uint32_t local = smem_u32(part + r * (BN / 4) + (lane ^ (r & 31)));
          add[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < ep.M && n < ep.N) {
            if (ep.addmat) {
              const float* mp = ep.addmat + row * ep.ld_addmat + n;
              add[u] = (full4 && madd_vec) ? __ldg(reinterpret_cast<const float4*>(mp)) : load4_guarded(mp, n, ep.N);
            }
            if (ep.rowadd) {
              const float* rp = ep.rowadd + (long long)((unsigned)row / (unsigned)ep.rows_per_group) * ep.ld_rowadd + n;
              const float4 t = (full4 && radd_vec) ? __ldg(reinterpret_cast<const float4*>(rp)) : load4_guarded(rp, n, ep.N);
              add[u].x += t.x; add[u].y += t.y; add[u].z += t.z; add[u].w += t.w;
            }
          }
          float4 t[8];  // all slices requested up front, summed in slice order: the sum is the same on every run
#pragma unroll
          for (int sl = 0; sl < 8; ++sl)
            if (sl < split_k) t[sl] = ld_dsmem_f4(local, sl);
          x[u] = t[0];
#pragma unroll
          for (int sl = 1; sl < 8; ++sl)
            if (sl < split_k) { x[u].x += t[sl].x; x[u].y += t[sl].y; x[u].z += t[sl].z; x[u].w += t[sl].w; }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = ks * rows_cta + ew * rows_warp + k + u;
          const long long row = m0 + r;
          if (row < ep.M && n < ep.N) {
            float4 y = x[u];
            y.x += b4.x; y.y += b4.y; y.z += b4.z; y.w += b4.w;
            y.x += add[u].x; y.y += add[u].y; y.z += add[u].z; y.w += add[u].w;
            if (ep.c) {
              float* dst = ep.c + row * ep.ldc + n;
              if (full4 && c_vec) {
                *reinterpret_cast<float4*>(dst) = y;
              } else {
                dst[0] = y.x;
                if (n + 1 < ep.N) dst[1] = y.y;
                if (n + 2 < ep.N) dst[2] = y.z;
                if (n + 3 < ep.N) dst[3] = y.w;
              }
            }
            if (ep.hi) {
              __nv_bfloat16 h[4], l[4];
              split_bf16(y.x, h[0], l[0]);
              split_bf16(y.y, h[1], l[1]);
              split_bf16(y.z, h[2], l[2]);
              split_bf16(y.w, h[3], l[3]);
              __nv_bfloat16* dh = ep.hi + row * ep.ldp + n;
              __nv_bfloat16* dl = ep.lo ? ep.lo + row * ep.ldp + n : nullptr;
              if (full4 && p_vec) {
                *reinterpret_cast<uint2*>(dh) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
                if (dl) *reinterpret_cast<uint2*>(dl) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
              } else {
                for (int q = 0; q < 4 && n + q < ep.N; ++q) {
                  dh[q] = h[q];
                  if (dl) dl[q] = l[q];
                }
              }
            }
          }
        }
      }
    }
    for (int tile = walker; !SK && tile < num_tiles; tile += walkers, ++j) {
      const int m0 = (tile / tiles_n) * (CG * BM) + cta_rank * BM, n0 = (tile % tiles_n) * BN;
      const int as = j & 1;
      const int row0 = m0 + quarter * 32 + r_off;  // this lane's rows: row0 + 4*i, i = 0..7
      int nvalid = (ep.M - row0 + 3) >> 2;
      nvalid = nvalid < 0 ? 0 : (nvalid > 8 ? 8 : nvalid);
      const float* radd[8];
      if (ep.rowadd) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          radd[i] = ep.rowadd + (long long)((unsigned)(row0 + 4 * i) / (unsigned)ep.rows_per_group) * ep.ld_rowadd;
      }
      // the bias of this lane's columns does not depend on the MMAs: fetched before the wait for the accumulator (phase
      // trace: with one tile per SM the epilogue is 40 % of the kernel and sat on this load once per 32-column chunk)
      // (not in the fp32-A kernels: their 448 threads leave 128 registers per thread and the epilogue would spill)
      constexpr int NB4 = AF ? 1 : BN / 64;
      float4 bias4[NB4];
      if (!AF) {
#pragma unroll
        for (int cc = 0; cc < NB4; ++cc) {
          const int n = n0 + (half * (BN / 64) + cc) * 32 + c4 * 4;
          bias4[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias && n < ep.N)
            bias4[cc] = (n + 3 < ep.N && bias_vec) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : load4_guarded(ep.bias + n, n, ep.N);
        }
      }
      mbar_wait(&acc_full[as], (j >> 1) & 1);
      tcgen05_fence_after();
      ISC_TRACE(threadIdx.x == 64 && j < 2, 8 + 2 * j);
#pragma unroll 1
      for (int cc = 0; cc < BN / 64; ++cc) {
        const int c = half * (BN / 64) + cc;
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + c * 32, v);
        ISC_TRACE(threadIdx.x == 64 && j == 0 && cc == 0, 3);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sc[lane * 8 + (q ^ (lane & 7))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        ISC_TRACE(threadIdx.x == 64 && j == 0 && cc == 0, 4);
        const int n = n0 + c * 32 + c4 * 4;
        if (n < ep.N) {
          const bool full4 = n + 3 < ep.N;
          float4 b4;
          if (AF) {
            b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias) b4 = (full4 && bias_vec) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : load4_guarded(ep.bias + n, n, ep.N);
          } else {
            b4 = bias4[0];  // (a select chain, not an indexed register array: the chunk loop is not unrolled)
#pragma unroll
            for (int k = 1; k < NB4; ++k)
              if (cc == k) b4 = bias4[k];
          }
          // the addends of all eight rows are requested up front: issued one by one in front of their use, each load's
          // latency is exposed (ncu: the epilogue sat on the dependent FADDs). One register array serves rowadd or, when
          // there is no rowadd, addmat (no caller passes both; if one did, addmat is loaded in the loop).
          float4 add4[8];
          const bool pre_madd = ep.addmat && !ep.rowadd;
          if (ep.rowadd) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              add4[i] = i < nvalid ? ((full4 && radd_vec) ? __ldg(reinterpret_cast<const float4*>(radd[i] + n)) : load4_guarded(radd[i] + n, n, ep.N))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (pre_madd) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float* mp = ep.addmat + (long long)(row0 + 4 * i) * ep.ld_addmat + n;
              add4[i] = i < nvalid ? ((full4 && madd_vec) ? __ldg(reinterpret_cast<const float4*>(mp)) : load4_guarded(mp, n, ep.N))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < nvalid) {
              const int r = 4 * i + r_off;
              const long long row = row0 + 4 * i;
              float4 x = sc[r * 8 + (c4 ^ (r & 7))];
              x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
              if (ep.rowadd || pre_madd) {
                const float4 t = add4[i];
                x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
              }
              if (ep.addmat && !pre_madd) {
                const float* mp = ep.addmat + row * ep.ld_addmat + n;
                const float4 t = (full4 && madd_vec) ? __ldg(reinterpret_cast<const float4*>(mp)) : load4_guarded(mp, n, ep.N);
                x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
              }
              x.x = act_ct<ACT>(x.x); x.y = act_ct<ACT>(x.y); x.z = act_ct<ACT>(x.z); x.w = act_ct<ACT>(x.w);
              if (ep.zero_border) {
                const int g = (int)(row & 255), gy = g >> 4, gx = g & 15;
                if (gy == 0 || gy == 15 || gx == 0 || gx == 15) x = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              if (ep.c) {
                float* dst = ep.c + row * ep.ldc + n;
                if (full4 && c_vec) {
                  *reinterpret_cast<float4*>(dst) = x;
                } else {
                  dst[0] = x.x;
                  if (n + 1 < ep.N) dst[1] = x.y;
                  if (n + 2 < ep.N) dst[2] = x.z;
                  if (n + 3 < ep.N) dst[3] = x.w;
                }
              }
              if (H16) {
                float4 y = x;
                if (ep.h16.expneg2) { y.x = __expf(-2.0f * y.x); y.y = __expf(-2.0f * y.y); y.z = __expf(-2.0f * y.z); y.w = __expf(-2.0f * y.w); }
                y.x *= ep.h16.scale; y.y *= ep.h16.scale; y.z *= ep.h16.scale; y.w *= ep.h16.scale;
                const float lo_ = fminf(fminf(y.x, y.y), fminf(y.z, y.w)), hi_ = fmaxf(fmaxf(y.x, y.y), fmaxf(y.z, y.w));
                __half* d16 = static_cast<__half*>(ep.h16.out) + row * ep.h16.ld + n;
                if (full4 && ((reinterpret_cast<uintptr_t>(d16) & 7) == 0)) {
                  const __half2 a = __floats2half2_rn(y.x, y.y), b = __floats2half2_rn(y.z, y.w);
                  *reinterpret_cast<uint2*>(d16) = make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
                  if (!(lo_ >= ep.h16.vmin && hi_ <= ep.h16.vmax)) atomicOr(ep.h16.flags + (unsigned)row / (unsigned)ep.h16.rows_per_flag, 1);
                } else {
                  const float yv[4] = {y.x, y.y, y.z, y.w};
                  for (int q = 0; q < 4 && n + q < ep.N; ++q) {
                    d16[q] = __float2half_rn(yv[q]);
                    if (!(yv[q] >= ep.h16.vmin && yv[q] <= ep.h16.vmax)) atomicOr(ep.h16.flags + (unsigned)row / (unsigned)ep.h16.rows_per_flag, 1);
                  }
                }
              }
              if (ep.hi) {
                __nv_bfloat16 h[4], l[4];
                split_bf16(x.x, h[0], l[0]);
                split_bf16(x.y, h[1], l[1]);
                split_bf16(x.z, h[2], l[2]);
                split_bf16(x.w, h[3], l[3]);
                __nv_bfloat16* dh = ep.hi + row * ep.ldp + n;
                __nv_bfloat16* dl = ep.lo ? ep.lo + row * ep.ldp + n : nullptr;
                if (full4 && p_vec) {
                  // packed in registers (taking the arrays' address sent them through the local-memory stack)
                  *reinterpret_cast<uint2*>(dh) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
                  if (dl) *reinterpret_cast<uint2*>(dl) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
                } else {
                  for (int q = 0; q < 4 && n + q < ep.N; ++q) {
                    dh[q] = h[q];
                    if (dl) dl[q] = l[q];
                  }
                }
              }
            }
          }
        }
        __syncwarp();  // staging tile is reused by the next chunk
        ISC_TRACE(threadIdx.x == 64 && j == 0 && cc < 2, 5 + cc);
      }
      // release the accumulator stage to the (leader's) MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cta(&acc_empty[as], 0);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
      }
      ISC_TRACE(threadIdx.x == 64 && j < 2, 9 + 2 * j);
    }
  }

  tcgen05_fence_before();
  if (CG == 2 || SK || EPI == EPI_GATE) cluster_sync_all();  // no CTA of the pair leaves while the other may still signal into its smem (SK: read it)
  else __syncthreads();
  ISC_TRACE(threadIdx.x == 0, 15);
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// bf16 plane [rows, cols] with leading dimension ld (elements) -> 2D map, box box_k x box_rows, swizzle = row bytes.
static int make_map(CUtensorMap* map, const __nv_bfloat16* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int box_k) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable from the driver");
    return ISC_ERR_DEVICE;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("gemm_tc: operand plane must be 16-byte aligned with a 16-byte multiple row pitch");
    return ISC_ERR_ARG;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld cols %lld ld %lld)", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
    return ISC_ERR_ARG;
  }
  return 0;
}

struct Maps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

// fp32 matrix [rows, cols] -> 2D map, box 32 floats (one 128-byte swizzle row) x 128 rows
static int make_map_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable from the driver");
    return ISC_ERR_DEVICE;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 4) % 16 != 0) {
    set_error("gemm_tc: fp32 operand must be 16-byte aligned with a 16-byte multiple row pitch");
    return ISC_ERR_ARG;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, static_cast<cuuint32_t>(BM)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rows %lld cols %lld ld %lld)", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
    return ISC_ERR_ARG;
  }
  return 0;
}

template <int PASSES, int BN, int ACT, int EPI, int CG, int AF = 0, int H16 = 0, int SK = 0>
static int launch_kernel(const Maps& m, const EpiParams& ep, int grid, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<PASSES, BN, ACT, EPI, CG, AF, H16, SK>;
  constexpr int smem = Cfg<PASSES, BN, CG, AF>::kSmemBytes;
  constexpr int NUM_THREADS = Cfg<PASSES, BN, CG, AF>::kThreads;
  ISC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (CG == 2 || SK || EPI == EPI_GATE) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = SK ? ep.split_k : (EPI == EPI_GATE ? 4 : 2);
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    ISC_CUDA(cudaLaunchKernelEx(&cfg, kern, m.a_hi, m.a_lo, m.b_hi, m.b_lo, ep));
  } else {
    ISC_CUDA(launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), smem, stream, m.a_hi, m.a_lo, m.b_hi, m.b_lo, ep));
  }
  ISC_LAUNCH_CHECK();
  return 0;
}

template <int PASSES, int BN, int CG>
static int make_maps(Maps& m, const Operand& A, const Operand& W, int M, int N, int K) {
  using C = Cfg<PASSES, BN, CG>;
  constexpr int BK = C::kBK;
  ISC_TRY(make_map(&m.a_hi, A.hi, M, K, A.ldp, BM, BK));
  ISC_TRY(make_map(&m.b_hi, W.hi, N, K, W.ldp, C::kBRows, BK));
  if (PASSES == 3) {
    ISC_TRY(make_map(&m.a_lo, A.lo, M, K, A.ldp, BM, BK));
    ISC_TRY(make_map(&m.b_lo, W.lo, N, K, W.ldp, C::kBRows, BK));
  } else {
    m.a_lo = m.a_hi;
    m.b_lo = m.b_hi;
  }
  return 0;
}

// ---- split-K: GEMMs with few tiles and a long K loop (the training path's M = 256 step GEMMs: 8-32 tiles of 16-32
// k-blocks; its weight-gradient contractions over batch x regions rows: 16-64 tiles of 784 k-blocks) leave most SMs idle
// while a few CTAs pull their whole operand panels through one SM's L2 port. Such a tile is given to a cluster of 2 / 4 / 8
// CTAs, one K slice each, which sum their partial tiles through distributed shared memory.
// Returns the cluster size for an M x N x K GEMM on 128 x 128 tiles with 64-wide k-blocks (1 = do not split).
// ISC_SPLITK=0 switches it off.
static int splitk_slices(int M, int N, int K) {
  static const bool on = !(getenv("ISC_SPLITK") && atoi(getenv("ISC_SPLITK")) == 0);
  if (!on) return 1;
  const int tiles = ((M + BM - 1) / BM) * ((N + 127) / 128), num_kb = (K + 63) / 64;
  if (tiles * 2 > num_sms() || num_kb < 12) return 1;
  int s = 8;  // at least four k-blocks per slice: below that, fill and the exchange cost more than the slices save
  while (s > 1 && (s * tiles > num_sms() || s * 4 > num_kb)) s >>= 1;
  return s;
}

// persistent grid: one CTA (CG = 2: one CTA pair) per SM (pair) walks the tile list
template <int BN, int CG>
static int persistent_grid(int M, int N) {
  const int tiles = ((N + BN - 1) / BN) * ((M + CG * BM - 1) / (CG * BM));
  const int slots = num_sms() / CG;
  return CG * (tiles < slots ? tiles : slots);
}

template <int PASSES, int BN, int CG>
static int launch(const Operand& A, const Operand& W, const Dest& Cd, int M, int N, int K, const Epilogue& e,
                  cudaStream_t stream) {
  Maps m;
  ISC_TRY((make_maps<PASSES, BN, CG>(m, A, W, M, N, K)));
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = e.bias;
  ep.rowadd = e.rowadd;
  ep.ld_rowadd = e.ld_rowadd;
  ep.rows_per_group = e.rows_per_group > 0 ? e.rows_per_group : 1;
  ep.addmat = e.addmat;
  ep.ld_addmat = e.ld_addmat;
  ep.act = e.act;
  ep.c = Cd.f32;
  ep.ldc = Cd.ld;
  ep.hi = Cd.hi;
  ep.lo = Cd.lo;
  ep.ldp = Cd.ldp;
  ep.h16 = Cd.h16;
  ep.M = M;
  ep.N = N;
  ep.K = K;
  const int grid = persistent_grid<BN, CG>(M, N);
  ProfScope ps(ISC_K_GEMM_TC, 2.0 * M * N * K * PASSES, stream);
  if (PASSES == 3 && BN == 128 && CG == 1 && e.act == ACT_NONE && !ep.h16.out) {
    const int slices = splitk_slices(M, N, K);
    if (slices > 1) {
      ep.split_k = slices;
      return launch_kernel<PASSES, 128, ACT_NONE, EPI_STD, 1, 0, 0, (PASSES == 3 && BN == 128 && CG == 1) ? 1 : 0>(m, ep, grid * slices, stream);
    }
  }
  if (ep.h16.out) {  // fp16 copy of the result (the prologue's projected features): only the two activations it is used with
    if (e.act == ACT_RELU) return launch_kernel<PASSES, BN, ACT_RELU, EPI_STD, CG, 0, 1>(m, ep, grid, stream);
    if (e.act == ACT_EXPNEG2_RELU) return launch_kernel<PASSES, BN, ACT_EXPNEG2_RELU, EPI_STD, CG, 0, 1>(m, ep, grid, stream);
    set_error("gemm_tc: fp16 output is only built for the ReLU / exp(-2 ReLU) epilogues");
    return ISC_ERR_ARG;
  }
  switch (e.act) {
    case ACT_RELU: return launch_kernel<PASSES, BN, ACT_RELU, EPI_STD, CG>(m, ep, grid, stream);
    case ACT_TANH: return launch_kernel<PASSES, BN, ACT_TANH, EPI_STD, CG>(m, ep, grid, stream);
    case ACT_EXPNEG2_RELU: return launch_kernel<PASSES, BN, ACT_EXPNEG2_RELU, EPI_STD, CG>(m, ep, grid, stream);
    case ACT_SIGMOID: return launch_kernel<PASSES, BN, ACT_SIGMOID, EPI_STD, CG>(m, ep, grid, stream);
    default: return launch_kernel<PASSES, BN, ACT_NONE, EPI_STD, CG>(m, ep, grid, stream);
  }
}

template <int PASSES, int CG>
static int launch_logits(const Operand& A, const Operand& W, int M, int N, int K, const float* bias,
                         const LogitsSelect& sel, cudaStream_t stream) {
  constexpr int BN = 256;
  Maps m;
  ISC_TRY((make_maps<PASSES, BN, CG>(m, A, W, M, N, K)));
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = bias;
  ep.rows_per_group = 1;
  ep.M = M;
  ep.N = N;
  ep.K = K;
  ep.sel = sel;
  const int grid = persistent_grid<BN, CG>(M, N);
  ProfScope ps(ISC_K_GEMM_TC, 2.0 * M * N * K * PASSES, stream);
  if (sel.k_sel == 8) return launch_kernel<PASSES, BN, ACT_NONE, EPI_LOGITS8, CG>(m, ep, grid, stream);
  return launch_kernel<PASSES, BN, ACT_NONE, EPI_LOGITS, CG>(m, ep, grid, stream);
}

// A given as fp32 (split into operand planes inside the kernel), 128x256 tiles; CG = 2: 256x256 tiles on CTA pairs — each
// CTA converts its own 128 rows of A and loads half of the B rows, which takes the kernel from L2 -> SM port bound
// (16 KB of fp32 A + 32 KB of B planes per 0.4 us of MMAs and SM) to MMA bound
template <int PASSES, int CG>
static int launch_af32(const float* A, int64_t lda, const Operand& W, const Dest& Cd, int M, int N, int K, const Epilogue& e,
                       cudaStream_t stream) {
  constexpr int BN = 256;
  using C = Cfg<PASSES, BN, CG, 1>;
  Maps m;
  ISC_TRY(make_map_f32(&m.a_hi, A, M, K, lda));
  m.a_lo = m.a_hi;
  ISC_TRY(make_map(&m.b_hi, W.hi, N, K, W.ldp, C::kBRows, C::kBK));
  if (PASSES == 3) ISC_TRY(make_map(&m.b_lo, W.lo, N, K, W.ldp, C::kBRows, C::kBK));
  else m.b_lo = m.b_hi;
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = e.bias;
  ep.rowadd = e.rowadd;
  ep.ld_rowadd = e.ld_rowadd;
  ep.rows_per_group = e.rows_per_group > 0 ? e.rows_per_group : 1;
  ep.addmat = e.addmat;
  ep.ld_addmat = e.ld_addmat;
  ep.act = e.act;
  ep.c = Cd.f32;
  ep.ldc = Cd.ld;
  ep.hi = Cd.hi;
  ep.lo = Cd.lo;
  ep.ldp = Cd.ldp;
  ep.h16 = Cd.h16;
  ep.M = M;
  ep.N = N;
  ep.K = K;
  const int grid = persistent_grid<BN, CG>(M, N);
  ProfScope ps(ISC_K_GEMM_TC, 2.0 * M * N * K * PASSES, stream);
  if (ep.h16.out) {
    if (e.act == ACT_RELU) return launch_kernel<PASSES, BN, ACT_RELU, EPI_STD, CG, 1, 1>(m, ep, grid, stream);
    set_error("gemm_tc_af32: fp16 output is only built for the ReLU epilogue");
    return ISC_ERR_ARG;
  }
  if (e.act == ACT_RELU) return launch_kernel<PASSES, BN, ACT_RELU, EPI_STD, CG, 1>(m, ep, grid, stream);
  return launch_kernel<PASSES, BN, ACT_NONE, EPI_STD, CG, 1>(m, ep, grid, stream);
}

// 3x3 convolution (stride 1, zero padding 1) over 16x16-gridded rows as one GEMM, 128x256 tiles. A: fp32 [M][C]
// (AF32 != 0, split in-kernel) or bf16 planes [M][C]; W planes [N][9*C].
template <int PASSES, int AF32>
static int launch_conv(const float* A32, const Operand& A, int64_t lda, const Operand& W, const Dest& Cd, int M, int N, int C_in,
                       const Epilogue& e, int zero_border, cudaStream_t stream) {
  constexpr int BN = 256;
  using C = Cfg<PASSES, BN, 1, AF32>;
  Maps m;
  if (AF32) {
    ISC_TRY(make_map_f32(&m.a_hi, A32, M, C_in, lda));
    m.a_lo = m.a_hi;
  } else {
    ISC_TRY(make_map(&m.a_hi, A.hi, M, C_in, A.ldp, BM, C::kBK));
    if (PASSES == 3) ISC_TRY(make_map(&m.a_lo, A.lo, M, C_in, A.ldp, BM, C::kBK));
    else m.a_lo = m.a_hi;
  }
  ISC_TRY(make_map(&m.b_hi, W.hi, N, 9LL * C_in, W.ldp, BN, C::kBK));
  if (PASSES == 3) ISC_TRY(make_map(&m.b_lo, W.lo, N, 9LL * C_in, W.ldp, BN, C::kBK));
  else m.b_lo = m.b_hi;
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = e.bias;
  ep.rows_per_group = 1;
  ep.act = e.act;
  ep.c = Cd.f32;
  ep.ldc = Cd.ld;
  ep.hi = Cd.hi;
  ep.lo = Cd.lo;
  ep.ldp = Cd.ldp;
  ep.M = M;
  ep.N = N;
  ep.K = 9 * C_in;
  ep.seg_kb = C_in / C::kBK;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) ep.seg_off[(dy + 1) * 3 + (dx + 1)] = dy * 16 + dx;
  ep.zero_border = zero_border;
  const int grid = persistent_grid<BN, 1>(M, N);
  ProfScope ps(ISC_K_GEMM_TC, 2.0 * M * N * 9.0 * C_in * PASSES, stream);
  if (e.act == ACT_RELU) return launch_kernel<PASSES, BN, ACT_RELU, EPI_STD, 1, AF32>(m, ep, grid, stream);
  return launch_kernel<PASSES, BN, ACT_NONE, EPI_STD, 1, AF32>(m, ep, grid, stream);
}

// Number of 256-column tiles per 128-row block for the mixed-width LSTM tile list (lstm_tile): the static round-robin
// walk is simulated for every choice and the cheapest one kept (cost of a wide tile 2, of a narrow one 1.15 — it
// moves more operand bytes per flop — plus 0.25 per tile for fill and epilogue).
static int lstm_wide_tiles(int M, int cg) {
  const int tiles_m = (M + cg * BM - 1) / (cg * BM), slots = num_sms() / cg, groups = (4 * H) / 256;
  int best = 0;
  double best_cost = 1e30;
  for (int w = groups; w >= 0; --w) {
    const int n_wide = tiles_m * w, total = n_wide + tiles_m * (groups - w) * 2;
    double worst = 0;
    for (int c = 0; c < slots && c < total; ++c) {
      double cost = 0;
      for (int t = c; t < total; t += slots) cost += (t < n_wide ? 2.0 : 1.15) + 0.25;
      worst = cost > worst ? cost : worst;
    }
    if (worst < best_cost - 1e-9) {
      best_cost = worst;
      best = w;
    }
  }
  return best;
}

template <int PASSES, int BN, int CG>
static int launch_lstm(const Operand& A, const Operand& W, int M, int K, const float* bias, const float* rowadd,
                       int64_t ld_rowadd, int rows_per_group, const LstmEpilogue& lstm, cudaStream_t stream) {
  using C = Cfg<PASSES, BN, CG>;
  const int N = 4 * H;
  // gate-interleaved weight rows: one 128-row box per 128 columns; a CTA pair loads half of a tile's rows per CTA in
  // 64-row boxes
  const int b_box = CG == 2 ? 64 : 128;
  Maps m;
  ISC_TRY(make_map(&m.a_hi, A.hi, M, K, A.ldp, BM, C::kBK));
  ISC_TRY(make_map(&m.b_hi, W.hi, N, K, W.ldp, b_box, C::kBK));
  if (PASSES == 3) {
    ISC_TRY(make_map(&m.a_lo, A.lo, M, K, A.ldp, BM, C::kBK));
    ISC_TRY(make_map(&m.b_lo, W.lo, N, K, W.ldp, b_box, C::kBK));
  } else {
    m.a_lo = m.a_hi;
    m.b_lo = m.b_hi;
  }
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = bias;
  ep.rowadd = rowadd;
  ep.ld_rowadd = ld_rowadd;
  ep.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  ep.M = M;
  ep.N = N;
  ep.K = K;
  ep.lstm = lstm;
  int grid = persistent_grid<BN, CG>(M, N);
  if (BN == 256) {
    ep.lstm_wide = lstm_wide_tiles(M, CG);
    // experiment hook: ISC_LSTM_WIDE1 / ISC_LSTM_WIDE2 force the number of 256-column tiles per row block for the
    // K <= 1024 (attention LSTM) / K > 1024 (language LSTM) gate GEMM
    static const int w1 = getenv("ISC_LSTM_WIDE1") ? atoi(getenv("ISC_LSTM_WIDE1")) : -1;
    static const int w2 = getenv("ISC_LSTM_WIDE2") ? atoi(getenv("ISC_LSTM_WIDE2")) : -1;
    const int wf = K <= 1024 ? w1 : w2;
    if (wf >= 0 && wf <= N / 256) ep.lstm_wide = wf;
    const int tiles = ((M + CG * BM - 1) / (CG * BM)) * (ep.lstm_wide + (N / 256 - ep.lstm_wide) * 2);
    const int slots = num_sms() / CG;
    grid = CG * (tiles < slots ? tiles : slots);
  }
  ProfScope ps(ISC_K_GEMM_TC, 2.0 * M * N * K * PASSES, stream);
  return launch_kernel<PASSES, BN, ACT_NONE, EPI_LSTM, CG>(m, ep, grid, stream);
}

// CTA pairs pay off on large GEMMs only (measured: 8192^3 1.41 -> 1.57 PFLOP/s x3-equivalent, logits-shaped 85 -> 75 us,
// no change on the M = 3072 step GEMMs, which are bound by wave quantisation and fill/drain, not by operand traffic).
static bool pair_pays(int M, int N, bool wide) {
  const long long pair_tiles = (long long)((M + 2 * BM - 1) / (2 * BM)) * ((N + (wide ? 255 : 127)) / (wide ? 256 : 128));
  return pair_tiles >= 3LL * (num_sms() / 2);
}
// ISC_GEMM_PAIR=0 forces single-CTA tiles, =1 forces CTA pairs wherever M spans more than one 128-row tile
static int pair_mode() {
  static int mode = -2;
  if (mode == -2) {
    const char* e = getenv("ISC_GEMM_PAIR");
    mode = e ? atoi(e) : -1;
  }
  return mode;
}

// 128x256 tiles move 25 % fewer operand bytes per flop (a tile costs ~1.5x a 128x128 one for 2x the work) but
// quantise worse on 148 SMs: pick the shape with the smaller waves x tile-cost estimate.
static bool use_wide_tiles(int M, int N) {
  if (N < 256) return false;
  static const int force = getenv("ISC_GEMM_WIDE") ? atoi(getenv("ISC_GEMM_WIDE")) : -1;  // experiments: 0 narrow, 1 wide
  if (force >= 0 && M == 3072) return force != 0;
  const long long sms = num_sms(), tm = (M + BM - 1) / BM;
  const long long narrow = (tm * ((N + 127) / 128) + sms - 1) / sms * 2;
  const long long wide = (tm * ((N + 255) / 256) + sms - 1) / sms * 3;
  return wide < narrow;
}

}  // namespace tc

#ifdef ISC_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int isc_debug_gemm_trace(void* buf) {
  unsigned long long* p = static_cast<unsigned long long*>(buf);
  return (int)cudaMemcpyToSymbol(tc::g_trace, &p, sizeof(p));
}
#endif

int gemm_tc(const Operand& A, const Operand& W, const Dest& C, int M, int N, int K, int passes, const Epilogue& ep,
            cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  ISC_REQUIRE(K > 0 && K % 8 == 0, "gemm_tc: K=%d must be a positive multiple of 8", K);
  ISC_REQUIRE(A.hi && W.hi, "gemm_tc: bf16 hi planes missing");
  const bool wide = tc::use_wide_tiles(M, N);
  const bool pair = M > tc::BM && (tc::pair_mode() == 1 || (tc::pair_mode() < 0 && tc::pair_pays(M, N, wide)));
  if (passes == 3) {
    ISC_REQUIRE(A.lo && W.lo, "gemm_tc: bf16 lo planes missing for the 3-pass mode");
    if (ep.act == ACT_NONE && !C.h16.out && tc::splitk_slices(M, N, K) > 1) return tc::launch<3, 128, 1>(A, W, C, M, N, K, ep, stream);
    if (pair) return wide ? tc::launch<3, 256, 2>(A, W, C, M, N, K, ep, stream) : tc::launch<3, 128, 2>(A, W, C, M, N, K, ep, stream);
    return wide ? tc::launch<3, 256, 1>(A, W, C, M, N, K, ep, stream) : tc::launch<3, 128, 1>(A, W, C, M, N, K, ep, stream);
  }
  if (pair) return wide ? tc::launch<1, 256, 2>(A, W, C, M, N, K, ep, stream) : tc::launch<1, 128, 2>(A, W, C, M, N, K, ep, stream);
  return wide ? tc::launch<1, 256, 1>(A, W, C, M, N, K, ep, stream) : tc::launch<1, 128, 1>(A, W, C, M, N, K, ep, stream);
}

int gemm_tc_gate(const Operand& A, const Operand& W, int M, int K, int passes, const float* bias, const float* addmat,
                 int64_t ld_addmat, const GateEpilogue& gate, const Dest& out, cudaStream_t stream) {
  if (M <= 0) return 0;
  ISC_REQUIRE(K > 0 && K % 8 == 0, "gemm_tc_gate: K=%d must be a positive multiple of 8", K);
  ISC_REQUIRE(A.hi && W.hi && (passes != 3 || (A.lo && W.lo)), "gemm_tc_gate: operand planes missing");
  ISC_REQUIRE(gate.alpha && gate.alpha_b && gate.cs, "gemm_tc_gate: gate parameters missing");
  auto aligned = [](const void* p, int64_t ld, int elem) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * elem) % 16 == 0; };
  ISC_REQUIRE(aligned(gate.cs, gate.ld_cs, 4) && aligned(gate.alpha, 4, 4) && (!bias || aligned(bias, 4, 4)) &&
                  (!addmat || aligned(addmat, ld_addmat, 4)) && (!out.f32 || aligned(out.f32, out.ld, 4)) &&
                  (!out.hi || ((reinterpret_cast<uintptr_t>(out.hi) & 7) == 0 && (out.ldp & 3) == 0)) &&
                  (!out.lo || (reinterpret_cast<uintptr_t>(out.lo) & 7) == 0),
              "gemm_tc_gate: operands must be 16-byte aligned with 16-byte multiple row pitches");
  tc::Maps m;
  tc::EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = bias;
  ep.addmat = addmat;
  ep.ld_addmat = ld_addmat;
  ep.rows_per_group = 1;
  ep.c = out.f32;
  ep.ldc = out.ld;
  ep.hi = out.hi;
  ep.lo = out.lo;
  ep.ldp = out.ldp;
  ep.M = M;
  ep.N = H;
  ep.K = K;
  ep.gate = gate;
  const int grid = ((M + tc::BM - 1) / tc::BM) * (H / 128);  // one 128 x 128 tile per CTA, clusters of 4 along a row block
  ProfScope ps(ISC_K_GEMM_TC, 2.0 * M * H * K * passes, stream);
  if (passes == 3) {
    ISC_TRY((tc::make_maps<3, 128, 1>(m, A, W, M, H, K)));
    return tc::launch_kernel<3, 128, ACT_TANH, tc::EPI_GATE, 1>(m, ep, grid, stream);
  }
  ISC_TRY((tc::make_maps<1, 128, 1>(m, A, W, M, H, K)));
  return tc::launch_kernel<1, 128, ACT_TANH, tc::EPI_GATE, 1>(m, ep, grid, stream);
}

int gemm_tc_af32(const float* A, int64_t lda, const Operand& W, const Dest& C, int M, int N, int K, int passes,
                 const Epilogue& ep, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  ISC_REQUIRE(K > 0 && K % 8 == 0, "gemm_tc_af32: K=%d must be a positive multiple of 8", K);
  ISC_REQUIRE(A && W.hi && (ep.act == ACT_RELU || ep.act == ACT_NONE), "gemm_tc_af32: operands missing / unsupported activation");
  // CTA pairs (ISC_AF_PAIR=1): measured SLOWER than single-CTA tiles on the prologue's region-embedding GEMM (M = 18816,
  // K = 2048: 113 -> ~190 us; the peer converters' remote arrivals and the five 32-column stages cost more than the halved
  // weight traffic returns), so off by default; the tests pass with it forced on
  static const int pair_env = getenv("ISC_AF_PAIR") ? atoi(getenv("ISC_AF_PAIR")) : 0;
  const bool pair = M > tc::BM && pair_env == 1;
  if (passes == 3) {
    ISC_REQUIRE(W.lo, "gemm_tc_af32: bf16 lo plane missing for the 3-pass mode");
    return pair ? tc::launch_af32<3, 2>(A, lda, W, C, M, N, K, ep, stream) : tc::launch_af32<3, 1>(A, lda, W, C, M, N, K, ep, stream);
  }
  return pair ? tc::launch_af32<1, 2>(A, lda, W, C, M, N, K, ep, stream) : tc::launch_af32<1, 1>(A, lda, W, C, M, N, K, ep, stream);
}

int gemm_tc_conv3x3(const float* A32, int64_t lda, const Operand& A, const Operand& W, const Dest& C, int M, int N, int C_in,
                    int passes, const Epilogue& ep, int zero_border, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  ISC_REQUIRE(M % 256 == 0 && C_in % 64 == 0 && N % 8 == 0, "gemm_tc_conv3x3: M must be images * 256, C a multiple of 64");
  ISC_REQUIRE((A32 || A.hi) && W.hi && (ep.act == ACT_RELU || ep.act == ACT_NONE) && !ep.rowadd && !ep.addmat,
              "gemm_tc_conv3x3: operands missing / unsupported epilogue");
  if (passes == 3) {
    ISC_REQUIRE(W.lo && (A32 || A.lo), "gemm_tc_conv3x3: bf16 lo planes missing for the 3-pass mode");
    return A32 ? tc::launch_conv<3, 1>(A32, A, lda, W, C, M, N, C_in, ep, zero_border, stream)
               : tc::launch_conv<3, 0>(nullptr, A, 0, W, C, M, N, C_in, ep, zero_border, stream);
  }
  return A32 ? tc::launch_conv<1, 1>(A32, A, lda, W, C, M, N, C_in, ep, zero_border, stream)
             : tc::launch_conv<1, 0>(nullptr, A, 0, W, C, M, N, C_in, ep, zero_border, stream);
}

int gemm_tc_lstm(const Operand& A, const Operand& W, int M, int K, int passes, const float* bias, const float* rowadd,
                 int64_t ld_rowadd, int rows_per_group, const LstmEpilogue& lstm, cudaStream_t stream) {
  if (M <= 0) return 0;
  ISC_REQUIRE(K > 0 && K % 8 == 0, "gemm_tc_lstm: K=%d must be a positive multiple of 8", K);
  static const bool mixed = getenv("ISC_LSTM_UNIFORM_TILES") == nullptr;  // set to use the uniform 128-column tile list
  // CTA pairs (a third fewer operand bytes per flop) once M spans several 256-row blocks; ISC_LSTM_PAIR=0/1 forces
  static const int pair_env = getenv("ISC_LSTM_PAIR") ? atoi(getenv("ISC_LSTM_PAIR")) : -1;
  const bool pair = M > tc::BM && (pair_env == 1 || (pair_env < 0 && M >= 8 * tc::BM));
  ISC_REQUIRE(A.hi && W.hi && lstm.c_prev && lstm.h_out && lstm.c_out, "gemm_tc_lstm: planes / state buffers missing");
  ISC_REQUIRE(!lstm.x_hi || ((lstm.ldx % 8) == 0 && (lstm.x_col % 8) == 0 && (reinterpret_cast<uintptr_t>(lstm.x_hi) & 15) == 0),
              "gemm_tc_lstm: operand planes must be 16-byte aligned");
  ISC_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm_tc_lstm: bias must be 16-byte aligned");
  ISC_REQUIRE(!rowadd || ((reinterpret_cast<uintptr_t>(rowadd) & 15) == 0 && (ld_rowadd & 3) == 0),
              "gemm_tc_lstm: rowadd must be 16-byte aligned");
  if (passes == 3) {
    ISC_REQUIRE(A.lo && W.lo, "gemm_tc_lstm: bf16 lo planes missing for the 3-pass mode");
    if (!mixed) return tc::launch_lstm<3, 128, 1>(A, W, M, K, bias, rowadd, ld_rowadd, rows_per_group, lstm, stream);
    if (pair) return tc::launch_lstm<3, 256, 2>(A, W, M, K, bias, rowadd, ld_rowadd, rows_per_group, lstm, stream);
    return tc::launch_lstm<3, 256, 1>(A, W, M, K, bias, rowadd, ld_rowadd, rows_per_group, lstm, stream);
  }
  if (!mixed) return tc::launch_lstm<1, 128, 1>(A, W, M, K, bias, rowadd, ld_rowadd, rows_per_group, lstm, stream);
  if (pair) return tc::launch_lstm<1, 256, 2>(A, W, M, K, bias, rowadd, ld_rowadd, rows_per_group, lstm, stream);
  return tc::launch_lstm<1, 256, 1>(A, W, M, K, bias, rowadd, ld_rowadd, rows_per_group, lstm, stream);
}

int gemm_tc_logits(const Operand& A, const Operand& W, int M, int N, int K, int passes, const float* bias,
                   const LogitsSelect& sel, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  ISC_REQUIRE(K > 0 && K % 8 == 0, "gemm_tc_logits: K=%d must be a positive multiple of 8", K);
  ISC_REQUIRE(A.hi && W.hi && sel.rec && sel.np == logits_slices(N) && (sel.k_sel == 4 || sel.k_sel == 8),
              "gemm_tc_logits: planes / records missing");
  // CTA pairs: a 256x256 pair tile takes a third fewer operand bytes per SM (the step GEMMs are bound by the L2 -> SM
  // port); worth it once there are several rounds of pair tiles
  static const int lp_env = getenv("ISC_LOGITS_PAIR") ? atoi(getenv("ISC_LOGITS_PAIR")) : -1;
  const bool pair = M > tc::BM && (tc::pair_mode() == 1 || lp_env == 1 || (lp_env < 0 && tc::pair_mode() < 0 && tc::pair_pays(M, N, true)));
  if (passes == 3) {
    ISC_REQUIRE(A.lo && W.lo, "gemm_tc_logits: bf16 lo planes missing for the 3-pass mode");
    return pair ? tc::launch_logits<3, 2>(A, W, M, N, K, bias, sel, stream) : tc::launch_logits<3, 1>(A, W, M, N, K, bias, sel, stream);
  }
  return pair ? tc::launch_logits<1, 2>(A, W, M, N, K, bias, sel, stream) : tc::launch_logits<1, 1>(A, W, M, N, K, bias, sel, stream);
}

}  // namespace isc
