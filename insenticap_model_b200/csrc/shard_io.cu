// Feature shards: the input side of the path. The reference keeps one HDF5 dataset per
// image and opens the file for every item (/root/reference/dataloader.py:164-204); at decode rates of 1e5 images/s that
// is the bottleneck. A shard is ONE flat file of fixed-size records — fc_feats [D] then att_feats [L][D] per image, fp32
// (bit-exact with the reference's arrays) or bf16 (what ISC_PREC_BF16 computes in; half the bytes) — that is mmap'ed
// once; a batch is gathered record by record into the caller's PINNED staging buffers by a few host threads, from where
// one cudaMemcpyAsync per tensor moves it (host-only code). Where the platform can page-lock a read-only file mapping,
// isc_shard_pin + isc_shard_copy_to_device skip the staging pass and DMA each record straight from the page cache.
//
// layout (little endian):  [0,64) header | names: n NUL-terminated strings | pad to 256 | records, stride = record_bytes
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/isc.h"
#include "common.cuh"

namespace {

struct ShardHeader {
  char magic[8];  // "ISCFEAT1"
  uint32_t version, dtype;  // dtype: ISC_SHARD_F32 / ISC_SHARD_BF16
  uint32_t feat_dim, n_regions;
  uint64_t n_images, names_bytes, data_offset, record_bytes;
  uint8_t pad[8];
};
static_assert(sizeof(ShardHeader) == 64, "shard header is 64 bytes");

struct Shard {
  int fd = -1;
  const uint8_t* map = nullptr;
  size_t map_bytes = 0;
  ShardHeader h;
  bool pinned = false;   // mapping registered in place
  uint8_t* copy = nullptr;  // or: the whole file read into page-locked memory (then map points here)
  const uint8_t* dev_view = nullptr;  // device-side address of the page-locked records (zero-copy reads by the gather kernel)
  std::vector<const char*> names;
  std::unordered_map<std::string, int64_t> index;
};

size_t elem_bytes(uint32_t dtype) { return dtype == ISC_SHARD_F32 ? 4 : 2; }
bool known_dtype(uint32_t dtype) { return dtype == ISC_SHARD_F32 || dtype == ISC_SHARD_BF16 || dtype == ISC_SHARD_F16; }
uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Page-locked records -> device tensors in ONE launch: a small persistent grid walks (item, 32 KB chunk) work units and
// reads the records straight from host memory over PCIe (zero-copy), four 16-byte loads in flight per thread. One
// cudaMemcpyAsync per record instead costs the issuing thread the whole DMA time (512 calls per batch, throttled by the
// copy queue: 8.8 ms of a 12.2 ms batch period in profiles/loader_pipeline.py), which serialised the loader's Python with
// its copies.
constexpr int kGatherThreads = 512;
constexpr int kGatherCtas = 32;
__global__ void __launch_bounds__(kGatherThreads) shard_gather_kernel(const uint8_t* __restrict__ records, uint64_t record_bytes,
                                                                      const int64_t* __restrict__ idx, int64_t n, uint32_t src_off,
                                                                      uint32_t item_bytes, uint8_t* __restrict__ dst) {
  constexpr uint32_t kChunk = kGatherThreads * 16 * 4;
  const uint32_t chunks = (item_bytes + kChunk - 1) / kChunk;
  const int64_t units = n * chunks;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int64_t item = u / chunks;
    const uint32_t c0 = (uint32_t)(u - item * chunks) * kChunk;
    const uint8_t* src = records + (uint64_t)idx[item] * record_bytes + src_off;
    uint8_t* out = dst + (uint64_t)item * item_bytes;
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t off = c0 + (k * kGatherThreads + threadIdx.x) * 16;
      if (off < item_bytes) v[k] = __ldcs(reinterpret_cast<const uint4*>(src + off));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t off = c0 + (k * kGatherThreads + threadIdx.x) * 16;
      if (off < item_bytes) __stcs(reinterpret_cast<uint4*>(out + off), v[k]);
    }
  }
}

uint16_t f32_to_bf16_rne(float f) {  // same rounding as __float2bfloat16_rn (NaN kept quiet)
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

void close_shard(Shard* s) {
  if (s->map && s->pinned) cudaHostUnregister(const_cast<uint8_t*>(s->map));
  if (s->copy) {
    cudaFreeHost(s->copy);
    s->map = nullptr;
  }
  if (s->map) munmap(const_cast<uint8_t*>(s->map), s->map_bytes);
  if (s->fd >= 0) close(s->fd);
  delete s;
}

}  // namespace

extern "C" {

int isc_shard_write(const char* path, int dtype, int feat_dim, int n_regions, int64_t n_images, const char* const* names,
                    const float* fc_feats, const float* att_feats) {
  ISC_REQUIRE(path && names && fc_feats && att_feats, "shard_write: NULL argument");
  ISC_REQUIRE(known_dtype((uint32_t)dtype) && feat_dim > 0 && n_regions > 0 && n_images > 0,
              "shard_write: bad dtype / dims");
  ShardHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "ISCFEAT1", 8);
  h.version = 1;
  h.dtype = (uint32_t)dtype;
  h.feat_dim = (uint32_t)feat_dim;
  h.n_regions = (uint32_t)n_regions;
  h.n_images = (uint64_t)n_images;
  for (int64_t i = 0; i < n_images; ++i) {
    ISC_REQUIRE(names[i] && names[i][0], "shard_write: empty name at %lld", (long long)i);
    h.names_bytes += strlen(names[i]) + 1;
  }
  const size_t eb = elem_bytes(h.dtype);
  const uint64_t rec_elems = (uint64_t)feat_dim * (1 + (uint64_t)n_regions);
  h.record_bytes = round_up(rec_elems * eb, 256);
  h.data_offset = round_up(sizeof(h) + h.names_bytes, 256);
  FILE* f = fopen(path, "wb");
  if (!f) {
    isc::set_error("shard_write: cannot create %s: %s", path, strerror(errno));
    return ISC_ERR_ARG;
  }
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  for (int64_t i = 0; ok && i < n_images; ++i) ok = fwrite(names[i], strlen(names[i]) + 1, 1, f) == 1;
  std::vector<uint8_t> rec(h.record_bytes, 0);
  std::vector<uint8_t> zeros(256, 0);
  const uint64_t gap = h.data_offset - (sizeof(h) + h.names_bytes);
  if (ok && gap) ok = fwrite(zeros.data(), gap, 1, f) == 1;
  for (int64_t i = 0; ok && i < n_images; ++i) {
    const float* fc = fc_feats + (uint64_t)i * feat_dim;
    const float* att = att_feats + (uint64_t)i * n_regions * feat_dim;
    if (h.dtype == ISC_SHARD_F32) {
      memcpy(rec.data(), fc, (size_t)feat_dim * 4);
      memcpy(rec.data() + (size_t)feat_dim * 4, att, (size_t)n_regions * feat_dim * 4);
    } else if (h.dtype == ISC_SHARD_BF16) {
      uint16_t* o = reinterpret_cast<uint16_t*>(rec.data());
      for (int j = 0; j < feat_dim; ++j) o[j] = f32_to_bf16_rne(fc[j]);
      for (uint64_t j = 0; j < (uint64_t)n_regions * feat_dim; ++j) o[feat_dim + j] = f32_to_bf16_rne(att[j]);
    } else {  // IEEE half, round to nearest even (cuda_fp16.h's host conversion)
      __half* o = reinterpret_cast<__half*>(rec.data());
      for (int j = 0; j < feat_dim; ++j) o[j] = __float2half_rn(fc[j]);
      for (uint64_t j = 0; j < (uint64_t)n_regions * feat_dim; ++j) o[feat_dim + j] = __float2half_rn(att[j]);
    }
    ok = fwrite(rec.data(), h.record_bytes, 1, f) == 1;
  }
  if (fclose(f) != 0) ok = false;
  if (!ok) {
    isc::set_error("shard_write: short write to %s: %s", path, strerror(errno));
    return ISC_ERR_ARG;
  }
  return 0;
}

int isc_shard_open(const char* path, isc_shard_t* out) {
  ISC_REQUIRE(path && out, "shard_open: NULL argument");
  *out = nullptr;
  Shard* s = new Shard();
  s->fd = open(path, O_RDONLY);
  if (s->fd < 0) {
    isc::set_error("shard_open: cannot open %s: %s", path, strerror(errno));
    close_shard(s);
    return ISC_ERR_ARG;
  }
  struct stat st;
  if (fstat(s->fd, &st) != 0 || (size_t)st.st_size < sizeof(ShardHeader)) {
    isc::set_error("shard_open: %s is shorter than a shard header", path);
    close_shard(s);
    return ISC_ERR_ARG;
  }
  s->map_bytes = (size_t)st.st_size;
  void* m = mmap(nullptr, s->map_bytes, PROT_READ, MAP_SHARED, s->fd, 0);
  if (m == MAP_FAILED) {
    s->map = nullptr;
    isc::set_error("shard_open: mmap of %s failed: %s", path, strerror(errno));
    close_shard(s);
    return ISC_ERR_ARG;
  }
  s->map = static_cast<const uint8_t*>(m);
  memcpy(&s->h, s->map, sizeof(ShardHeader));
  const ShardHeader& h = s->h;
  const bool sane = memcmp(h.magic, "ISCFEAT1", 8) == 0 && h.version == 1 &&
                    known_dtype(h.dtype) && h.feat_dim > 0 && h.n_regions > 0 && h.feat_dim <= (1u << 24) && h.n_regions <= (1u << 24) &&
                    h.record_bytes >= (uint64_t)h.feat_dim * (1 + (uint64_t)h.n_regions) * elem_bytes(h.dtype) &&
                    // every bound is checked without a product or sum that could wrap for a crafted header
                    h.record_bytes > 0 && h.names_bytes <= s->map_bytes - sizeof(ShardHeader) &&
                    h.data_offset >= sizeof(ShardHeader) + h.names_bytes && h.data_offset <= s->map_bytes &&
                    h.n_images <= (s->map_bytes - h.data_offset) / h.record_bytes;
  if (!sane) {
    isc::set_error("shard_open: %s is not a version-1 ISCFEAT1 shard (bad magic, header or truncated file)", path);
    close_shard(s);
    return ISC_ERR_ARG;
  }
  const char* p = reinterpret_cast<const char*>(s->map + sizeof(ShardHeader));
  const char* end = p + h.names_bytes;
  s->names.reserve(h.n_images);
  s->index.reserve(h.n_images * 2);
  for (uint64_t i = 0; i < h.n_images; ++i) {
    const void* nul = p < end ? memchr(p, 0, (size_t)(end - p)) : nullptr;
    if (!nul) {
      isc::set_error("shard_open: name table of %s ends after %llu of %llu names", path, (unsigned long long)i,
                     (unsigned long long)h.n_images);
      close_shard(s);
      return ISC_ERR_ARG;
    }
    s->names.push_back(p);
    s->index.emplace(std::string(p), (int64_t)i);
    p = static_cast<const char*>(nul) + 1;
  }
  *out = s;
  return 0;
}

int isc_shard_close(isc_shard_t shard) {
  if (shard) close_shard(static_cast<Shard*>(shard));
  return 0;
}

int isc_shard_info(isc_shard_t shard, int64_t* n_images, int* feat_dim, int* n_regions, int* dtype) {
  ISC_REQUIRE(shard, "shard_info: NULL shard");
  const Shard* s = static_cast<const Shard*>(shard);
  if (n_images) *n_images = (int64_t)s->h.n_images;
  if (feat_dim) *feat_dim = (int)s->h.feat_dim;
  if (n_regions) *n_regions = (int)s->h.n_regions;
  if (dtype) *dtype = (int)s->h.dtype;
  return 0;
}

int64_t isc_shard_find(isc_shard_t shard, const char* name) {
  if (!shard || !name) return -1;
  const Shard* s = static_cast<const Shard*>(shard);
  auto it = s->index.find(name);
  return it == s->index.end() ? -1 : it->second;
}

const char* isc_shard_name(isc_shard_t shard, int64_t index) {
  const Shard* s = static_cast<const Shard*>(shard);
  if (!s || index < 0 || (uint64_t)index >= s->h.n_images) return nullptr;
  return s->names[(size_t)index];
}

int isc_shard_gather(isc_shard_t shard, const int64_t* indices, int64_t n, void* fc_dst, void* att_dst, int n_threads) {
  ISC_REQUIRE(shard && indices && n >= 0 && (fc_dst || att_dst), "shard_gather: NULL argument");
  const Shard* s = static_cast<const Shard*>(shard);
  const ShardHeader& h = s->h;
  for (int64_t i = 0; i < n; ++i)
    ISC_REQUIRE(indices[i] >= 0 && (uint64_t)indices[i] < h.n_images, "shard_gather: index %lld at position %lld out of range [0, %llu)",
                (long long)indices[i], (long long)i, (unsigned long long)h.n_images);
  const size_t fc_bytes = (size_t)h.feat_dim * elem_bytes(h.dtype);
  const size_t att_bytes = fc_bytes * h.n_regions;
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const uint8_t* rec = s->map + h.data_offset + (uint64_t)indices[i] * h.record_bytes;
      if (fc_dst) memcpy(static_cast<uint8_t*>(fc_dst) + (size_t)i * fc_bytes, rec, fc_bytes);
      if (att_dst) memcpy(static_cast<uint8_t*>(att_dst) + (size_t)i * att_bytes, rec + fc_bytes, att_bytes);
    }
  };
  int nt = n_threads < 1 ? 1 : n_threads;
  if (nt > n) nt = (int)(n > 0 ? n : 1);
  if (nt == 1) {
    work(0, n);
    return 0;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (int t = 0; t < nt; ++t) pool.emplace_back(work, n * t / nt, n * (t + 1) / nt);
  for (auto& th : pool) th.join();
  return 0;
}

// Zero-staging path: page-lock the mapping once, then DMA records straight out of the page cache.
int isc_shard_pin(isc_shard_t shard) {
  ISC_REQUIRE(shard, "shard_pin: NULL shard");
  Shard* s = static_cast<Shard*>(shard);
  if (s->pinned || s->copy) return 0;
  if (cudaHostRegister(const_cast<uint8_t*>(s->map), s->map_bytes,
                       cudaHostRegisterReadOnly | cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) {
    s->pinned = true;
    void* dv = nullptr;
    if (cudaHostGetDevicePointer(&dv, const_cast<uint8_t*>(s->map), 0) == cudaSuccess) s->dev_view = static_cast<const uint8_t*>(dv);
    else cudaGetLastError();
    return 0;
  }
  cudaGetLastError();  // not every platform can page-lock a read-only file mapping: keep the shard in page-locked RAM
  uint8_t* buf = nullptr;
  ISC_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&buf), s->map_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
  const int nt = 8;
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([=] {
      const size_t lo = s->map_bytes * t / nt, hi = s->map_bytes * (t + 1) / nt;
      memcpy(buf + lo, s->map + lo, hi - lo);
    });
  for (auto& th : pool) th.join();
  const ptrdiff_t shift = buf - s->map;
  for (auto& nm : s->names) nm += shift;
  munmap(const_cast<uint8_t*>(s->map), s->map_bytes);
  s->map = s->copy = buf;
  void* dv = nullptr;
  if (cudaHostGetDevicePointer(&dv, buf, 0) == cudaSuccess) s->dev_view = static_cast<const uint8_t*>(dv);
  else cudaGetLastError();
  return 0;
}

int isc_shard_copy_to_device(isc_shard_t shard, const int64_t* indices, int64_t n, void* fc_dst, void* att_dst,
                             isc_stream_t stream) {
  ISC_REQUIRE(shard && indices && n >= 0 && (fc_dst || att_dst), "shard_copy_to_device: NULL argument");
  const Shard* s = static_cast<const Shard*>(shard);
  const ShardHeader& h = s->h;
  for (int64_t i = 0; i < n; ++i)
    ISC_REQUIRE(indices[i] >= 0 && (uint64_t)indices[i] < h.n_images,
                "shard_copy_to_device: index %lld at position %lld out of range [0, %llu)", (long long)indices[i], (long long)i,
                (unsigned long long)h.n_images);
  const size_t fc_bytes = (size_t)h.feat_dim * elem_bytes(h.dtype);
  const size_t att_bytes = fc_bytes * h.n_regions;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ISC_SHARD_COPY=kernel: the zero-copy gather kernel above instead of one cudaMemcpyAsync per record. Measured
  // (profiles/shard_copy_bench.py, 512 fp16 records = 0.41 GB): the per-record DMA is steady (11.6 ms every time) but
  // throttles the ISSUING thread for the whole DMA time; the kernel returns at once and has the same median rate
  // (12.6 ms) but outliers of tens to hundreds of ms on this pool — so the DMA loop stays the default and the loader
  // issues it from a thread of its own (dataloader.DevicePrefetcher).
  static const int mode = [] {
    const char* e = getenv("ISC_SHARD_COPY");
    return (e && strcmp(e, "kernel") == 0) ? 2 : 1;
  }();
  if (n > 0 && mode == 2 && s->dev_view && fc_bytes % 16 == 0 && h.data_offset % 16 == 0 && h.record_bytes % 16 == 0 &&
      att_bytes <= 0xffffffffull) {
    // the indices travel through a stream-ordered allocation (the caller's array is pageable: its copy is staged before
    // cudaMemcpyAsync returns), the records are read in place by the gather kernel
    int64_t* d_idx = nullptr;
    ISC_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_idx), (size_t)n * sizeof(int64_t), st));
    ISC_CUDA(cudaMemcpyAsync(d_idx, indices, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    const uint8_t* recs = s->dev_view + h.data_offset;
    if (fc_dst)
      shard_gather_kernel<<<kGatherCtas, kGatherThreads, 0, st>>>(recs, h.record_bytes, d_idx, n, 0, (uint32_t)fc_bytes,
                                                                 static_cast<uint8_t*>(fc_dst));
    if (att_dst)
      shard_gather_kernel<<<kGatherCtas, kGatherThreads, 0, st>>>(recs, h.record_bytes, d_idx, n, (uint32_t)fc_bytes,
                                                                 (uint32_t)att_bytes, static_cast<uint8_t*>(att_dst));
    ISC_LAUNCH_CHECK();
    ISC_CUDA(cudaFreeAsync(d_idx, st));
    return 0;
  }
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* rec = s->map + h.data_offset + (uint64_t)indices[i] * h.record_bytes;
    if (fc_dst)
      ISC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(fc_dst) + (size_t)i * fc_bytes, rec, fc_bytes, cudaMemcpyHostToDevice, st));
    if (att_dst)
      ISC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(att_dst) + (size_t)i * att_bytes, rec + fc_bytes, att_bytes,
                               cudaMemcpyHostToDevice, st));
  }
  return 0;
}

}  // extern "C"
