// Host-side CUtensorMap construction for the TMA kernels, with a process-wide cache.
//
// A tensor map is a pure function of (base pointer, dtype, rank, dims, strides, box, swizzle): the forward of one
// (b, H, W) shape re-issues the same few hundred maps every call (PyTorch's caching allocator hands the same
// activation addresses back), so every map is encoded through the driver once and found by hash afterwards — the
// per-(pointer, shape) plan cache SURVEY 8(b) asks for, without a handle the caller has to carry.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace pgt {

struct TmapKey {
  uint64_t base;
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
  uint32_t rank, dtype, swizzle, pad;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
    return (size_t)h;
  }
};

inline PFN_cuTensorMapEncodeTiled_v12000 tmap_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }();
  return fn;
}

// One cache for the whole library (defined in api.cu).
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>& tmap_cache();
std::mutex& tmap_cache_mutex();
void tmap_cache_stats(long long* hits, long long* misses);
void tmap_cache_count(bool hit);

enum { TMAP_SW128 = 0, TMAP_SW_NONE = 1 };

// dims / box in elements (innermost first), strides_bytes for dims 1..rank-1.  dtype: PGT_BF16 or PGT_F32.
inline int tmap_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int dtype = PGT_BF16, int swizzle = TMAP_SW128) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = reinterpret_cast<uint64_t>(base);
  key.rank = (uint32_t)rank; key.dtype = (uint32_t)dtype; key.swizzle = (uint32_t)swizzle;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i > 0) key.strides[i - 1] = strides_bytes[i - 1];
  }
  {
    std::lock_guard<std::mutex> g(tmap_cache_mutex());
    auto& c = tmap_cache();
    auto it = c.find(key);
    if (it != c.end()) { *map = it->second; tmap_cache_count(true); return PGT_OK; }
  }
  auto fn = tmap_encode_fn();
  if (fn == nullptr) return PGT_ERR_DRIVER;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, dtype == PGT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                  const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == TMAP_SW128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return PGT_ERR_DRIVER;
  {
    std::lock_guard<std::mutex> g(tmap_cache_mutex());
    auto& c = tmap_cache();
    if (c.size() > (1u << 16)) c.clear();          // bound the cache (a few hundred maps per (b, H, W) in practice)
    c.emplace(key, *map);
    tmap_cache_count(false);
  }
  return PGT_OK;
}

// bf16 row-major matrix [rows, cols] with row pitch `ld` elements, box = 64 columns (one 128-byte swizzled row) x box_rows.
inline int tmap_rows_bf16(CUtensorMap* map, const void* base, long long ld, long long rows, int cols, int box_rows) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)ld * 2};
  const uint32_t box[2] = {64, (uint32_t)box_rows};
  return tmap_encode(map, base, 2, dims, strides, box);
}

// Per-device one-time kernel attribute setup (cudaFuncSetAttribute is per device, the library is per process).
struct PerDeviceOnce {
  std::mutex m;
  uint64_t done = 0;                                // bit d: device d configured (64 devices are plenty)
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> g(m);
    if (dev < 64 && (done >> dev) & 1ull) return cudaSuccess;
    e = f();
    if (e == cudaSuccess && dev < 64) done |= 1ull << dev;
    return e;
  }
};

}  // namespace pgt
