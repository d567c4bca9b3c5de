// Tensor-map (descriptor-based) TMA: host-side encoding of CUtensorMap objects and the device-side tile store.
//
// The tile kernels own one ROW per thread while the feature arrays in HBM want whole rows written as contiguous
// lines.  A tile store through a tensor map does that transposition in the copy engine: the threads write their
// rows into a shared-memory box in the map's swizzle pattern (16-byte chunk index XOR row bits: bank-conflict free
// for one-row-per-lane writers), ONE thread issues cp.async.bulk.tensor for the whole [rows x bytes] box, and the
// engine clips the box against the tensor's bounds (ragged last tile of a sample).
//
// libpilegnn links cudart statically and has no link-time dependency on libcuda: cuTensorMapEncodeTiled is looked
// up through cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>

namespace pile {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tmap_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// byte tensor [n2][n1][inner] (inner contiguous; stride1 / stride2 in bytes, multiples of 16), box [1][box1][inner]
// (or [box2][box1][inner])
// with the swizzle that matches `inner` (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B).  Returns 0 or a cudaError.
inline int tmap_encode_rows(CUtensorMap* out, void* base, uint32_t inner, uint64_t n1, uint64_t stride1, uint64_t n2,
                            uint64_t stride2, uint32_t box1, uint32_t box2 = 1) {
  EncodeTiledFn enc = tmap_encoder();
  if (enc == nullptr) return (int)cudaErrorNotSupported;
  const cuuint64_t dims[3] = {inner, n1, n2};
  const cuuint64_t strides[2] = {stride1, stride2};
  const cuuint32_t box[3] = {inner, box1, box2};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// small cache of encoded maps (encoding costs a few microseconds of host time per call)
template <int NMAPS>
struct TmapCache {
  struct Entry {
    const void* base = nullptr;
    long long k0 = -1, k1 = -1;
    CUtensorMap m[NMAPS];
  };
  static constexpr int CAP = 8;
  Entry e[CAP];
  int next = 0;
  std::mutex mu;
  // fill(maps) -> int status is called on a miss
  template <typename Fill>
  int get(const void* base, long long k0, long long k1, CUtensorMap (&out)[NMAPS], Fill fill) {
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < CAP; ++i)
      if (e[i].base == base && e[i].k0 == k0 && e[i].k1 == k1) {
        for (int j = 0; j < NMAPS; ++j) out[j] = e[i].m[j];
        return 0;
      }
    Entry& s = e[next];
    const int st = fill(s.m);
    if (st) { s.base = nullptr; return st; }
    s.base = base; s.k0 = k0; s.k1 = k1;
    next = (next + 1) % CAP;
    for (int j = 0; j < NMAPS; ++j) out[j] = s.m[j];
    return 0;
  }
};

#ifdef __CUDACC__
namespace tc {
// shared-memory box -> global tile at coordinates (c0 bytes, c1 rows, c2 planes); bulk async-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t smem_addr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_addr), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed (writes visible)
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
}  // namespace tc
#endif

}  // namespace pile
