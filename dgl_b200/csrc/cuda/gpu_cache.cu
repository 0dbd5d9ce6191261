// Set-associative GPU feature cache (sm_100a): 32-way sets, one warp per key, lane = way.
// Replaces the reference's HugeCTR gpu_cache wrapper (/root/reference/src/runtime/cuda/gpu_cache.cu:41-180,
// third_party/HugeCTR/gpu_cache/src/nv_gpu_cache.cu, graphbolt/src/cuda/extension/gpu_cache.cu).
// query : ballot over the set's 32 tags; on a hit the warp copies the row with 16-byte lanes and refreshes
//         the slot's timestamp.  replace: per-set spin lock, victim = empty way or the oldest timestamp.
#include "common.cuh"

namespace dglb {

__device__ __forceinline__ uint64_t cache_hash(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  return k;
}

__device__ __forceinline__ void warp_copy_row(const char* src, char* dst, int64_t row_bytes, int lane) {
  if ((row_bytes & 15) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    for (int64_t c = lane; c < row_bytes / 16; c += 32) reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
  } else {
    for (int64_t c = lane; c < row_bytes; c += 32) dst[c] = src[c];
  }
}

__global__ void __launch_bounds__(256) gpucache_query_kernel(const int64_t* __restrict__ tags, int32_t* __restrict__ stamp,
                                                             const char* __restrict__ values, int epoch,
                                                             int64_t num_sets, const int64_t* __restrict__ qkeys,
                                                             char* __restrict__ out, int64_t n,
                                                             int32_t* __restrict__ hit, int64_t row_bytes) {
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n;
       i += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t key = qkeys[i];
    const int64_t set = (int64_t)(cache_hash((uint64_t)key) % (uint64_t)num_sets);
    const int64_t slot = set * 32 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, tags[slot] == key);
    if (m) {
      const int way = __ffs(m) - 1;
      const int64_t s = set * 32 + way;
      warp_copy_row(values + s * row_bytes, out + i * row_bytes, row_bytes, lane);
      if (lane == 0) {
        stamp[s] = epoch;
        hit[i] = 1;
      }
    } else if (lane == 0) {
      hit[i] = 0;
    }
  }
}

__global__ void __launch_bounds__(256) gpucache_replace_kernel(int64_t* __restrict__ tags, int32_t* __restrict__ stamp,
                                                               char* __restrict__ values, int32_t* __restrict__ locks,
                                                               int epoch, int64_t num_sets,
                                                               const int64_t* __restrict__ keys,
                                                               const char* __restrict__ vals, int64_t n,
                                                               int64_t row_bytes) {
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n;
       i += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t key = keys[i];
    const int64_t set = (int64_t)(cache_hash((uint64_t)key) % (uint64_t)num_sets);
    if (lane == 0) {
      while (atomicCAS(&locks[set], 0, 1) != 0) { __nanosleep(64); }
    }
    __syncwarp();
    __threadfence();
    const int64_t slot = set * 32 + lane;
    const int64_t tag = *(volatile int64_t*)&tags[slot];
    const int32_t st = *(volatile int32_t*)&stamp[slot];
    unsigned m = __ballot_sync(0xffffffffu, tag == key);
    int way;
    if (m) {
      way = __ffs(m) - 1;
    } else {
      const unsigned e = __ballot_sync(0xffffffffu, tag == -1);
      if (e) {
        way = __ffs(e) - 1;
      } else {  // oldest timestamp: warp arg-min
        int best = st, bway = lane;
        for (int o = 16; o > 0; o >>= 1) {
          const int ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int ow = __shfl_xor_sync(0xffffffffu, bway, o);
          if (ob < best || (ob == best && ow < bway)) { best = ob; bway = ow; }
        }
        way = bway;
      }
    }
    const int64_t s = set * 32 + way;
    warp_copy_row(vals + i * row_bytes, values + s * row_bytes, row_bytes, lane);
    __syncwarp();
    if (lane == 0) {
      tags[s] = key;
      stamp[s] = epoch;
      __threadfence();
      atomicExch(&locks[set], 0);
    }
    __syncwarp();
  }
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_gpucache_query(const void* tags, void* stamp, const void* values, int epoch, int64_t num_sets,
                                    int ways, const void* qkeys, void* out, int64_t n, void* hit, int64_t row_bytes,
                                    void* stream) {
  DGLB_TRY
  DGLB_CHECK(ways == 32, "gpu cache is 32-way");
  if (n == 0) return 0;
  int64_t blocks = (n + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gpucache_query_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const int64_t*)tags, (int32_t*)stamp, (const char*)values, epoch, num_sets, (const int64_t*)qkeys, (char*)out, n, (int32_t*)hit, row_bytes);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_gpucache_replace(void* tags, void* stamp, void* values, void* locks, int epoch, int64_t num_sets,
                                      int ways, const void* keys, const void* vals, int64_t n, int64_t row_bytes,
                                      void* stream) {
  DGLB_TRY
  DGLB_CHECK(ways == 32, "gpu cache is 32-way");
  if (n == 0) return 0;
  int64_t blocks = (n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  gpucache_replace_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((int64_t*)tags, (int32_t*)stamp, (char*)values, (int32_t*)locks, epoch, num_sets, (const int64_t*)keys, (const char*)vals, n, row_bytes);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
