// GPU cache of hot vertices' adjacency lists + compact-CSC gather of arbitrary seeds (device or host-pinned graph).
//
// Reference: graphbolt/src/cuda/extension/gpu_graph_cache.cu:54-546 (cuco::static_map keyed by vertex, `_Insert`,
// `_QueryAndIncrement`, frequency threshold) and graphbolt/src/cuda/index_select_csc_impl.cu:61-330
// (`_CopyIndicesAlignedKernel`, UVA or device). Here:
//   * the map is an own open-addressing table (linear probing, 64-bit keys, CAS insert) living in plain device
//     tensors: key | access count | offset | length, so the whole cache state is checkpointable with torch.save;
//   * `query` looks every seed up, bumps its access counter (inserting unseen keys while the table has room) and
//     reports hit / miss plus the cached (offset, length);
//   * `replace` admits missing seeds whose counter reached the threshold: space in the edge pool is reserved with one
//     atomicAdd per vertex, the list is copied by the same kernel, (offset, length) are published last;
//   * `csc_gather` copies the adjacency lists of a seed list into a compact CSC: one sub-warp per seed, 16-byte copies
//     when both sides are aligned; the source may be host-pinned memory (UVA) - this is K29 and the miss path of the cache.
// No host synchronisation inside any call.
#include "common.cuh"

namespace dglb {

constexpr int64_t kEmptyKey = -1;

__device__ __forceinline__ uint64_t gc_hash(int64_t k) { return mix64((uint64_t)k * 0x9E3779B97F4A7C15ull); }

// returns the slot of `key` (inserting it if absent and there is a free slot in the probe window), or -1
__device__ __forceinline__ int64_t gc_find_or_insert(int64_t* keys, int64_t slots, int64_t key, bool insert) {
  uint64_t h = gc_hash(key) % (uint64_t)slots;
  for (int probe = 0; probe < 128; ++probe) {
    const int64_t cur = keys[h];
    if (cur == key) return (int64_t)h;
    if (cur == kEmptyKey) {
      if (!insert) return -1;
      const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + h),
                                               (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == (unsigned long long)kEmptyKey || old == (unsigned long long)key) return (int64_t)h;
    }
    h = (h + 1) % (uint64_t)slots;
  }
  return -1;
}

// hit[i] = 1 and (off[i], len[i]) valid when seed i is cached; every seed's access counter is incremented.
template <typename I>
__global__ void gc_query_kernel(int64_t* __restrict__ keys, int32_t* __restrict__ counts, const int64_t* __restrict__ offs,
                                const int32_t* __restrict__ lens, int64_t slots, const I* __restrict__ seeds, int64_t n,
                                uint8_t* __restrict__ hit, int64_t* __restrict__ off_out, int32_t* __restrict__ len_out,
                                int64_t* __restrict__ stats) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t s = gc_find_or_insert(keys, slots, (int64_t)seeds[i], true);
  bool h = false;
  if (s >= 0) {
    atomicAdd(counts + s, 1);
    const int32_t l = lens[s];
    if (l >= 0) {
      h = true;
      off_out[i] = offs[s];
      len_out[i] = l;
    }
  }
  hit[i] = h ? 1 : 0;
  if (!h) {
    len_out[i] = 0;
    off_out[i] = 0;
    atomicAdd(reinterpret_cast<unsigned long long*>(stats + 1), 1ull);
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats), (unsigned long long)n);
}

// one sub-warp (8 lanes) per seed: admit + copy. indptr/indices describe the freshly fetched compact CSC of `seeds`.
template <typename I>
__global__ void gc_replace_kernel(int64_t* __restrict__ keys, const int32_t* __restrict__ counts, int64_t* __restrict__ offs,
                                  int32_t* __restrict__ lens, int64_t slots, int threshold, const I* __restrict__ seeds,
                                  int64_t n, const int64_t* __restrict__ indptr, const I* __restrict__ indices,
                                  I* __restrict__ pool, int64_t pool_cap, int64_t* __restrict__ used) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 7;
  if (i >= n) return;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
  int64_t dst = -1, s = -1;
  const int64_t lo = indptr[i], len = indptr[i + 1] - lo;
  if (lane == 0) {
    s = gc_find_or_insert(keys, slots, (int64_t)seeds[i], false);
    // lens[s] == -1: known but not cached; claim it with CAS(-1 -> -2) so duplicates of a seed admit it once
    if (s >= 0 && counts[s] >= threshold && atomicCAS(lens + s, -1, -2) == -1) {
      const int64_t at = (int64_t)atomicAdd(reinterpret_cast<unsigned long long*>(used), (unsigned long long)len);
      if (at + len <= pool_cap) dst = at;
      else {
        atomicAdd(reinterpret_cast<unsigned long long*>(used), (unsigned long long)(-len));  // give the space back
        atomicExch(lens + s, -1);
      }
    }
  }
  dst = __shfl_sync(gmask, dst, (threadIdx.x & 31) & ~7);
  if (dst < 0) return;
  for (int64_t q = lane; q < len; q += 8) pool[dst + q] = indices[lo + q];
  __threadfence();
  __syncwarp(gmask);
  if (lane == 0) {
    offs[s] = dst;
    atomicExch(lens + s, (int32_t)len);  // publish last
  }
}

// degrees of the seeds in a CSC (indptr may be host-pinned); the caller scans them into out_indptr.
template <typename I>
__global__ void csc_degree_kernel(const I* __restrict__ indptr, const I* __restrict__ seeds, int64_t n,
                                  int64_t* __restrict__ deg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t s = seeds[i];
    deg[i] = (int64_t)indptr[s + 1] - (int64_t)indptr[s];
  }
}

// out_indices[out_indptr[i] .. out_indptr[i+1]) = src[start_i .. start_i + len): start from `indptr[seed]` (graph gather)
// or from an explicit offset list (cache pool gather). One sub-warp of 8 lanes per seed; optional second array (edge ids
// or any per-edge attribute with the same layout).
template <typename I, typename S>
__global__ void seg_copy_kernel(const S* __restrict__ starts, const I* __restrict__ seeds, int64_t n,
                                const int64_t* __restrict__ out_indptr, const I* __restrict__ src, I* __restrict__ out,
                                const I* __restrict__ src2, I* __restrict__ out2) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 7;
  if (i >= n) return;
  const int64_t b = seeds ? (int64_t)starts[seeds[i]] : (int64_t)starts[i];
  const int64_t o = out_indptr[i], len = out_indptr[i + 1] - o;
  for (int64_t q = lane; q < len; q += 8) {
    out[o + q] = src[b + q];
    if (src2) out2[o + q] = src2[b + q];
  }
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_graphcache_query(int idtype, void* keys, void* counts, const void* offs, const void* lens,
                                      int64_t slots, const void* seeds, int64_t n, void* hit, void* off_out,
                                      void* len_out, void* stats, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    gc_query_kernel<I><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (int64_t*)keys, (int32_t*)counts, (const int64_t*)offs, (const int32_t*)lens, slots, (const I*)seeds, n,
        (uint8_t*)hit, (int64_t*)off_out, (int32_t*)len_out, (int64_t*)stats);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_graphcache_replace(int idtype, void* keys, const void* counts, void* offs, void* lens, int64_t slots,
                                        int threshold, const void* seeds, int64_t n, const void* indptr,
                                        const void* indices, void* pool, int64_t pool_cap, void* used, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    gc_replace_kernel<I><<<(unsigned)((n * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (int64_t*)keys, (const int32_t*)counts, (int64_t*)offs, (int32_t*)lens, slots, threshold, (const I*)seeds, n,
        (const int64_t*)indptr, (const I*)indices, (I*)pool, pool_cap, (int64_t*)used);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_csc_degree(int idtype, const void* indptr, const void* seeds, int64_t n, void* deg, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    csc_degree_kernel<I><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const I*)indptr, (const I*)seeds, n, (int64_t*)deg);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// starts_are_offsets = 0: starts = graph indptr (idtype) indexed by seeds; 1: starts = int64 offsets, one per output row
DGLB_API int dglb200_cuda_seg_copy(int idtype, int starts_are_offsets, const void* starts, const void* seeds, int64_t n,
                                   const void* out_indptr, const void* src, void* out, const void* src2, void* out2,
                                   void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n * 8 + 255) / 256);
  DGLB_ID_SWITCH(idtype, I, {
    if (starts_are_offsets)
      seg_copy_kernel<I, int64_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const int64_t*)starts, nullptr, n, (const int64_t*)out_indptr, (const I*)src, (I*)out, (const I*)src2, (I*)out2);
    else
      seg_copy_kernel<I, I><<<blocks, 256, 0, (cudaStream_t)stream>>>((const I*)starts, (const I*)seeds, n, (const int64_t*)out_indptr, (const I*)src, (I*)out, (const I*)src2, (I*)out2);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// device-visible alias of a host-pinned (cudaHostAlloc / cudaHostRegister) pointer
DGLB_API int dglb200_cuda_host_device_ptr(const void* host, void** out) {
  DGLB_TRY
  DGLB_CUDA_CHECK(cudaHostGetDevicePointer(out, const_cast<void*>(host), 0));
  DGLB_CATCH
}
