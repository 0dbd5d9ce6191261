// Structure kernels for sm_100a: vectorised row gather/scatter (device or UVA host-pinned
// source), order-preserving hash compaction (unique + relabel for block/MFG construction),
// id -> position mapping, CSR helpers.
// Behavioural counterparts in the reference: src/runtime/cuda/cuda_hashtable.cu:179-295,
// src/graph/transform/cuda/cuda_map_edges.cuh:80, src/array/cuda/array_index_select.cuh:15-64,
// src/array/cuda/uvm/array_index_select_uvm.cuh:22, graphbolt/src/cuda/index_select_impl.cu:34-205.
// Design differences: one table entry stores (key, first-occurrence position); ordering is
// recovered with a flag + exclusive scan, so no per-block count/compact passes and no host sync
// (the unique count stays on the device for the caller to consume).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace dglb {

// ------------------------------------------------------------------ gather / scatter
template <typename I, int VB>
__global__ void __launch_bounds__(256) gather_rows_kernel(const char* __restrict__ src, const I* __restrict__ idx,
                                                          int64_t n, char* __restrict__ dst, int64_t row_bytes,
                                                          int64_t src_rows, int sub,
                                                          const int64_t* __restrict__ dst_pos = nullptr) {
  using V = typename std::conditional<VB == 16, uint4, typename std::conditional<VB == 8, uint2,
            typename std::conditional<VB == 4, uint32_t, uint8_t>::type>::type>::type;
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int64_t nvec = row_bytes / VB;
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; r < n;
       r += (int64_t)gridDim.x * rows_per_block) {
    int64_t s = idx[r];
    // dst_pos: the gather walks the source in the (sorted) order of idx and writes row r to its original position
    const int64_t orow = dst_pos ? dst_pos[r] : r;
    if (s < 0 || s >= src_rows) {  // out-of-range ids gather zeros (padding rows)
      V z;
      memset(&z, 0, sizeof(V));
      for (int64_t c = lane; c < nvec; c += sub) reinterpret_cast<V*>(dst + orow * row_bytes)[c] = z;
      continue;
    }
    const V* sp = reinterpret_cast<const V*>(src + s * row_bytes);
    V* dp = reinterpret_cast<V*>(dst + orow * row_bytes);
    int64_t c = lane;
    for (; c + 3 * sub < nvec; c += 4 * sub) {  // 4 loads in flight per lane
      V a = sp[c], b = sp[c + sub], d = sp[c + 2 * sub], e = sp[c + 3 * sub];
      dp[c] = a; dp[c + sub] = b; dp[c + 2 * sub] = d; dp[c + 3 * sub] = e;
    }
    for (; c < nvec; c += sub) dp[c] = sp[c];
  }
}

template <typename I, int VB>
__global__ void __launch_bounds__(256) scatter_rows_kernel(char* __restrict__ dst, const I* __restrict__ idx, int64_t n,
                                                           const char* __restrict__ src, int64_t row_bytes, int sub) {
  using V = typename std::conditional<VB == 16, uint4, typename std::conditional<VB == 8, uint2,
            typename std::conditional<VB == 4, uint32_t, uint8_t>::type>::type>::type;
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int64_t nvec = row_bytes / VB;
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; r < n;
       r += (int64_t)gridDim.x * rows_per_block) {
    const int64_t d = idx[r];
    const V* sp = reinterpret_cast<const V*>(src + r * row_bytes);
    V* dp = reinterpret_cast<V*>(dst + d * row_bytes);
    for (int64_t c = lane; c < nvec; c += sub) dp[c] = sp[c];
  }
}

static int pick_vb(int64_t row_bytes, const void* a, const void* b) {
  const uintptr_t x = (uintptr_t)a | (uintptr_t)b | (uintptr_t)row_bytes;
  if (x % 16 == 0) return 16;
  if (x % 8 == 0) return 8;
  if (x % 4 == 0) return 4;
  return 1;
}

template <typename I>
void launch_gather(const void* src, const I* idx, int64_t n, void* dst, int64_t row_bytes, int64_t src_rows,
                   int max_blocks, cudaStream_t st, const int64_t* dst_pos = nullptr) {
  const int vb = pick_vb(row_bytes, src, dst);
  const int64_t nvec = row_bytes / vb;
  int sub = 1;
  while (sub < nvec && sub < 32) sub <<= 1;
  const int rpb = 256 / sub;
  int64_t blocks = (n + rpb - 1) / rpb;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
#define G(VBV) gather_rows_kernel<I, VBV><<<(unsigned)blocks, 256, 0, st>>>((const char*)src, idx, n, (char*)dst, row_bytes, src_rows, sub, dst_pos)
  if (vb == 16) G(16); else if (vb == 8) G(8); else if (vb == 4) G(4); else G(1);
#undef G
}

template <typename I>
void launch_scatter(void* dst, const I* idx, int64_t n, const void* src, int64_t row_bytes, cudaStream_t st) {
  const int vb = pick_vb(row_bytes, src, dst);
  const int64_t nvec = row_bytes / vb;
  int sub = 1;
  while (sub < nvec && sub < 32) sub <<= 1;
  const int rpb = 256 / sub;
  int64_t blocks = (n + rpb - 1) / rpb;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
#define S(VBV) scatter_rows_kernel<I, VBV><<<(unsigned)blocks, 256, 0, st>>>((char*)dst, idx, n, (const char*)src, row_bytes, sub)
  if (vb == 16) S(16); else if (vb == 8) S(8); else if (vb == 4) S(4); else S(1);
#undef S
}

// ------------------------------------------------------------------ ordered hash table
__device__ __forceinline__ uint64_t hash64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  return k;
}

template <typename I>
__device__ __forceinline__ I atomic_cas(I* addr, I cmp, I val);
template <>
__device__ __forceinline__ int32_t atomic_cas<int32_t>(int32_t* addr, int32_t cmp, int32_t val) {
  return atomicCAS(addr, cmp, val);
}
template <>
__device__ __forceinline__ int64_t atomic_cas<int64_t>(int64_t* addr, int64_t cmp, int64_t val) {
  return (int64_t)atomicCAS((unsigned long long*)addr, (unsigned long long)cmp, (unsigned long long)val);
}

// n_total is read from device memory when n_dev != nullptr (sync-free pipelines)
template <typename I>
__global__ void __launch_bounds__(256) ht_insert_kernel(const I* __restrict__ first, int64_t nfirst,
                                                        const I* __restrict__ ids, int64_t n_cap,
                                                        const int64_t* __restrict__ n_dev, I* __restrict__ keys,
                                                        int32_t* __restrict__ minpos, int32_t* __restrict__ slot,
                                                        uint64_t mask) {
  const int64_t n = nfirst + (n_dev ? *n_dev : n_cap);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const I id = i < nfirst ? first[i] : ids[i - nfirst];
    if (id < 0) {  // padding
      slot[i] = -1;
      continue;
    }
    uint64_t h = hash64((uint64_t)id) & mask;
    while (true) {
      const I old = atomic_cas<I>(&keys[h], (I)-1, id);
      if (old == (I)-1 || old == id) {
        atomicMin(&minpos[h], (int32_t)i);
        slot[i] = (int32_t)h;
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

template <typename I>
__global__ void __launch_bounds__(256) ht_flag_kernel(int64_t nfirst, int64_t n_cap, const int64_t* __restrict__ n_dev,
                                                      const int32_t* __restrict__ minpos,
                                                      const int32_t* __restrict__ slot, int32_t* __restrict__ flag,
                                                      int64_t n_alloc) {
  const int64_t n = nfirst + (n_dev ? *n_dev : n_cap);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_alloc; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t f = 0;
    if (i < n) {
      const int32_t s = slot[i];
      f = (s >= 0 && minpos[s] == (int32_t)i) ? 1 : 0;
    }
    flag[i] = f;
  }
}

template <typename I>
__global__ void __launch_bounds__(256) ht_assign_kernel(const I* __restrict__ first, int64_t nfirst,
                                                        const I* __restrict__ ids, int64_t n_cap,
                                                        const int64_t* __restrict__ n_dev,
                                                        const int32_t* __restrict__ slot,
                                                        const int32_t* __restrict__ flag,
                                                        const int32_t* __restrict__ pos, int32_t* __restrict__ local,
                                                        I* __restrict__ out_unique, int64_t* __restrict__ out_count,
                                                        int64_t n_alloc, int64_t out_cap, int32_t* __restrict__ overflow) {
  const int64_t n = nfirst + (n_dev ? *n_dev : n_cap);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (flag[i]) {
      const I id = i < nfirst ? first[i] : ids[i - nfirst];
      if (pos[i] < out_cap) {
        out_unique[pos[i]] = id;
        local[slot[i]] = pos[i];
      } else {
        local[slot[i]] = -1;  // dropped: output capacity exceeded (flagged below)
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t total = n_alloc > 0 ? (int64_t)pos[n_alloc - 1] + flag[n_alloc - 1] : 0;
    *out_count = total < out_cap ? total : out_cap;
    if (total > out_cap && overflow) atomicExch(overflow, 1);
  }
}

template <typename I>
__global__ void __launch_bounds__(256) ht_lookup_kernel(int64_t nfirst, int64_t n_cap,
                                                        const int64_t* __restrict__ n_dev,
                                                        const int32_t* __restrict__ slot,
                                                        const int32_t* __restrict__ local, I* __restrict__ out_local) {
  const int64_t n = n_dev ? *n_dev : n_cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cap; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n) {
      const int32_t s = slot[nfirst + i];
      out_local[i] = s >= 0 ? (I)local[s] : (I)-1;
    } else {
      out_local[i] = (I)-1;
    }
  }
}

template <typename I>
__global__ void __launch_bounds__(256) ht_insert_kv_kernel(const I* __restrict__ universe, int64_t nu,
                                                           I* __restrict__ keys, int32_t* __restrict__ vals,
                                                           uint64_t mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nu; i += (int64_t)gridDim.x * blockDim.x) {
    const I id = universe[i];
    uint64_t h = hash64((uint64_t)id) & mask;
    while (true) {
      const I old = atomic_cas<I>(&keys[h], (I)-1, id);
      if (old == (I)-1 || old == id) {
        atomicMin(&vals[h], (int32_t)i);
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

template <typename I>
__global__ void __launch_bounds__(256) ht_find_kernel(const I* __restrict__ ids, int64_t n, const I* __restrict__ keys,
                                                      const int32_t* __restrict__ vals, uint64_t mask,
                                                      I* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const I id = ids[i];
    I res = -1;
    if (id >= 0) {
      uint64_t h = hash64((uint64_t)id) & mask;
      while (true) {
        const I k = keys[h];
        if (k == id) { res = (I)vals[h]; break; }
        if (k == (I)-1) break;
        h = (h + 1) & mask;
      }
    }
    out[i] = res;
  }
}

static inline unsigned grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace dglb

using namespace dglb;

// src may be device memory or host-pinned memory (then `uva`=1 and the device alias is resolved here)
DGLB_API int dglb200_cuda_gather_rows(int idtype, const void* src, const void* idx, int64_t n, void* dst,
                                      int64_t row_bytes, int64_t src_rows, int uva, int max_blocks, const void* dst_pos,
                                      void* stream) {
  DGLB_TRY
  const void* sp = src;
  if (uva) {
    void* dp = nullptr;
    DGLB_CUDA_CHECK(cudaHostGetDevicePointer(&dp, const_cast<void*>(src), 0));
    sp = dp;
  }
  if (max_blocks <= 0) max_blocks = 148 * 32;
  DGLB_ID_SWITCH(idtype, I, { launch_gather<I>(sp, (const I*)idx, n, dst, row_bytes, src_rows, max_blocks, (cudaStream_t)stream, (const int64_t*)dst_pos); });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_scatter_rows(int idtype, void* dst, const void* idx, int64_t n, const void* src,
                                       int64_t row_bytes, int uva, void* stream) {
  DGLB_TRY
  void* dp = dst;
  if (uva) DGLB_CUDA_CHECK(cudaHostGetDevicePointer(&dp, dst, 0));
  DGLB_ID_SWITCH(idtype, I, { launch_scatter<I>(dp, (const I*)idx, n, src, row_bytes, (cudaStream_t)stream); });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int64_t dglb200_cuda_scan_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  return (int64_t)bytes;
}

// Order-preserving unique + relabel. Workspace (all caller-allocated):
//   keys[cap] (idtype, pre-filled -1), minpos[cap] int32 (pre-filled INT_MAX), local[cap] int32,
//   slot/flag/pos [nfirst+n] int32, scan temp. cap must be a power of two >= 2*(nfirst+n).
// n_dev (optional device int64) overrides n at run time for sync-free pipelines.
DGLB_API int dglb200_cuda_unique_compact(int idtype, const void* first, int64_t nfirst, const void* ids, int64_t n,
                                         const void* n_dev, void* keys, void* minpos, void* local, int64_t cap,
                                         void* slot, void* flag, void* pos, void* scan_temp, int64_t scan_bytes,
                                         void* out_unique, void* out_local, void* out_count, int64_t out_cap,
                                         void* overflow, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = nfirst + n;
  DGLB_CHECK((cap & (cap - 1)) == 0 && cap >= 2 * N, "hash capacity must be a power of two >= 2N");
  if (N == 0) {
    DGLB_CUDA_CHECK(cudaMemsetAsync(out_count, 0, 8, st));
    return 0;
  }
  const uint64_t mask = (uint64_t)cap - 1;
  {
    // (re)initialise the table and the output inside the stream so the call is graph-replayable
    const size_t isz = idtype == kI32 ? 4 : 8;
    DGLB_CUDA_CHECK(cudaMemsetAsync(keys, 0xFF, (size_t)cap * isz, st));
    DGLB_CUDA_CHECK(cudaMemsetAsync(minpos, 0x7F, (size_t)cap * 4, st));
    DGLB_CUDA_CHECK(cudaMemsetAsync(out_unique, 0xFF, (size_t)(out_cap > 0 && out_cap < N ? out_cap : N) * isz, st));
  }
  DGLB_ID_SWITCH(idtype, I, {
    ht_insert_kernel<I><<<grid_for(N), 256, 0, st>>>((const I*)first, nfirst, (const I*)ids, n, (const int64_t*)n_dev,
                                                     (I*)keys, (int32_t*)minpos, (int32_t*)slot, mask);
    ht_flag_kernel<I><<<grid_for(N), 256, 0, st>>>(nfirst, n, (const int64_t*)n_dev, (const int32_t*)minpos,
                                                   (const int32_t*)slot, (int32_t*)flag, N);
    size_t tb = (size_t)scan_bytes;
    DGLB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_temp, tb, (const int32_t*)flag, (int32_t*)pos, (int)N, st));
    ht_assign_kernel<I><<<grid_for(N), 256, 0, st>>>((const I*)first, nfirst, (const I*)ids, n, (const int64_t*)n_dev,
                                                     (const int32_t*)slot, (const int32_t*)flag, (const int32_t*)pos,
                                                     (int32_t*)local, (I*)out_unique, (int64_t*)out_count, N,
                                                     out_cap > 0 ? out_cap : N, (int32_t*)overflow);
    if (n > 0)
      ht_lookup_kernel<I><<<grid_for(n), 256, 0, st>>>(nfirst, n, (const int64_t*)n_dev, (const int32_t*)slot,
                                                       (const int32_t*)local, (I*)out_local);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// out[i] = position of ids[i] in universe (-1 if absent). keys pre-filled -1, vals pre-filled INT_MAX.
DGLB_API int dglb200_cuda_map_ids(int idtype, const void* universe, int64_t nu, const void* ids, int64_t n, void* keys,
                                  void* vals, int64_t cap, void* out, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_CHECK((cap & (cap - 1)) == 0 && cap >= 2 * nu, "hash capacity must be a power of two >= 2*nu");
  const uint64_t mask = (uint64_t)cap - 1;
  {
    const size_t isz = idtype == kI32 ? 4 : 8;
    DGLB_CUDA_CHECK(cudaMemsetAsync(keys, 0xFF, (size_t)cap * isz, st));
    DGLB_CUDA_CHECK(cudaMemsetAsync(vals, 0x7F, (size_t)cap * 4, st));
  }
  DGLB_ID_SWITCH(idtype, I, {
    if (nu > 0) ht_insert_kv_kernel<I><<<grid_for(nu), 256, 0, st>>>((const I*)universe, nu, (I*)keys, (int32_t*)vals, mask);
    if (n > 0) ht_find_kernel<I><<<grid_for(n), 256, 0, st>>>((const I*)ids, n, (const I*)keys, (const int32_t*)vals, mask, (I*)out);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
