// Sparse-structure construction on the GPU: COO -> CSR / CSC (stable radix sort with the edge-id permutation),
// coalescing of duplicate coordinates (with value reduction), CSR + CSR and a hash-table SpGEMM (CSR x CSR).
//
// Replaces the torch.sort / torch.sparse.mm calls the graph index and dgl.sparse used on CUDA in round 1, and plays the
// role of the reference's src/array/cuda/{coo2csr,coo_sort,csr_sort,csr_sum,csr_mm}.cu (cuSPARSE-based there). The only
// library primitive is CUB's device radix sort / scan (header-only, ships with the toolkit); everything else - key
// packing, head flags, segment reduction, the per-row hash accumulation of the SpGEMM - is this file's kernels.
// Temporary storage is stream-ordered (cudaMallocAsync), so calls enqueue without host synchronisation; routines whose
// OUTPUT size is data dependent return it in a device counter and the caller trims the (upper-bound sized) outputs.
#include <cub/cub.cuh>

#include "common.cuh"

namespace dglb {
namespace sb {

// The default stream-ordered pool gives memory back to the driver at every synchronisation (release threshold 0), which
// made each cudaMallocAsync a ~1.8 ms driver call in a loop that builds one CSR per sampled block (measured with
// torch.profiler on the GAT mini-batch pipeline). Keep the pool's memory: set the threshold once per device.
static void keep_pool_memory() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[dev] = true;
}

struct Scratch {
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t s) : st(s) { keep_pool_memory(); }
  template <typename T>
  T* get(size_t n) {
    void* p = nullptr;
    DGLB_CUDA_CHECK(cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st));
    ptrs.push_back(p);
    return static_cast<T*>(p);
  }
  ~Scratch() {
    for (void* p : ptrs) cudaFreeAsync(p, st);
  }
};

static int bits_for(int64_t n) {
  int b = 1;
  while (b < 63 && ((int64_t)1 << b) < n) ++b;
  return b;
}

template <typename I>
__global__ void iota_kernel(I* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (I)i;
}

template <typename I>
__global__ void gather_kernel(const I* __restrict__ src, const I* __restrict__ perm, I* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[perm[i]];
}

// indptr[r] = first position p with sorted_row[p] >= r  (sorted_row ascending): one binary search per row boundary
template <typename I>
__global__ void row_bounds_kernel(const I* __restrict__ sorted_row, int64_t nnz, int64_t num_rows, I* __restrict__ indptr) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= num_rows; r += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)sorted_row[mid] < r) lo = mid + 1; else hi = mid;
    }
    indptr[r] = (I)lo;
  }
}

template <typename I>
__global__ void expand_rows_kernel(const I* __restrict__ indptr, int64_t num_rows, I* __restrict__ row) {
  // one warp per row keeps long rows coalesced
  const int lane = threadIdx.x & 31;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < num_rows; r += ((int64_t)gridDim.x * blockDim.x) >> 5)
    for (int64_t p = indptr[r] + lane; p < indptr[r + 1]; p += 32) row[p] = (I)r;
}

static unsigned grid_for(int64_t n, int block = 256) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, 148 * 16));
}

// stable sort of (row, edge id) by row; out: indptr, indices = col[perm], eids = perm
template <typename I>
void coo2csr(const I* row, const I* col, int64_t nnz, int64_t num_rows, I* indptr, I* indices, I* eids, cudaStream_t st) {
  Scratch s(st);
  I* ids = s.get<I>(nnz);
  I* srow = s.get<I>(nnz);
  iota_kernel<I><<<grid_for(nnz), 256, 0, st>>>(ids, nnz);
  size_t tb = 0;
  const int end_bit = bits_for(std::max<int64_t>(num_rows, 2));
  cub::DeviceRadixSort::SortPairs(nullptr, tb, row, srow, ids, eids, (int)nnz, 0, end_bit, st);
  void* tmp = s.get<char>(tb);
  cub::DeviceRadixSort::SortPairs(tmp, tb, row, srow, ids, eids, (int)nnz, 0, end_bit, st);
  gather_kernel<I><<<grid_for(nnz), 256, 0, st>>>(col, eids, indices, nnz);
  row_bounds_kernel<I><<<grid_for(num_rows + 1), 256, 0, st>>>(srow, nnz, num_rows, indptr);
}

// ------------------------------------------------------------------------------------------------------------ coalesce
template <typename I>
__global__ void pack_keys_kernel(const I* __restrict__ row, const I* __restrict__ col, int64_t n, int64_t num_cols,
                                 uint64_t* __restrict__ key) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    key[i] = (uint64_t)row[i] * (uint64_t)num_cols + (uint64_t)col[i];
}

__global__ void head_flags_kernel(const uint64_t* __restrict__ key, int64_t n, int32_t* __restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// pos[i] = inclusive scan of flags: output slot of element i is pos[i] - 1. Values of equal keys are summed.
template <typename I>
__global__ void emit_unique_kernel(const uint64_t* __restrict__ key, const int32_t* __restrict__ flag,
                                   const int64_t* __restrict__ pos, const int64_t* __restrict__ perm,
                                   const float* __restrict__ val, int64_t n, int64_t num_cols, int vdim,
                                   I* __restrict__ out_row, I* __restrict__ out_col, float* __restrict__ out_val,
                                   int64_t* __restrict__ out_count, int64_t* __restrict__ out_first) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = pos[i] - 1;
    if (flag[i]) {
      out_row[slot] = (I)(key[i] / (uint64_t)num_cols);
      out_col[slot] = (I)(key[i] % (uint64_t)num_cols);
      if (out_first) out_first[slot] = perm[i];
    }
    if (val)
      for (int d = 0; d < vdim; ++d) atomicAdd(out_val + slot * vdim + d, val[perm[i] * vdim + d]);
    if (i == n - 1) *out_count = pos[i];
  }
}

// sort by (row, col), merge duplicates; out arrays sized nnz (upper bound), out_val pre-zeroed by the caller
template <typename I>
void coalesce(const I* row, const I* col, const float* val, int vdim, int64_t nnz, int64_t num_rows, int64_t num_cols,
              I* out_row, I* out_col, float* out_val, int64_t* out_first, int64_t* out_count, cudaStream_t st) {
  Scratch s(st);
  uint64_t* key = s.get<uint64_t>(nnz);
  uint64_t* skey = s.get<uint64_t>(nnz);
  int64_t* ids = s.get<int64_t>(nnz);
  int64_t* perm = s.get<int64_t>(nnz);
  int32_t* flag = s.get<int32_t>(nnz);
  int64_t* pos = s.get<int64_t>(nnz);
  pack_keys_kernel<I><<<grid_for(nnz), 256, 0, st>>>(row, col, nnz, num_cols, key);
  iota_kernel<int64_t><<<grid_for(nnz), 256, 0, st>>>(ids, nnz);
  const int end_bit = std::min(64, bits_for(std::max<int64_t>(num_rows, 2)) + bits_for(std::max<int64_t>(num_cols, 2)));
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, key, skey, ids, perm, (int)nnz, 0, end_bit, st);
  void* tmp = s.get<char>(tb);
  cub::DeviceRadixSort::SortPairs(tmp, tb, key, skey, ids, perm, (int)nnz, 0, end_bit, st);
  head_flags_kernel<<<grid_for(nnz), 256, 0, st>>>(skey, nnz, flag);
  size_t sb_ = 0;
  cub::DeviceScan::InclusiveSum(nullptr, sb_, flag, pos, (int)nnz, st);
  void* stmp = s.get<char>(sb_);
  cub::DeviceScan::InclusiveSum(stmp, sb_, flag, pos, (int)nnz, st);
  emit_unique_kernel<I><<<grid_for(nnz), 256, 0, st>>>(skey, flag, pos, perm, val, nnz, num_cols, vdim, out_row, out_col,
                                                       out_val, out_count, out_first);
}

// ------------------------------------------------------------------------------------------------------------- SpGEMM
// C = A @ B, all CSR. Phase 1: upper bound of every row's non-zeros (sum of the B-row lengths it touches). Phase 2: one
// warp per row accumulates into the row's private open-addressing table (keys = column, values = fp32 partial sums)
// carved out of one scratch buffer at the exclusive scan of the table sizes. Phase 3: compaction into COO triplets
// (the caller sorts them into canonical CSR with `coalesce`).
template <typename I>
__global__ void spgemm_bound_kernel(const I* __restrict__ ap, const I* __restrict__ ai, const I* __restrict__ bp,
                                    int64_t rows, int64_t* __restrict__ cap) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    int64_t ub = 0;
    for (int64_t p = ap[r]; p < ap[r + 1]; ++p) ub += bp[ai[p] + 1] - bp[ai[p]];
    int64_t c = 0;
    if (ub > 0) {
      c = 8;
      while (c < 2 * ub) c <<= 1;  // load factor <= 0.5
    }
    cap[r] = c;
  }
}

template <typename I>
__global__ void spgemm_accumulate_kernel(const I* __restrict__ ap, const I* __restrict__ ai, const float* __restrict__ av,
                                         const I* __restrict__ bp, const I* __restrict__ bi, const float* __restrict__ bv,
                                         int64_t rows, const int64_t* __restrict__ off, const int64_t* __restrict__ cap,
                                         int64_t* __restrict__ keys, float* __restrict__ vals, int64_t* __restrict__ row_nnz) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t c = cap[r];
    if (c == 0) {
      if (lane == 0) row_nnz[r] = 0;
      continue;
    }
    int64_t* k = keys + off[r];
    float* v = vals + off[r];
    int64_t fresh = 0;
    for (int64_t p = ap[r]; p < ap[r + 1]; ++p) {
      const int64_t mid = ai[p];
      const float a = av ? av[p] : 1.f;
      for (int64_t q = bp[mid] + lane; q < bp[mid + 1]; q += 32) {
        const int64_t col = bi[q];
        const float contrib = a * (bv ? bv[q] : 1.f);
        int64_t h = (col * 0x9E3779B97F4A7C15ull >> 20) & (c - 1);
        while (true) {
          const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(k + h), ~0ull, (unsigned long long)col);
          if (prev == ~0ull) ++fresh;
          if (prev == ~0ull || prev == (unsigned long long)col) {
            atomicAdd(v + h, contrib);
            break;
          }
          h = (h + 1) & (c - 1);
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    if (lane == 0) row_nnz[r] = fresh;
  }
}

template <typename I>
__global__ void spgemm_compact_kernel(int64_t rows, const int64_t* __restrict__ off, const int64_t* __restrict__ cap,
                                      const int64_t* __restrict__ out_off, const int64_t* __restrict__ keys,
                                      const float* __restrict__ vals, I* __restrict__ out_row, I* __restrict__ out_col,
                                      float* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    int64_t w = out_off[r];
    for (int64_t base = 0; base < cap[r]; base += 32) {
      const int64_t h = base + lane;
      const bool live = h < cap[r] && keys[off[r] + h] != -1;
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) {
        const int64_t slot = w + __popc(m & ((1u << lane) - 1));
        out_row[slot] = (I)r;
        out_col[slot] = (I)keys[off[r] + h];
        out_val[slot] = vals[off[r] + h];
      }
      w += __popc(m);
    }
  }
}

}  // namespace sb
}  // namespace dglb

using namespace dglb;
using namespace dglb::sb;

// row / col: [nnz] ids (row need not be sorted). Outputs: indptr [num_rows + 1], indices [nnz], eids [nnz].
DGLB_API int dglb200_cuda_coo2csr(int idtype, const void* row, const void* col, int64_t nnz, int64_t num_rows, void* indptr,
                                  void* indices, void* eids, void* stream) {
  DGLB_TRY
  DGLB_CHECK(nnz < (int64_t)1 << 31, "coo2csr: more than 2^31 - 1 edges need a segmented build");
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_ID_SWITCH(idtype, I, {
    if (nnz == 0) {
      DGLB_CUDA_CHECK(cudaMemsetAsync(indptr, 0, (size_t)(num_rows + 1) * sizeof(I), st));
    } else {
      coo2csr<I>((const I*)row, (const I*)col, nnz, num_rows, (I*)indptr, (I*)indices, (I*)eids, st);
    }
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// CSR (indptr, indices) of an [num_rows x num_cols] matrix -> CSR of its transpose; out_eids[p] = position of the
// transposed entry p in the input
DGLB_API int dglb200_cuda_csr_transpose(int idtype, const void* indptr, const void* indices, int64_t nnz, int64_t num_rows,
                                        int64_t num_cols, void* out_indptr, void* out_indices, void* out_eids, void* stream) {
  DGLB_TRY
  DGLB_CHECK(nnz < (int64_t)1 << 31, "csr_transpose: more than 2^31 - 1 entries");
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_ID_SWITCH(idtype, I, {
    if (nnz == 0) {
      DGLB_CUDA_CHECK(cudaMemsetAsync(out_indptr, 0, (size_t)(num_cols + 1) * sizeof(I), st));
    } else {
      Scratch s(st);
      I* row = s.get<I>(nnz);
      expand_rows_kernel<I><<<grid_for(num_rows * 32), 256, 0, st>>>((const I*)indptr, num_rows, row);
      coo2csr<I>((const I*)indices, row, nnz, num_cols, (I*)out_indptr, (I*)out_indices, (I*)out_eids, st);
    }
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// Sort COO entries by (row, col) and merge duplicates. val: fp32 [nnz, vdim] or null. Outputs sized nnz; out_val must be
// zero-filled; out_first[k] = input position of one entry of group k (null to skip); out_count: device int64.
DGLB_API int dglb200_cuda_coo_coalesce(int idtype, const void* row, const void* col, const void* val, int vdim, int64_t nnz,
                                       int64_t num_rows, int64_t num_cols, void* out_row, void* out_col, void* out_val,
                                       void* out_first, void* out_count, void* stream) {
  DGLB_TRY
  DGLB_CHECK(nnz < (int64_t)1 << 31, "coalesce: more than 2^31 - 1 entries");
  cudaStream_t st = (cudaStream_t)stream;
  if (nnz == 0) {
    DGLB_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st));
    return 0;
  }
  DGLB_ID_SWITCH(idtype, I, {
    coalesce<I>((const I*)row, (const I*)col, (const float*)val, vdim, nnz, num_rows, num_cols, (I*)out_row, (I*)out_col,
                (float*)out_val, (int64_t*)out_first, (int64_t*)out_count, st);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// SpGEMM phase A: per-row table capacities (exclusive-scanned into `off`, total in off[rows]) - the caller allocates the
// tables from the total. cap / off: int64 [rows] / [rows + 1].
DGLB_API int dglb200_cuda_spgemm_plan(int idtype, const void* ap, const void* ai, const void* bp, int64_t rows, void* cap,
                                      void* off, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_ID_SWITCH(idtype, I, {
    spgemm_bound_kernel<I><<<grid_for(rows), 256, 0, st>>>((const I*)ap, (const I*)ai, (const I*)bp, rows, (int64_t*)cap);
  });
  Scratch s(st);
  size_t tb = 0;
  DGLB_CUDA_CHECK(cudaMemsetAsync(off, 0, sizeof(int64_t), st));
  cub::DeviceScan::InclusiveSum(nullptr, tb, (const int64_t*)cap, (int64_t*)off + 1, (int)rows, st);
  void* tmp = s.get<char>(tb);
  cub::DeviceScan::InclusiveSum(tmp, tb, (const int64_t*)cap, (int64_t*)off + 1, (int)rows, st);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// phase B: accumulate (keys pre-filled with -1, vals with 0), row_nnz [rows] and its exclusive scan out_off [rows + 1]
DGLB_API int dglb200_cuda_spgemm_accumulate(int idtype, const void* ap, const void* ai, const void* av, const void* bp,
                                            const void* bi, const void* bv, int64_t rows, const void* off, const void* cap,
                                            void* keys, void* vals, void* row_nnz, void* out_off, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_ID_SWITCH(idtype, I, {
    spgemm_accumulate_kernel<I><<<grid_for(rows * 32), 256, 0, st>>>(
        (const I*)ap, (const I*)ai, (const float*)av, (const I*)bp, (const I*)bi, (const float*)bv, rows, (const int64_t*)off,
        (const int64_t*)cap, (int64_t*)keys, (float*)vals, (int64_t*)row_nnz);
  });
  Scratch s(st);
  size_t tb = 0;
  DGLB_CUDA_CHECK(cudaMemsetAsync(out_off, 0, sizeof(int64_t), st));
  cub::DeviceScan::InclusiveSum(nullptr, tb, (const int64_t*)row_nnz, (int64_t*)out_off + 1, (int)rows, st);
  void* tmp = s.get<char>(tb);
  cub::DeviceScan::InclusiveSum(tmp, tb, (const int64_t*)row_nnz, (int64_t*)out_off + 1, (int)rows, st);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// phase C: tables -> COO triplets (row-grouped, columns unordered inside a row)
DGLB_API int dglb200_cuda_spgemm_compact(int idtype, int64_t rows, const void* off, const void* cap, const void* out_off,
                                         const void* keys, const void* vals, void* out_row, void* out_col, void* out_val,
                                         void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_ID_SWITCH(idtype, I, {
    spgemm_compact_kernel<I><<<grid_for(rows * 32), 256, 0, st>>>(rows, (const int64_t*)off, (const int64_t*)cap,
                                                                 (const int64_t*)out_off, (const int64_t*)keys,
                                                                 (const float*)vals, (I*)out_row, (I*)out_col, (float*)out_val);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
