// Frontier compaction for the sync-free mini-batch engine: ONE cooperative kernel per layer.
//
//   S_{l-1} = S_l  ++  { neighbour ids not in S_l }          (dst nodes stay a prefix of the src nodes)
//   nbr_local[d, k] = position of nbr_global[d, k] in S_{l-1}
//
// The reference does this with a sort-free hash map too (to_block / unique_and_compact,
// /root/reference/src/graph/transform/cuda/cuda_to_block.cu:59-239, graphbolt/src/cuda/unique_and_compact_impl.cu:56-313)
// but as a chain of kernels with host-visible counts. Here the order of the NEW nodes is not significant for the engine
// (rows of an activation matrix), so positions come from a warp-aggregated atomic counter instead of flag/scan/assign
// passes, and the four phases (insert, assign, relabel, clean-up) are separated by grid-wide barriers inside a single
// cooperative launch: no memsets, no scan, no host involvement, CUDA-graph capturable. The hash table is left clean
// (keys = -1, vals = kPending) for the next call, so it is initialised exactly once at allocation time.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dglb {

constexpr int32_t kPending = -2;

__device__ __forceinline__ uint32_t hash32(uint32_t k) {
  k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
  return k;
}

// slot[] encoding: -1 = padding item, h >= 0 = found/inserted by someone else, -(h + 2) = this item won the insertion
__global__ void __launch_bounds__(256) frontier_compact_kernel(
    const int32_t* __restrict__ first, const int64_t* __restrict__ n_first_dev, int64_t first_cap,
    const int32_t* __restrict__ ids, int fanout, int32_t* __restrict__ keys, int32_t* __restrict__ vals, uint32_t mask,
    int32_t* __restrict__ slot, int32_t* __restrict__ out_unique, int64_t out_cap, int32_t* __restrict__ out_local,
    int64_t* __restrict__ out_count, int32_t* __restrict__ counter, int32_t* __restrict__ overflow) {
  cg::grid_group grid = cg::this_grid();
  const int64_t nf = min(*n_first_dev, first_cap);
  const int64_t nn = nf * fanout;       // live neighbour items
  const int64_t N = nf + nn;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;

  // ---- phase 0: insert. Seeds own position i; neighbour items only remember whether they created the entry.
  for (int64_t i = tid; i < N; i += nthreads) {
    const bool is_seed = i < nf;
    const int32_t id = is_seed ? first[i] : ids[i - nf];
    if (id < 0) {
      slot[i] = -1;
      continue;
    }
    uint32_t h = hash32((uint32_t)id) & mask;
    while (true) {
      const int32_t old = atomicCAS(&keys[h], -1, id);
      if (old == -1 || old == id) {
        if (is_seed) {
          vals[h] = (int32_t)i;  // a seed's position always wins over "pending"
          out_unique[i] = id;
          slot[i] = (int32_t)h;
        } else {
          slot[i] = old == -1 ? -(int32_t)(h + 2) : (int32_t)h;
        }
        break;
      }
      h = (h + 1) & mask;
    }
  }
  grid.sync();
  // ---- phase 1: every entry created by a neighbour item that no seed claimed gets the next free position
  for (int64_t i0 = nf + (tid - lane); i0 < N; i0 += nthreads) {  // warp-uniform trip count
    const int64_t i = i0 + lane;
    int32_t h = -1;
    if (i < N) {
      const int32_t s = slot[i];
      if (s <= -2 && vals[-(s + 2)] == kPending) h = -(s + 2);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, h >= 0);
    if (ballot) {
      int32_t base = 0;
      if (lane == 0) base = atomicAdd(counter, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (h >= 0) {
        const int64_t pos = nf + base + __popc(ballot & ((1u << lane) - 1u));
        if (pos < out_cap) {
          vals[h] = (int32_t)pos;
          out_unique[pos] = keys[h];
        } else {
          vals[h] = -1;  // dropped: output capacity exceeded
          atomicExch(overflow, 1);
        }
      }
    }
  }
  grid.sync();
  // ---- phase 2: relabel the neighbour lists, publish the count, pad the tail of the node list
  const int64_t total = min((int64_t)nf + (int64_t)(*counter), out_cap);
  for (int64_t j = tid; j < first_cap * fanout; j += nthreads) {
    int32_t loc = -1;
    if (j < nn) {
      const int32_t s = slot[nf + j];
      if (s != -1) loc = vals[s >= 0 ? s : -(s + 2)];
    }
    out_local[j] = loc;
  }
  for (int64_t p = total + tid; p < out_cap; p += nthreads) out_unique[p] = -1;
  if (tid == 0) *out_count = total;
  grid.sync();
  // ---- phase 3: leave the table clean for the next call
  for (int64_t i = tid; i < N; i += nthreads) {
    const int32_t s = slot[i];
    if (s != -1) {
      const int32_t h = s >= 0 ? s : -(s + 2);
      keys[h] = -1;
      vals[h] = kPending;
    }
  }
  if (tid == 0) *counter = 0;
}

}  // namespace dglb

using namespace dglb;

// keys/vals: int32[hcap], hcap a power of two >= 2*(first_cap*(1+fanout)); keys must be filled with -1 and vals
// with -2 before the first use - the kernel restores that state itself. slot: int32[first_cap*(1+fanout)].
// counter: int32[1] zero on first use. out_unique: int32[out_cap]; out_local: int32[first_cap*fanout]; out_count: int64[1].
DGLB_API int dglb200_frontier_compact(const void* first, const void* n_first_dev, int64_t first_cap, const void* ids,
                                      int fanout, void* keys, void* vals, int64_t hcap, void* slot, void* out_unique,
                                      int64_t out_cap, void* out_local, void* out_count, void* counter, void* overflow,
                                      void* stream) {
  DGLB_TRY
  const int64_t N = first_cap * (1 + (int64_t)fanout);
  DGLB_CHECK((hcap & (hcap - 1)) == 0 && hcap >= 2 * N && hcap <= (1ll << 31), "hash capacity must be a power of two >= 2N");
  if (first_cap == 0) return 0;
  static int max_blocks = 0;
  if (!max_blocks) {
    int dev = 0, sms = 0, per_sm = 0;
    DGLB_CUDA_CHECK(cudaGetDevice(&dev));
    DGLB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DGLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frontier_compact_kernel, 256, 0));
    max_blocks = sms * (per_sm < 4 ? per_sm : 4);
  }
  int64_t blocks = (N + 255) / 256;
  if (blocks > max_blocks) blocks = max_blocks;
  uint32_t mask = (uint32_t)(hcap - 1);
  void* args[] = {(void*)&first, (void*)&n_first_dev, (void*)&first_cap, (void*)&ids, (void*)&fanout, (void*)&keys,
                  (void*)&vals, (void*)&mask, (void*)&slot, (void*)&out_unique, (void*)&out_cap, (void*)&out_local,
                  (void*)&out_count, (void*)&counter, (void*)&overflow};
  DGLB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)frontier_compact_kernel, dim3((unsigned)blocks), dim3(256), args, 0,
                                              (cudaStream_t)stream));
  DGLB_CATCH
}

// ---------------------------------------------------------------------------------------------------------------------
// ELL block -> CSC arrays: out[indptr[d] + k] = nbr[d, k] for k < cnt[d]. indptr is the exclusive scan of cnt (computed
// by the caller); optional second payload (edge ids, any integer width through a byte copy of `esize` bytes).
namespace dglb {
template <typename I>
__global__ void __launch_bounds__(256) ell_to_csc_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ eid,
                                                         const int32_t* __restrict__ cnt, const int64_t* __restrict__ indptr,
                                                         int64_t n, int fanout, I* __restrict__ out_idx, I* __restrict__ out_eid,
                                                         const I* __restrict__ eid_map) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * fanout; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = i / fanout;
    const int k = (int)(i - d * fanout);
    if (k < cnt[d]) {
      const int64_t p = indptr[d] + k;
      out_idx[p] = (I)nbr[i];
      if (out_eid) {
        const int64_t e = eid[i];
        out_eid[p] = eid_map ? eid_map[e] : (I)e;
      }
    }
  }
}
}  // namespace dglb

DGLB_API int dglb200_ell_to_csc(int idtype, const void* nbr, const void* eid, const void* cnt, const void* indptr, int64_t n,
                                int fanout, void* out_idx, void* out_eid, const void* eid_map, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  const unsigned grid = (unsigned)std::min<int64_t>((n * fanout + 255) / 256, 148 * 16);
  DGLB_ID_SWITCH(idtype, I, {
    ell_to_csc_kernel<I><<<grid, 256, 0, (cudaStream_t)stream>>>((const int32_t*)nbr, (const int32_t*)eid, (const int32_t*)cnt,
                                                                 (const int64_t*)indptr, n, fanout, (I*)out_idx, (I*)out_eid,
                                                                 (const I*)eid_map);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
