// Visit-frequency top-k for random-walk neighbourhoods (PinSAGE): for every seed, count how often each node was
// visited by that seed's walks and keep the K most visited (count descending, node id ascending on ties).
//
// Reference: src/graph/sampling/randomwalks/frequency_hashmap.cu:36-167 + :240-420 - a global-memory hash map over all
// (node, seed) pairs, a device-wide radix sort of the counts per seed and a compaction pass.
// Design for sm_100a: a seed's visits are few (walks x traversals, typically 10^2..10^3), so the whole problem of one
// seed fits in shared memory. One CTA per seed builds an open-addressing table in shared memory with shared-memory
// atomics, folds (count, id) of every slot into one 64-bit word in place and extracts the top K with K block-wide
// arg-max rounds over those words (one 8-byte shared load per slot, one shared atomicMax per thread and two barriers per
// round; the winner's owner clears its slot). Global traffic is exactly one read of the visit row and one write of K
// results per seed; there is no device-wide sort, no scratch allocation and no host synchronisation, so the call is
// CUDA-graph capturable.
#include "common.cuh"

namespace dglb {

// visits: [n_seeds, M] (negative = no visit). out_node / out_cnt: [n_seeds, K], empty slots: node = -1, count = 0.
// Shared memory: H 64-bit keys + H 32-bit counts, H = power of two >= 2 * M.
template <typename I>
__global__ void __launch_bounds__(128) visit_topk_kernel(const I* __restrict__ visits, int64_t n_seeds, int M, int K, int H,
                                                          I* __restrict__ out_node, int* __restrict__ out_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  long long* keys = reinterpret_cast<long long*>(smem_raw);
  int* cnts = reinterpret_cast<int*>(keys + H);
  __shared__ unsigned long long best[2];  // double-buffered by round parity: two barriers per round suffice
  const unsigned mask = (unsigned)H - 1u;
  for (int64_t seed = blockIdx.x; seed < n_seeds; seed += gridDim.x) {
    for (int s = threadIdx.x; s < H; s += blockDim.x) {
      keys[s] = -1;
      cnts[s] = 0;
    }
    __syncthreads();
    const I* row = visits + seed * (int64_t)M;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      const long long v = (long long)row[j];
      if (v < 0) continue;
      unsigned slot = (unsigned)(mix64((uint64_t)v) & mask);
      while (true) {
        const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&keys[slot]),
                                                    (unsigned long long)(-1ll), (unsigned long long)v);
        if (prev == -1 || prev == v) {
          atomicAdd(&cnts[slot], 1);
          break;
        }
        slot = (slot + 1) & mask;
      }
    }
    __syncthreads();
    // fold (count, key) of every slot into one comparable word, in place: count (24 bits) | inverted node id (40 bits).
    // A max over the words picks the highest count and, on ties, the smallest id; empty slots become 0.
    unsigned long long* packed = reinterpret_cast<unsigned long long*>(keys);
    for (int s = threadIdx.x; s < H; s += blockDim.x) {
      const int c = cnts[s];
      packed[s] = c > 0 ? (((unsigned long long)c << 40) | (0xFFFFFFFFFFull - (unsigned long long)keys[s])) : 0ull;
    }
    if (threadIdx.x == 0) best[0] = 0ull;
    __syncthreads();
    for (int k = 0; k < K; ++k) {
      unsigned long long mine = 0ull;
      int mine_slot = -1;
      for (int s = threadIdx.x; s < H; s += blockDim.x) {
        const unsigned long long p = packed[s];
        if (p > mine) {
          mine = p;
          mine_slot = s;
        }
      }
      if (mine) atomicMax(&best[k & 1], mine);
      __syncthreads();
      const unsigned long long b = best[k & 1];
      if (b == 0ull) {
        for (int r = k + (int)threadIdx.x; r < K; r += blockDim.x) {
          out_node[seed * K + r] = I(-1);
          out_cnt[seed * K + r] = 0;
        }
        break;
      }
      if (mine == b) {  // node ids are distinct inside a table, so exactly one thread owns the winner
        packed[mine_slot] = 0ull;
        out_node[seed * K + k] = I((long long)(0xFFFFFFFFFFull - (b & 0xFFFFFFFFFFull)));
        out_cnt[seed * K + k] = (int)(b >> 40);
      }
      if (threadIdx.x == 0) best[(k + 1) & 1] = 0ull;  // last read in round k-1, next written after the barrier below
      __syncthreads();  // slot clearing and the reset are visible before the next round's scan
    }
    __syncthreads();
  }
}

}  // namespace dglb

using namespace dglb;

// Returns the shared-memory bytes a launch needs through *smem_bytes (0 = shape not supported: fall back).
DGLB_API int dglb200_cuda_visit_topk(int idtype, const void* visits, int64_t n_seeds, int M, int K, void* out_node,
                                     void* out_cnt, void* stream) {
  DGLB_TRY
  if (n_seeds == 0 || K == 0) return 0;
  DGLB_CHECK(M > 0 && M < (1 << 24), "visit_topk: visits per seed must be in [1, 2^24)");
  int H = 64;
  while (H < 2 * M) H <<= 1;
  const size_t smem = (size_t)H * (sizeof(long long) + sizeof(int));
  DGLB_CHECK(smem <= 200 * 1024, "visit_topk: more than ~8k visits per seed do not fit the shared-memory table");
  const unsigned grid = (unsigned)std::min<int64_t>(n_seeds, 148 * 16);
  DGLB_ID_SWITCH(idtype, I, {
    auto kern = visit_topk_kernel<I>;
    if (smem > 48 * 1024) DGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, (cudaStream_t)stream>>>((const I*)visits, n_seeds, M, K, H, (I*)out_node, (int*)out_cnt);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
