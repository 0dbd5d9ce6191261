// Graph-algorithm kernels (sm_100a): random walks (metapath / restart / edge-weighted), node2vec walks, LABOR-0 layer
// sampling, global uniform negative sampling, farthest-point sampling, brute-force k-NN, neighbour matching
// (edge coarsening), index-space partition ops (owner / permutation by owner), CSR<->COO structure conversions.
//
// Reference behaviour: src/graph/sampling/randomwalks/randomwalk_gpu.cu:40-98 (thread per walk, cuRAND),
// node2vec_randomwalk.cc, src/array/cuda/labor_sampling.cu, negative_sampling.cu:26, src/geometry/cuda/
// geometry_op_impl.cu:31 (fps), edge_coarsening_impl.cu:29-81, src/graph/transform/cuda/knn.cu:261-766,
// src/partition/cuda/partition_op.cu:33-218, src/array/cuda/{coo2csr,csr2coo}.cu, graphbolt/src/cuda/expand_indptr.cu.
// Design differences: all randomness is counter-based (seed, stream, counter) so results do not depend on the launch
// shape; sampling kernels write device-side counts (no host synchronisation inside); walks keep up to 16 relations'
// CSR pointers in a kernel-parameter table so a whole metapath walk is ONE launch.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace dglb {

constexpr int kMaxWalkRels = 16;

template <typename I>
struct WalkRels {
  const I* indptr[kMaxWalkRels];
  const I* indices[kMaxWalkRels];
  const I* eids[kMaxWalkRels];   // may be null: edge id = position
  const float* prob[kMaxWalkRels];  // may be null: uniform
};

// traces: [n, L+1], eids_out: [n, L] (may be null). step_rel[i] = relation walked at step i. restart: [L] or null.
template <typename I>
__global__ void random_walk_kernel(WalkRels<I> rels, const int32_t* __restrict__ step_rel,
                                   const float* __restrict__ restart, const I* __restrict__ seeds, int64_t n, int L,
                                   uint64_t seed, I* __restrict__ traces, I* __restrict__ eids_out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  I cur = seeds[w];
  traces[w * (L + 1)] = cur;
  for (int s = 0; s < L; ++s) {
    I nxt = -1, eo = -1;
    if (cur >= 0) {
      bool stop = false;
      if (restart) stop = rng_uniform(seed, (uint64_t)w, 2 * (uint64_t)s + 1) < restart[s];
      if (!stop) {
        const int r = step_rel[s];
        const I* ip = rels.indptr[r];
        const int64_t b = ip[cur], e = ip[cur + 1];
        const int64_t deg = e - b;
        if (deg > 0) {
          int64_t p = -1;
          const float* pr = rels.prob[r];
          const I* ei = rels.eids[r];
          if (!pr) {
            p = b + (int64_t)(((rng_u64(seed, (uint64_t)w, 2 * (uint64_t)s) >> 11) * (unsigned __int128)deg) >> 53);
          } else {  // inverse CDF over the row (two passes; rows are short compared to the walk count)
            double tot = 0.0;
            for (int64_t q = b; q < e; ++q) tot += (double)pr[ei ? ei[q] : q];
            if (tot > 0.0) {
              double target = (double)rng_uniform(seed, (uint64_t)w, 2 * (uint64_t)s) * tot, acc = 0.0;
              for (int64_t q = b; q < e; ++q) {
                const double ww = (double)pr[ei ? ei[q] : q];
                acc += ww;
                if (ww > 0.0) p = q;  // last positive edge is the fallback for rounding at the top end
                if (acc > target && ww > 0.0) break;
              }
            }
          }
          if (p >= 0) {
            nxt = rels.indices[r][p];
            eo = ei ? ei[p] : (I)p;
          }
        }
      }
    }
    cur = nxt;
    traces[w * (L + 1) + s + 1] = cur;
    if (eids_out) eids_out[w * L + s] = eo;
  }
}

// node2vec second-order walk by rejection sampling (reference node2vec_randomwalk.cc): candidate x ~ U(N(cur));
// accept with prob 1/p if x == prev, 1 if x in N(prev), 1/q otherwise (all divided by the maximum of the three).
template <typename I>
__global__ void node2vec_walk_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                     const I* __restrict__ eids, const I* __restrict__ seeds, int64_t n, int L,
                                     float inv_p, float inv_q, uint64_t seed, I* __restrict__ traces,
                                     I* __restrict__ eids_out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const float maxw = fmaxf(fmaxf(inv_p, 1.f), inv_q);
  I prev = -1, cur = seeds[w];
  traces[w * (L + 1)] = cur;
  uint64_t ctr = 0;
  for (int s = 0; s < L; ++s) {
    I nxt = -1, eo = -1;
    if (cur >= 0) {
      const int64_t b = indptr[cur], deg = indptr[cur + 1] - b;
      if (deg > 0) {
        for (int tries = 0; tries < 1 << 16; ++tries) {
          const int64_t p = b + (int64_t)(((rng_u64(seed, (uint64_t)w, ctr++) >> 11) * (unsigned __int128)deg) >> 53);
          const I x = indices[p];
          float acc = inv_q;
          if (prev < 0) acc = maxw;  // first step is a first-order step
          else if (x == prev) acc = inv_p;
          else {
            const int64_t pb = indptr[prev], pe = indptr[prev + 1];
            for (int64_t q = pb; q < pe; ++q)
              if (indices[q] == x) { acc = 1.f; break; }
          }
          if (rng_uniform(seed, (uint64_t)w, ctr++) * maxw <= acc) {
            nxt = x;
            eo = eids ? eids[p] : (I)p;
            break;
          }
        }
      }
    }
    prev = cur;
    cur = nxt;
    traces[w * (L + 1) + s + 1] = cur;
    if (eids_out) eids_out[w * L + s] = eo;
  }
}

// ---------------------------------------------------------------------------------------------- LABOR-0
// Edge (t -> s) of seed s is kept iff r_t <= fanout / deg(s), r_t = U(seed, t) shared by every seed (and every layer
// using the same seed), so overlapping neighbourhoods pick the same vertices. Pass 1 counts, pass 2 (after an
// exclusive scan of the counts) writes (src, dst position, edge id). One sub-warp of 8 lanes per seed.
template <typename I, bool FILL>
__global__ void labor0_kernel(const I* __restrict__ indptr, const I* __restrict__ indices, const I* __restrict__ eids,
                              const I* __restrict__ seeds, int64_t n_seeds, int fanout, uint64_t seed,
                              int64_t* __restrict__ counts, const int64_t* __restrict__ offsets, I* __restrict__ out_src,
                              I* __restrict__ out_seedpos, I* __restrict__ out_eid) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 7;
  if (i >= n_seeds) return;
  const I s = seeds[i];
  const int64_t b = indptr[s], e = indptr[s + 1], deg = e - b;
  const float thr = fanout < 0 || deg <= fanout ? 2.f : (float)fanout / (float)deg;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
  int64_t base = FILL ? offsets[i] : 0, cnt = 0;
  for (int64_t q0 = b; q0 < e; q0 += 8) {
    const int64_t q = q0 + lane;
    bool keep = false;
    I t = 0;
    if (q < e) {
      t = indices[q];
      keep = rng_uniform(seed, 0x4c41424full, (uint64_t)t) <= thr;
    }
    const unsigned m = __ballot_sync(gmask, keep) & gmask;
    if (FILL && keep) {
      const int64_t pos = base + cnt + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
      out_src[pos] = t;
      out_seedpos[pos] = (I)i;
      out_eid[pos] = eids ? eids[q] : (I)q;
    }
    cnt += __popc(m);
  }
  if (!FILL && lane == 0) counts[i] = cnt;
}

// ---------------------------------------------------------------------------------------------- negative sampling
// One trial per thread: (u, v) uniform; rejected if it is an edge (linear / binary search in the CSR row of u) or a
// self loop. ok[i] = 1 marks accepted trials; de-duplication / truncation are done by the caller.
template <typename I>
__global__ void neg_sample_kernel(const I* __restrict__ indptr, const I* __restrict__ indices, int sorted_rows,
                                  int64_t n_src, int64_t n_dst, int exclude_self, int64_t trials, uint64_t seed,
                                  I* __restrict__ out_u, I* __restrict__ out_v, uint8_t* __restrict__ ok) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= trials) return;
  const int64_t u = (int64_t)(((rng_u64(seed, 1, (uint64_t)i) >> 11) * (unsigned __int128)n_src) >> 53);
  const int64_t v = (int64_t)(((rng_u64(seed, 2, (uint64_t)i) >> 11) * (unsigned __int128)n_dst) >> 53);
  bool good = !(exclude_self && u == v);
  if (good) {
    int64_t b = indptr[u], e = indptr[u + 1];
    if (sorted_rows) {
      while (b < e) {
        const int64_t mid = (b + e) >> 1;
        const int64_t x = indices[mid];
        if (x == v) { good = false; break; }
        if (x < v) b = mid + 1; else e = mid;
      }
    } else {
      for (int64_t q = b; q < e; ++q)
        if (indices[q] == v) { good = false; break; }
    }
  }
  out_u[i] = (I)u;
  out_v[i] = (I)v;
  ok[i] = good ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- farthest point sampling
// One CTA per point cloud; dist[] lives in global memory (L2-resident for typical clouds), every iteration updates the
// distances to the newest sample and reduces the arg-max with warp shuffles + one smem pass.
template <typename T>
__global__ void __launch_bounds__(1024) fps_kernel(const T* __restrict__ pos, int64_t N, int C, int npoints,
                                                   const int64_t* __restrict__ start, float* __restrict__ dist,
                                                   int64_t* __restrict__ out) {
  const int b = blockIdx.x;
  const T* P = pos + (int64_t)b * N * C;
  float* D = dist + (int64_t)b * N;
  __shared__ float s_val[32];
  __shared__ int64_t s_idx[32];
  __shared__ int64_t s_cur;
  __shared__ float s_c[16];
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) D[i] = 3.4e38f;
  if (threadIdx.x == 0) s_cur = start[b];
  __syncthreads();
  for (int it = 0; it < npoints; ++it) {
    const int64_t cur = s_cur;
    if (threadIdx.x == 0) out[(int64_t)b * npoints + it] = cur;
    if (threadIdx.x < C) s_c[threadIdx.x] = (float)P[cur * C + threadIdx.x];
    __syncthreads();
    float best = -1.f;
    int64_t bi = 0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
      float d = 0.f;
      for (int c = 0; c < C; ++c) {
        const float x = (float)P[i * C + c] - s_c[c];
        d = fmaf(x, x, d);
      }
      d = fminf(d, D[i]);
      D[i] = d;
      if (d > best) { best = d; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_down_sync(0xffffffffu, best, o);
      const int64_t oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int nw = (blockDim.x + 31) >> 5;
      best = threadIdx.x < nw ? s_val[threadIdx.x] : -2.f;
      bi = threadIdx.x < nw ? s_idx[threadIdx.x] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, best, o);
        const int64_t oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (threadIdx.x == 0) s_cur = bi;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- brute-force k-NN
// One thread per query; the points of the query's segment stream through shared memory in tiles of 256; each thread
// keeps its k best in a register/local max-list (k <= 64). Output: neighbour ids [n_query, k] (global point ids,
// -1 padded when the segment has fewer than k points), sorted by distance.
constexpr int kKnnTile = 256;
constexpr int kKnnMaxK = 64;
constexpr int kKnnMaxC = 16;

template <typename T>
__global__ void __launch_bounds__(kKnnTile) knn_brute_kernel(const T* __restrict__ data, const int64_t* __restrict__ data_off,
                                                              const T* __restrict__ query, const int64_t* __restrict__ query_off,
                                                              int C, int k, int64_t* __restrict__ out_idx,
                                                              float* __restrict__ out_dist) {
  // grid: (tiles of queries within the largest segment, segments)
  const int seg = blockIdx.y;
  const int64_t qb = query_off[seg], qe = query_off[seg + 1];
  const int64_t db = data_off[seg], de = data_off[seg + 1];
  const int64_t qi = qb + (int64_t)blockIdx.x * kKnnTile + threadIdx.x;
  if ((int64_t)blockIdx.x * kKnnTile >= qe - qb) return;
  const bool live = qi < qe;
  extern __shared__ float s_pts[];  // [kKnnTile, C]
  float q[kKnnMaxC];
  for (int c = 0; c < C; ++c) q[c] = live ? (float)query[qi * C + c] : 0.f;
  float bd[kKnnMaxK];
  int64_t bix[kKnnMaxK];
  for (int j = 0; j < k; ++j) { bd[j] = 3.4e38f; bix[j] = -1; }
  for (int64_t t0 = db; t0 < de; t0 += kKnnTile) {
    const int64_t cnt = min((int64_t)kKnnTile, de - t0);
    __syncthreads();
    for (int64_t x = threadIdx.x; x < cnt * C; x += kKnnTile) s_pts[x] = (float)data[t0 * C + x];
    __syncthreads();
    if (live) {
      for (int64_t j = 0; j < cnt; ++j) {
        float d = 0.f;
        for (int c = 0; c < C; ++c) {
          const float x = s_pts[j * C + c] - q[c];
          d = fmaf(x, x, d);
        }
        if (d < bd[k - 1]) {  // insertion into the sorted list
          int p = k - 1;
          while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bix[p] = bix[p - 1]; --p; }
          bd[p] = d;
          bix[p] = t0 + j;
        }
      }
    }
  }
  if (live)
    for (int j = 0; j < k; ++j) {
      out_idx[qi * k + j] = bix[j];
      if (out_dist) out_dist[qi * k + j] = bd[j];
    }
}

// ---------------------------------------------------------------------------------------------- neighbour matching
// Handshake matching rounds (edge coarsening): every unmatched vertex proposes to its heaviest unmatched neighbour
// (ties: hashed vertex priority so the result is not biased by the id order); mutual proposals are matched. The host
// loops a fixed number of rounds, leftovers become singletons.
template <typename I>
__global__ void match_propose_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                     const I* __restrict__ eids, const float* __restrict__ w, int64_t n, uint64_t seed,
                                     const I* __restrict__ match, I* __restrict__ proposal) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  I best = -1;
  if (match[v] < 0) {
    float bw = -3.4e38f;
    uint64_t bp = 0;
    for (int64_t q = indptr[v]; q < indptr[v + 1]; ++q) {
      const I u = indices[q];
      if (u == v || match[u] >= 0) continue;
      const float ww = w ? w[eids ? eids[q] : q] : 1.f;
      const uint64_t pr = mix64(seed ^ (uint64_t)(u < v ? ((uint64_t)u << 32 | (uint64_t)v) : ((uint64_t)v << 32 | (uint64_t)u)));
      if (ww > bw || (ww == bw && pr > bp)) { bw = ww; bp = pr; best = u; }
    }
  }
  proposal[v] = best;
}
template <typename I>
__global__ void match_accept_kernel(int64_t n, const I* __restrict__ proposal, I* __restrict__ match) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n || match[v] >= 0) return;
  const I u = proposal[v];
  if (u >= 0 && proposal[u] == (I)v) match[v] = v < u ? (I)v : u;
}
template <typename I>
__global__ void match_finish_kernel(int64_t n, I* __restrict__ match) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n && match[v] < 0) match[v] = (I)v;
}

// ---------------------------------------------------------------------------------------------- partition ops
// owner of each index: remainder (idx % P) or range (upper bound in the P+1 boundary array).
template <typename I>
__global__ void partition_owner_kernel(const I* __restrict__ idx, int64_t n, int P, const int64_t* __restrict__ range,
                                       int32_t* __restrict__ owner, int64_t* __restrict__ counts, I* __restrict__ local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t x = idx[i];
  int o;
  int64_t loc;
  if (!range) {
    o = (int)(x % P);
    loc = x / P;
  } else {
    int lo = 0, hi = P;  // largest o with range[o] <= x
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (range[mid] <= x) lo = mid; else hi = mid;
    }
    o = lo;
    loc = x - range[o];
  }
  owner[i] = o;
  if (local) local[i] = (I)loc;
  if (counts) atomicAdd(reinterpret_cast<unsigned long long*>(counts + o), 1ull);
}

template <typename I>
__global__ void iota_kernel(I* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (I)i;
}

// ---------------------------------------------------------------------------------------------- structure conversions
// CSR -> COO row ids ("expand indptr"): one thread per nonzero, binary search of the owning row; optionally offsets the
// row ids through node_ids (graphbolt expand_indptr semantics).
template <typename I>
__global__ void expand_indptr_kernel(const I* __restrict__ indptr, int64_t n_rows, int64_t nnz,
                                     const I* __restrict__ node_ids, I* __restrict__ rows) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int64_t lo = 0, hi = n_rows;  // largest r with indptr[r] <= e
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)indptr[mid] <= e) lo = mid; else hi = mid;
  }
  rows[e] = node_ids ? node_ids[lo] : (I)lo;
}

// row histogram for COO -> CSR (rows need not be sorted): counts[r + 1] += 1, scanned afterwards.
template <typename I>
__global__ void row_hist_kernel(const I* __restrict__ row, int64_t nnz, I* __restrict__ counts) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) {
    if (sizeof(I) == 8) atomicAdd(reinterpret_cast<unsigned long long*>(counts + row[e]), 1ull);
    else atomicAdd(reinterpret_cast<unsigned int*>(counts + row[e]), 1u);
  }
}

// ---------------------------------------------------------------------------------------------- keyed sum
// out[key[i], c] += val[i, c] for FEW distinct keys (relation ids, partition ids): every CTA privatises the whole
// [num_keys x C] table in shared memory, so the global atomics drop from n * C to gridDim * num_keys * C. This is the
// backward of `table[keys]` when millions of rows hit a handful of table rows (torch's index_add_ serialises on them).
template <typename I>
__global__ void __launch_bounds__(256) keyed_sum_kernel(const I* __restrict__ keys, int64_t n, const float* __restrict__ vals,
                                                        int C, int num_keys, float* __restrict__ out) {
  extern __shared__ float s_acc[];
  const int tab = num_keys * C;
  for (int i = threadIdx.x; i < tab; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int64_t total = n * C;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = x / C;
    const int c = (int)(x - i * C);
    atomicAdd(&s_acc[(int)keys[i] * C + c], vals[x]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tab; i += blockDim.x) {
    const float v = s_acc[i];
    if (v != 0.f) atomicAdd(out + i, v);
  }
}

static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)std::max<int64_t>(1, (n + t - 1) / t); }

}  // namespace dglb

using namespace dglb;

// rel_ptrs: host array of 4 * n_rels pointers (indptr, indices, eids|0, prob|0 per relation)
DGLB_API int dglb200_cuda_random_walk(int idtype, const void* const* rel_ptrs, int n_rels, const void* step_rel,
                                      const void* restart, const void* seeds, int64_t n, int L, int64_t seed,
                                      void* traces, void* eids_out, void* stream) {
  DGLB_TRY
  DGLB_CHECK(n_rels >= 1 && n_rels <= kMaxWalkRels, "random_walk: at most 16 distinct relations per metapath");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_ID_SWITCH(idtype, I, {
    WalkRels<I> rels;
    for (int r = 0; r < kMaxWalkRels; ++r) {
      rels.indptr[r] = r < n_rels ? (const I*)rel_ptrs[4 * r] : nullptr;
      rels.indices[r] = r < n_rels ? (const I*)rel_ptrs[4 * r + 1] : nullptr;
      rels.eids[r] = r < n_rels ? (const I*)rel_ptrs[4 * r + 2] : nullptr;
      rels.prob[r] = r < n_rels ? (const float*)rel_ptrs[4 * r + 3] : nullptr;
    }
    random_walk_kernel<I><<<blocks_for(n, 128), 128, 0, st>>>(rels, (const int32_t*)step_rel, (const float*)restart,
                                                              (const I*)seeds, n, L, (uint64_t)seed, (I*)traces, (I*)eids_out);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_node2vec_walk(int idtype, const void* indptr, const void* indices, const void* eids,
                                        const void* seeds, int64_t n, int L, double p, double q, int64_t seed,
                                        void* traces, void* eids_out, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    node2vec_walk_kernel<I><<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
        (const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, L, (float)(1.0 / p), (float)(1.0 / q),
        (uint64_t)seed, (I*)traces, (I*)eids_out);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// pass 0: counts[n_seeds] (int64). pass 1: offsets = exclusive scan of counts; fills out_src / out_seedpos / out_eid.
DGLB_API int dglb200_cuda_labor0(int idtype, int pass, const void* indptr, const void* indices, const void* eids,
                                 const void* seeds, int64_t n_seeds, int fanout, int64_t seed, void* counts,
                                 const void* offsets, void* out_src, void* out_seedpos, void* out_eid, void* stream) {
  DGLB_TRY
  if (n_seeds == 0) return 0;
  const unsigned blocks = blocks_for(n_seeds * 8, 256);
  DGLB_ID_SWITCH(idtype, I, {
    if (pass == 0)
      labor0_kernel<I, false><<<blocks, 256, 0, (cudaStream_t)stream>>>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n_seeds, fanout, (uint64_t)seed, (int64_t*)counts, nullptr, nullptr, nullptr, nullptr);
    else
      labor0_kernel<I, true><<<blocks, 256, 0, (cudaStream_t)stream>>>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n_seeds, fanout, (uint64_t)seed, nullptr, (const int64_t*)offsets, (I*)out_src, (I*)out_seedpos, (I*)out_eid);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_neg_sample(int idtype, const void* indptr, const void* indices, int sorted_rows, int64_t n_src,
                                     int64_t n_dst, int exclude_self, int64_t trials, int64_t seed, void* out_u,
                                     void* out_v, void* ok, void* stream) {
  DGLB_TRY
  if (trials == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    neg_sample_kernel<I><<<blocks_for(trials, 256), 256, 0, (cudaStream_t)stream>>>(
        (const I*)indptr, (const I*)indices, sorted_rows, n_src, n_dst, exclude_self, trials, (uint64_t)seed, (I*)out_u,
        (I*)out_v, (uint8_t*)ok);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// pos: [B, N, C] (fp32 / fp16 / bf16), start: [B] int64, dist: [B, N] fp32 scratch, out: [B, npoints] int64
DGLB_API int dglb200_cuda_fps(int dtype, const void* pos, int64_t B, int64_t N, int C, int npoints, const void* start,
                              void* dist, void* out, void* stream) {
  DGLB_TRY
  DGLB_CHECK(C >= 1 && C <= 16, "fps: coordinate dimension must be in [1, 16]");
  if (B == 0 || npoints == 0) return 0;
  const int threads = N >= 1024 ? 1024 : (int)std::max<int64_t>(32, (N + 31) / 32 * 32);
  DGLB_FLOAT_SWITCH(dtype, T, {
    fps_kernel<T><<<(unsigned)B, threads, 0, (cudaStream_t)stream>>>((const T*)pos, N, C, npoints, (const int64_t*)start, (float*)dist, (int64_t*)out);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// data: [n_data, C], query: [n_query, C]; *_off: [n_seg + 1] int64 segment offsets; max_q = longest query segment
DGLB_API int dglb200_cuda_knn_brute(int dtype, const void* data, const void* data_off, const void* query,
                                    const void* query_off, int n_seg, int64_t max_q, int C, int k, void* out_idx,
                                    void* out_dist, void* stream) {
  DGLB_TRY
  DGLB_CHECK(C >= 1 && C <= kKnnMaxC, "knn: feature dimension must be in [1, 16] for the brute-force kernel");
  DGLB_CHECK(k >= 1 && k <= kKnnMaxK, "knn: k must be in [1, 64] for the brute-force kernel");
  if (n_seg == 0 || max_q == 0) return 0;
  const dim3 grid((unsigned)((max_q + kKnnTile - 1) / kKnnTile), (unsigned)n_seg);
  DGLB_FLOAT_SWITCH(dtype, T, {
    knn_brute_kernel<T><<<grid, kKnnTile, kKnnTile * C * sizeof(float), (cudaStream_t)stream>>>(
        (const T*)data, (const int64_t*)data_off, (const T*)query, (const int64_t*)query_off, C, k, (int64_t*)out_idx, (float*)out_dist);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// match: [n] initialised to -1 by the caller; proposal: [n] scratch
DGLB_API int dglb200_cuda_neighbor_matching(int idtype, const void* indptr, const void* indices, const void* eids,
                                            const void* w, int64_t n, int64_t seed, int rounds, void* match,
                                            void* proposal, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = blocks_for(n, 256);
  DGLB_ID_SWITCH(idtype, I, {
    for (int r = 0; r < rounds; ++r) {
      match_propose_kernel<I><<<blocks, 256, 0, st>>>((const I*)indptr, (const I*)indices, (const I*)eids, (const float*)w, n,
                                                     (uint64_t)seed + (uint64_t)r * 0x9E3779B97F4A7C15ull, (const I*)match, (I*)proposal);
      match_accept_kernel<I><<<blocks, 256, 0, st>>>(n, (const I*)proposal, (I*)match);
    }
    match_finish_kernel<I><<<blocks, 256, 0, st>>>(n, (I*)match);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// owner[i] (int32), optional counts[P] (int64, zeroed by the caller) and local index; range = [P + 1] int64 or null
DGLB_API int dglb200_cuda_partition_owner(int idtype, const void* idx, int64_t n, int P, const void* range, void* owner,
                                          void* counts, void* local, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    partition_owner_kernel<I><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const I*)idx, n, P, (const int64_t*)range, (int32_t*)owner, (int64_t*)counts, (I*)local);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// stable permutation that groups positions by owner: perm (int64) such that owner[perm] is sorted. temp: scratch of
// dglb200_cuda_partition_perm_temp_bytes(n) bytes; keys_out: [n] int32 scratch.
DGLB_API int64_t dglb200_cuda_partition_perm_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int64_t*)nullptr,
                                  (int64_t*)nullptr, (int)n, 0, 32);
  return (int64_t)bytes;
}
DGLB_API int dglb200_cuda_partition_perm(const void* owner, int64_t n, int P, void* iota, void* keys_out, void* perm,
                                         void* temp, int64_t temp_bytes, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  iota_kernel<int64_t><<<blocks_for(n, 256), 256, 0, st>>>((int64_t*)iota, n);
  int bits = 1;
  while ((1 << bits) < P) ++bits;
  size_t tb = (size_t)temp_bytes;
  DGLB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(temp, tb, (const int32_t*)owner, (int32_t*)keys_out, (const int64_t*)iota,
                                                  (int64_t*)perm, (int)n, 0, bits, st));
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_expand_indptr(int idtype, const void* indptr, int64_t n_rows, int64_t nnz, const void* node_ids,
                                        void* rows, void* stream) {
  DGLB_TRY
  if (nnz == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    expand_indptr_kernel<I><<<blocks_for(nnz, 256), 256, 0, (cudaStream_t)stream>>>((const I*)indptr, n_rows, nnz, (const I*)node_ids, (I*)rows);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// indptr: [n_rows + 1], zeroed by the caller; after the call indptr[r + 1] holds the row count (the caller scans it)
DGLB_API int dglb200_cuda_row_hist(int idtype, const void* row, int64_t nnz, void* indptr_plus1, void* stream) {
  DGLB_TRY
  if (nnz == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    row_hist_kernel<I><<<blocks_for(nnz, 256), 256, 0, (cudaStream_t)stream>>>((const I*)row, nnz, (I*)indptr_plus1);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// out: [num_keys, C] fp32, accumulated into (caller zeroes it); num_keys * C <= 8192
DGLB_API int dglb200_cuda_keyed_sum(int idtype, const void* keys, int64_t n, const void* vals, int C, int num_keys,
                                    void* out, void* stream) {
  DGLB_TRY
  DGLB_CHECK((int64_t)num_keys * C <= 8192, "keyed_sum: table must fit 32 KB of shared memory");
  if (n == 0) return 0;
  int64_t blocks = (n * C + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  DGLB_ID_SWITCH(idtype, I, {
    keyed_sum_kernel<I><<<(unsigned)blocks, 256, (size_t)num_keys * C * sizeof(float), (cudaStream_t)stream>>>(
        (const I*)keys, n, (const float*)vals, C, num_keys, (float*)out);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
