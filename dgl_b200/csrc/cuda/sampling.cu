// Neighbor sampling kernels for sm_100a (graph on device or host-pinned/UVA).
//  * sample_ell_kernel : the hot path — uniform, without replacement, fanout <= 64; output is
//    a dense ELL block [n_rows, fanout] (pad -1) + per-row counts, sized by an upper bound so
//    NO host synchronisation is needed and multi-layer sampling chains on device-side counts.
//    One thread per seed runs Floyd's subset algorithm (k RNG draws, O(k^2) dedup in registers)
//    instead of the reference's block-per-row reservoir with atomics
//    (/root/reference/src/array/cuda/rowwise_sampling.cu:43-229, graphbolt neighbor_sampler.cu:76-145).
//  * sample_count / sample_fill : exact-size COO output for the DGL-compatible API, uniform or
//    weighted (exponential-race keys == A-Res), with or without replacement.
// RNG is counter based: (seed, global row, draw) -> value, so results do not depend on the launch
// shape and are reproducible.
#include "common.cuh"

namespace dglb {

constexpr int kMaxFloyd = 64;
constexpr int kTotalSlots = 16;  // out_total is int64[16]: the sum of the slots is the running edge count

template <typename I>
__global__ void __launch_bounds__(128) sample_ell_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                         const I* __restrict__ seeds, int64_t n_cap,
                                                         const int64_t* __restrict__ n_dev, int fanout, uint64_t seed,
                                                         const int64_t* __restrict__ seed_dev, I* __restrict__ out_nbr, I* __restrict__ out_eid,
                                                         int32_t* __restrict__ out_cnt, int64_t* __restrict__ out_total,
                                                         int pad_rows) {
  const int64_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = (pad_rows & 1) ? (i < n_cap) : (i < n);
  const uint64_t pos_key = (pad_rows & 2) ? 0ull : (uint64_t)i;
  if (seed_dev) seed += (uint64_t)(*seed_dev) * 0xD6E8FEB86659FD93ull;  // per-step stream under graph replay
  I* nb = out_nbr + i * fanout;
  I* ne = out_eid ? out_eid + i * fanout : nullptr;
  int cnt = 0;
  if (active && i < n) {
    const int64_t v = seeds[i];
    if (v >= 0) {
      const int64_t s = indptr[v], deg = indptr[v + 1] - s;
      if (deg <= fanout) {
        cnt = (int)deg;
        for (int k = 0; k < cnt; ++k) {
          nb[k] = indices[s + k];
          if (ne) ne[k] = (I)(s + k);
        }
      } else {
        cnt = fanout;
        uint32_t chosen[kMaxFloyd];
        int m = 0;
        for (int64_t j = deg - fanout; j < deg; ++j, ++m) {
          uint32_t t = rng_below(seed, (uint64_t)v * 0x9E37ull + pos_key, (uint64_t)m, (uint32_t)(j + 1));
          bool dup = false;
          for (int q = 0; q < m; ++q) dup |= (chosen[q] == t);
          chosen[m] = dup ? (uint32_t)j : t;
        }
        for (int k = 0; k < fanout; ++k) {
          nb[k] = indices[s + chosen[k]];
          if (ne) ne[k] = (I)(s + chosen[k]);
        }
      }
    }
  }
  if (active) {
    for (int k = cnt; k < fanout; ++k) {
      nb[k] = (I)-1;
      if (ne) ne[k] = (I)-1;
    }
    out_cnt[i] = cnt;
  }
  if (out_total) {  // running total of sampled edges (warp-aggregated atomic, spread over kTotalSlots addresses)
    int wsum = cnt;
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if ((threadIdx.x & 31) == 0 && wsum > 0)
      atomicAdd((unsigned long long*)out_total + (blockIdx.x & (kTotalSlots - 1)), (unsigned long long)wsum);
  }
}

// Sub-warp variant (fanout <= 32): G lanes cooperate on one seed. Every lane draws its own Floyd candidate, the
// sequential "already chosen?" rule is resolved with fanout-1 register shuffles (lane q's pick is final after step
// q-1), and the neighbour ids are fetched / written by G lanes at once - no local-memory scratch, and the three
// dependent memory round trips (seed -> indptr -> indices) are the whole critical path. Same RNG stream and the same
// picks as the thread-per-seed kernel.
template <typename I, int G>
__global__ void __launch_bounds__(256) sample_ell_group_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                               const I* __restrict__ seeds, int64_t n_cap,
                                                               const int64_t* __restrict__ n_dev, int fanout, uint64_t seed,
                                                               const int64_t* __restrict__ seed_dev, I* __restrict__ out_nbr,
                                                               I* __restrict__ out_eid, int32_t* __restrict__ out_cnt,
                                                               int64_t* __restrict__ out_total, int pad_rows) {
  const int64_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = gtid / G;
  const int k = (int)(gtid % G);
  const bool active = (pad_rows & 1) ? (i < n_cap) : (i < n);
  const uint64_t pos_key = (pad_rows & 2) ? 0ull : (uint64_t)i;  // bit 1: picks depend on the node only, not on its row
  if (seed_dev) seed += (uint64_t)(*seed_dev) * 0xD6E8FEB86659FD93ull;  // per-step stream under graph replay
  int cnt = 0;
  int64_t pick = -1;  // CSR position of this lane's neighbour
  int64_t v = -1, s = 0, deg = 0;
  if (active && i < n) {
    v = seeds[i];
    if (v >= 0) {
      s = indptr[v];
      deg = indptr[v + 1] - s;
    }
  }
  const bool floyd = deg > fanout;  // uniform inside the group
  uint32_t t = 0, val = 0, j = 0;
  if (floyd && k < fanout) {
    j = (uint32_t)(deg - fanout + k);
    t = rng_below(seed, (uint64_t)v * 0x9E37ull + pos_key, (uint64_t)k, j + 1);
    val = t;
  }
  bool dup = false;
  for (int q = 0; q + 1 < fanout; ++q) {  // warp-uniform trip count; groups without a Floyd row just ride along
    const uint32_t cq = __shfl_sync(0xffffffffu, val, q, G);
    if (k > q && t == cq) dup = true;
    if (k == q + 1) val = dup ? j : t;
  }
  if (deg > 0) {
    if (floyd) {
      cnt = fanout;
      if (k < fanout) pick = s + val;
    } else {
      cnt = (int)deg;
      if (k < cnt) pick = s + k;
    }
  }
  if (active && k < fanout) {
    const int64_t o = i * fanout + k;
    out_nbr[o] = pick >= 0 ? indices[pick] : (I)-1;
    if (out_eid) out_eid[o] = (I)pick;
  }
  if (active && k == 0) out_cnt[i] = cnt;
  if (out_total) {  // running total of sampled edges: one atomic per block, spread over kTotalSlots addresses
    __shared__ int s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    int wsum = k == 0 ? cnt : 0;
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if ((threadIdx.x & 31) == 0 && wsum > 0) atomicAdd(&s_sum, wsum);
    __syncthreads();
    if (threadIdx.x == 0 && s_sum > 0)
      atomicAdd((unsigned long long*)out_total + (blockIdx.x & (kTotalSlots - 1)), (unsigned long long)s_sum);
  }
}

template <typename I>
__global__ void __launch_bounds__(256) sample_count_kernel(const I* __restrict__ indptr, const I* __restrict__ eids,
                                                           const I* __restrict__ seeds, int64_t n, int64_t fanout,
                                                           int replace, const float* __restrict__ prob,
                                                           int64_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = seeds[i];
  const int64_t s = indptr[v], deg = indptr[v + 1] - s;
  int64_t c;
  if (deg == 0) {
    c = 0;
  } else if (prob) {
    int64_t nz = 0;
    for (int64_t k = 0; k < deg; ++k) nz += prob[eids ? (int64_t)eids[s + k] : s + k] > 0.f;
    c = nz == 0 ? 0 : (fanout < 0 ? nz : (replace ? fanout : min(fanout, nz)));
  } else {
    c = fanout < 0 ? deg : (replace ? fanout : min(fanout, deg));
  }
  counts[i] = c;
}

template <typename I>
__global__ void __launch_bounds__(128) sample_fill_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                          const I* __restrict__ eids, const I* __restrict__ seeds,
                                                          int64_t n, int64_t fanout, int replace,
                                                          const float* __restrict__ prob,
                                                          const int64_t* __restrict__ offsets, uint64_t seed,
                                                          I* __restrict__ out_src, I* __restrict__ out_dst,
                                                          I* __restrict__ out_eid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = seeds[i];
  const int64_t s = indptr[v], deg = indptr[v + 1] - s;
  const int64_t o = offsets[i], cnt = offsets[i + 1] - o;
  if (cnt == 0) return;
  const uint64_t stream = (uint64_t)v * 0x9E37ull + (uint64_t)i;
  auto emit = [&](int64_t slot, int64_t k) {
    out_src[o + slot] = indices[s + k];
    out_dst[o + slot] = (I)v;
    out_eid[o + slot] = eids ? eids[s + k] : (I)(s + k);
  };
  if (!prob) {
    if (!replace && cnt == deg) {
      for (int64_t k = 0; k < deg; ++k) emit(k, k);
    } else if (replace) {
      for (int64_t k = 0; k < cnt; ++k) emit(k, rng_below(seed, stream, k, (uint32_t)deg));
    } else if (cnt <= kMaxFloyd) {
      uint32_t chosen[kMaxFloyd];
      int m = 0;
      for (int64_t j = deg - cnt; j < deg; ++j, ++m) {
        uint32_t t = rng_below(seed, stream, m, (uint32_t)(j + 1));
        bool dup = false;
        for (int q = 0; q < m; ++q) dup |= (chosen[q] == t);
        chosen[m] = dup ? (uint32_t)j : t;
      }
      for (int k = 0; k < cnt; ++k) emit(k, chosen[k]);
    } else {
      // selection sampling (Knuth algorithm S): O(deg), exact uniform subset
      int64_t taken = 0;
      for (int64_t k = 0; k < deg && taken < cnt; ++k) {
        const float u = rng_uniform(seed, stream, k);
        if (u * (float)(deg - k) < (float)(cnt - taken)) emit(taken++, k);
      }
    }
  } else {
    auto w = [&](int64_t k) { return prob[eids ? (int64_t)eids[s + k] : s + k]; };
    if (replace) {
      float tot = 0.f;
      for (int64_t k = 0; k < deg; ++k) tot += fmaxf(w(k), 0.f);
      for (int64_t c = 0; c < cnt; ++c) {
        const float r = rng_uniform(seed, stream, c) * tot;
        float acc = 0.f;
        int64_t pick = -1;
        for (int64_t k = 0; k < deg; ++k) {
          const float wk = fmaxf(w(k), 0.f);
          if (wk <= 0.f) continue;
          acc += wk;
          pick = k;
          if (acc >= r) break;
        }
        emit(c, pick);
      }
    } else {
      // Exponential race (Efraimidis-Spirakis): keep the cnt smallest keys -log(u_k) / w_k. The keys are a pure function
      // of (seed, row, k), so rows that need more than KMAX picks (fanout = -1 or fanout > 128 on a hub) are served in
      // several passes: each pass collects the next KMAX smallest (key, k) pairs above the last emitted pair in a
      // sorted register/local buffer. Every slot [0, cnt) of the output is therefore written for any cnt.
      constexpr int KMAX = 128;
      float keys[KMAX];
      int32_t idx[KMAX];
      float last_key = -1.f;
      int64_t last_k = -1, emitted = 0;
      while (emitted < cnt) {
        const int kk = (int)min(cnt - emitted, (int64_t)KMAX);
        int m = 0;
        for (int64_t k = 0; k < deg; ++k) {
          const float wk = w(k);
          if (wk <= 0.f) continue;
          const float key = -__logf(rng_uniform(seed, stream, k)) / wk;
          if (key < last_key || (key == last_key && k <= last_k)) continue;  // emitted by an earlier pass
          if (m < kk) {
            int p = m++;
            while (p > 0 && keys[p - 1] > key) { keys[p] = keys[p - 1]; idx[p] = idx[p - 1]; --p; }
            keys[p] = key; idx[p] = (int32_t)k;
          } else if (key < keys[kk - 1]) {
            int p = kk - 1;
            while (p > 0 && keys[p - 1] > key) { keys[p] = keys[p - 1]; idx[p] = idx[p - 1]; --p; }
            keys[p] = key; idx[p] = (int32_t)k;
          }
        }
        if (m == 0) break;  // cannot happen: cnt <= number of positive-probability edges (sample_count_kernel)
        for (int c = 0; c < m; ++c) emit(emitted + c, idx[c]);
        emitted += m;
        last_key = keys[m - 1];
        last_k = idx[m - 1];
      }
    }
  }
}

}  // namespace dglb

using namespace dglb;

// graph arrays may be host-pinned (uva=1): device aliases are resolved here.
static const void* dev_alias(const void* p, int uva) {
  if (!uva || !p) return p;
  void* d = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(&d, const_cast<void*>(p), 0);
  if (e != cudaSuccess) throw std::runtime_error(std::string("cudaHostGetDevicePointer: ") + cudaGetErrorString(e));
  return d;
}

#define DGLB_ELL_GROUP(G)                                                                                         \
  sample_ell_group_kernel<I, G><<<(unsigned)((n_cap * G + 255) / 256), 256, 0, (cudaStream_t)stream>>>(            \
      (const I*)ip, (const I*)ix, (const I*)seeds, n_cap, (const int64_t*)n_dev, fanout, (uint64_t)seed,           \
      (const int64_t*)seed_dev, (I*)out_nbr, (I*)out_eid, (int32_t*)out_cnt, (int64_t*)out_total, pad_rows)

DGLB_API int dglb200_cuda_sample_ell(int idtype, const void* indptr, const void* indices, int uva, const void* seeds,
                                     int64_t n_cap, const void* n_dev, int fanout, int64_t seed, const void* seed_dev,
                                     void* out_nbr, void* out_eid, void* out_cnt, void* out_total, int pad_rows,
                                     void* stream) {
  DGLB_TRY
  DGLB_CHECK(fanout > 0 && fanout <= kMaxFloyd, "sample_ell supports 1 <= fanout <= 64");
  if (n_cap == 0) return 0;
  const void* ip = dev_alias(indptr, uva);
  const void* ix = dev_alias(indices, uva);
  DGLB_ID_SWITCH(idtype, I, {
    if (fanout <= 8) DGLB_ELL_GROUP(8);
    else if (fanout <= 16) DGLB_ELL_GROUP(16);
    else if (fanout <= 32) DGLB_ELL_GROUP(32);
    else
      sample_ell_kernel<I><<<(unsigned)((n_cap + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
          (const I*)ip, (const I*)ix, (const I*)seeds, n_cap, (const int64_t*)n_dev, fanout, (uint64_t)seed,
          (const int64_t*)seed_dev, (I*)out_nbr, (I*)out_eid, (int32_t*)out_cnt, (int64_t*)out_total, pad_rows);

  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_sample_count(int idtype, const void* indptr, const void* eids, int uva, const void* seeds,
                                       int64_t n, int64_t fanout, int replace, const void* prob, void* counts,
                                       void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  const void* ip = dev_alias(indptr, uva);
  const void* ep = dev_alias(eids, uva);
  DGLB_ID_SWITCH(idtype, I, {
    sample_count_kernel<I><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const I*)ip, (const I*)ep, (const I*)seeds, n, fanout, replace, (const float*)prob, (int64_t*)counts);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_sample_fill(int idtype, const void* indptr, const void* indices, const void* eids, int uva,
                                      const void* seeds, int64_t n, int64_t fanout, int replace, const void* prob,
                                      const void* offsets, int64_t seed, void* out_src, void* out_dst, void* out_eid,
                                      void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  const void* ip = dev_alias(indptr, uva);
  const void* ix = dev_alias(indices, uva);
  const void* ep = dev_alias(eids, uva);
  DGLB_ID_SWITCH(idtype, I, {
    sample_fill_kernel<I><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const I*)ip, (const I*)ix, (const I*)ep, (const I*)seeds, n, fanout, replace, (const float*)prob,
        (const int64_t*)offsets, (uint64_t)seed, (I*)out_src, (I*)out_dst, (I*)out_eid);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
