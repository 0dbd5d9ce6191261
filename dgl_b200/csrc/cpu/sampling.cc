// CPU neighbor sampling, order-preserving unique/compact (block construction),
// random walks. Behavioural spec: SURVEY.md A.5/A.6
// (/root/reference/python/dgl/sampling/neighbor.py:222-398,
//  /root/reference/src/array/cpu/rowwise_pick.h, src/graph/transform/to_block.cc).
// Fresh design: counter-based per-row RNG (deterministic under any thread
// count), two-phase count/fill so outputs are torch-allocated, and a flat
// open-addressing table for first-occurrence-ordered compaction.
#include "common.h"

namespace dglb {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed, uint64_t stream) {
    s = seed * 0x9E3779B97F4A7C15ull + (stream + 1) * 0xBF58476D1CE4E5B9ull;
    next();
    next();
  }
  inline uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
  inline double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

// number of picks for one row
template <typename I>
inline int64_t pick_count(int64_t deg, int64_t fanout, bool replace, const float* prob, const I* eids, int64_t start) {
  if (deg == 0) return 0;
  if (prob) {
    int64_t nz = 0;
    for (int64_t k = 0; k < deg; ++k) nz += prob[eids ? eids[start + k] : start + k] > 0.f;
    if (nz == 0) return 0;
    if (fanout < 0) return nz;
    return replace ? fanout : std::min<int64_t>(fanout, nz);
  }
  if (fanout < 0) return deg;
  return replace ? fanout : std::min<int64_t>(fanout, deg);
}

template <typename I>
void sample_count(const I* indptr, const I* eids, const I* seeds, int64_t n, int64_t fanout, bool replace,
                  const float* prob, int64_t* counts) {
  parallel_for(0, n, 1024, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      const int64_t v = seeds[i];
      const int64_t s = indptr[v], deg = indptr[v + 1] - s;
      counts[i] = pick_count<I>(deg, fanout, replace, prob, eids, s);
    }
  });
}

template <typename I>
void sample_fill(const I* indptr, const I* indices, const I* eids, const I* seeds, int64_t n, int64_t fanout,
                 bool replace, const float* prob, const int64_t* offsets, uint64_t seed, I* out_src, I* out_dst,
                 I* out_eid) {
  parallel_for_dynamic(0, n, 256, [&](int64_t b, int64_t e_) {
    std::vector<int64_t> tmp;
    std::vector<std::pair<double, int64_t>> keys;
    for (int64_t i = b; i < e_; ++i) {
      const int64_t v = seeds[i];
      const int64_t s = indptr[v], deg = indptr[v + 1] - s;
      const int64_t o = offsets[i], cnt = offsets[i + 1] - o;
      if (cnt == 0) continue;
      Rng rng(seed, (uint64_t)i);
      auto emit = [&](int64_t slot, int64_t k) {
        out_src[o + slot] = indices[s + k];
        out_dst[o + slot] = I(v);
        out_eid[o + slot] = eids ? eids[s + k] : I(s + k);
      };
      if (!prob) {
        if (cnt == deg && !replace) {
          for (int64_t k = 0; k < deg; ++k) emit(k, k);
        } else if (replace) {
          for (int64_t k = 0; k < cnt; ++k) emit(k, (int64_t)rng.below(deg));
        } else if (cnt * 4 < deg) {
          // Floyd's algorithm: cnt distinct values in [0, deg)
          tmp.clear();
          for (int64_t j = deg - cnt; j < deg; ++j) {
            int64_t t = (int64_t)rng.below(j + 1);
            bool dup = false;
            for (auto x : tmp) dup |= (x == t);
            tmp.push_back(dup ? j : t);
          }
          for (int64_t k = 0; k < cnt; ++k) emit(k, tmp[k]);
        } else {
          tmp.resize(deg);
          for (int64_t k = 0; k < deg; ++k) tmp[k] = k;
          for (int64_t k = 0; k < cnt; ++k) {
            int64_t j = k + (int64_t)rng.below(deg - k);
            std::swap(tmp[k], tmp[j]);
            emit(k, tmp[k]);
          }
        }
      } else {
        auto w = [&](int64_t k) { return (double)prob[eids ? eids[s + k] : s + k]; };
        if (replace) {
          tmp.clear();
          keys.clear();
          double tot = 0;
          for (int64_t k = 0; k < deg; ++k) {
            tot += std::max(0.0, w(k));
            keys.emplace_back(tot, k);
          }
          for (int64_t k = 0; k < cnt; ++k) {
            double r = rng.uniform() * tot;
            auto it = std::upper_bound(keys.begin(), keys.end(), std::make_pair(r, std::numeric_limits<int64_t>::max()));
            if (it == keys.end()) --it;
            while (w(it->second) <= 0 && it != keys.begin()) --it;
            emit(k, it->second);
          }
        } else {
          // exponential-race keys: smallest -log(u)/w wins (equivalent to A-Res)
          keys.clear();
          for (int64_t k = 0; k < deg; ++k) {
            double wk = w(k);
            if (wk > 0) keys.emplace_back(-std::log(std::max(rng.uniform(), 1e-300)) / wk, k);
          }
          if ((int64_t)keys.size() > cnt) std::partial_sort(keys.begin(), keys.begin() + cnt, keys.end());
          for (int64_t k = 0; k < cnt; ++k) emit(k, keys[k].second);
        }
      }
    }
  });
}

// Ordered hash table id -> local index. `first` ids (dst nodes, assumed unique) get
// [0, nfirst), remaining ids get indices in first-occurrence order (reference semantics:
// src/runtime/cuda/cuda_hashtable.cuh, transforms/to_block.py include_dst_in_src).
template <typename I>
struct IdTable {
  std::vector<I> keys;
  std::vector<I> vals;
  uint64_t mask;
  explicit IdTable(int64_t n) {
    uint64_t cap = 16;
    while ((int64_t)cap < 2 * n + 2) cap <<= 1;
    keys.assign(cap, I(-1));
    vals.assign(cap, I(-1));
    mask = cap - 1;
  }
  static inline uint64_t hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    return k;
  }
  // returns local index; inserts with `next` if absent
  inline I insert(I id, I& next) {
    uint64_t h = hash((uint64_t)id) & mask;
    while (true) {
      if (keys[h] == id) return vals[h];
      if (keys[h] == I(-1)) {
        keys[h] = id;
        vals[h] = next;
        return next++;
      }
      h = (h + 1) & mask;
    }
  }
  inline I find(I id) const {
    uint64_t h = hash((uint64_t)id) & mask;
    while (true) {
      if (keys[h] == id) return vals[h];
      if (keys[h] == I(-1)) return I(-1);
      h = (h + 1) & mask;
    }
  }
};

template <typename I>
int64_t unique_compact(const I* first, int64_t nfirst, const I* ids, int64_t n, I* out_unique, I* out_local) {
  IdTable<I> tab(nfirst + n);
  I next = 0;
  for (int64_t i = 0; i < nfirst; ++i) {
    I before = next;
    I loc = tab.insert(first[i], next);
    if (next != before) out_unique[loc] = first[i];
  }
  for (int64_t i = 0; i < n; ++i) {
    I before = next;
    I loc = tab.insert(ids[i], next);
    if (next != before) out_unique[loc] = ids[i];
    out_local[i] = loc;
  }
  return next;
}

// map ids through a table built from `universe` (universe[i] -> i); -1 if absent
template <typename I>
void map_ids(const I* universe, int64_t nu, const I* ids, int64_t n, I* out) {
  IdTable<I> tab(nu);
  I next = 0;
  for (int64_t i = 0; i < nu; ++i) tab.insert(universe[i], next);
  parallel_for(0, n, 4096, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) out[i] = tab.find(ids[i]);
  });
}

// Uniform random walk with optional restart and node2vec (p, q) bias on a CSR (out-edges).
template <typename I>
void random_walk(const I* indptr, const I* indices, const I* eids, const I* seeds, int64_t n, int64_t length,
                 double restart_prob, double p, double q, uint64_t seed, I* traces, I* trace_eids) {
  const bool n2v = (p != 1.0 || q != 1.0);
  parallel_for_dynamic(0, n, 256, [&](int64_t b, int64_t e_) {
    for (int64_t i = b; i < e_; ++i) {
      Rng rng(seed, (uint64_t)i);
      I cur = seeds[i], prev = I(-1);
      I* tr = traces + i * (length + 1);
      I* te = trace_eids ? trace_eids + i * length : nullptr;
      tr[0] = cur;
      int64_t step = 0;
      for (; step < length; ++step) {
        if (cur < 0) break;
        const int64_t s = indptr[cur], deg = indptr[cur + 1] - s;
        if (deg == 0) break;
        if (restart_prob > 0 && rng.uniform() < restart_prob) break;
        int64_t k;
        if (!n2v || prev < 0) {
          k = (int64_t)rng.below(deg);
        } else {
          // rejection sampling for node2vec
          const double mx = std::max(1.0, std::max(1.0 / p, 1.0 / q));
          const int64_t ps = indptr[prev], pe = indptr[prev + 1];
          while (true) {
            k = (int64_t)rng.below(deg);
            const I nxt = indices[s + k];
            double w;
            if (nxt == prev) w = 1.0 / p;
            else {
              bool adj = false;
              for (int64_t t = ps; t < pe && !adj; ++t) adj = indices[t] == nxt;
              w = adj ? 1.0 : 1.0 / q;
            }
            if (rng.uniform() * mx <= w) break;
          }
        }
        prev = cur;
        cur = indices[s + k];
        tr[step + 1] = cur;
        if (te) te[step] = eids ? eids[s + k] : I(s + k);
      }
      for (int64_t r = step; r < length; ++r) {
        tr[r + 1] = I(-1);
        if (te) te[r] = I(-1);
      }
    }
  });
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cpu_sample_count(int idtype, const void* indptr, const void* eids, const void* seeds, int64_t n,
                                      int64_t fanout, int replace, const void* prob, void* counts) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    sample_count<I>((const I*)indptr, (const I*)eids, (const I*)seeds, n, fanout, replace != 0, (const float*)prob,
                    (int64_t*)counts);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_sample_fill(int idtype, const void* indptr, const void* indices, const void* eids,
                                     const void* seeds, int64_t n, int64_t fanout, int replace, const void* prob,
                                     const void* offsets, int64_t seed, void* out_src, void* out_dst, void* out_eid) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    sample_fill<I>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, fanout, replace != 0,
                   (const float*)prob, (const int64_t*)offsets, (uint64_t)seed, (I*)out_src, (I*)out_dst,
                   (I*)out_eid);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_unique_compact(int idtype, const void* first, int64_t nfirst, const void* ids, int64_t n,
                                        void* out_unique, void* out_local, void* out_count) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    *(int64_t*)out_count = unique_compact<I>((const I*)first, nfirst, (const I*)ids, n, (I*)out_unique, (I*)out_local);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_map_ids(int idtype, const void* universe, int64_t nu, const void* ids, int64_t n, void* out) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, { map_ids<I>((const I*)universe, nu, (const I*)ids, n, (I*)out); });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_random_walk(int idtype, const void* indptr, const void* indices, const void* eids,
                                     const void* seeds, int64_t n, int64_t length, double restart_prob, double p,
                                     double q, int64_t seed, void* traces, void* trace_eids) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    random_walk<I>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, length, restart_prob, p,
                   q, (uint64_t)seed, (I*)traces, (I*)trace_eids);
  });
  DGLB_CATCH
}
