// Row-wise pickers beyond plain uniform / per-edge-weighted sampling:
//   * per-edge-type fan-outs on a homogenised graph   (reference src/array/cpu/rowwise_pick.h:230-520
//     `CSRRowWisePerEtypePick`, python/dgl/sampling/neighbor.py:69-220 `sample_etype_neighbors`)
//   * tag-biased sampling over tag-sorted neighbour lists (reference src/array/cpu/rowwise_sampling.cc:136-260
//     `GetSamplingBiasedPickFn`, python/dgl/sampling/neighbor.py:690-870)
//   * top-k by edge weight                               (reference src/array/cpu/rowwise_topk.cc:21-122)
// Same conventions as sampling.cc: counter-based per-row RNG (results independent of the thread count), a count
// phase and a fill phase so the outputs are allocated by torch, rows are seed positions, the neighbour column is
// written to `out_nb`, the seed id to `out_seed`.
#include "common.h"

namespace dglb {

namespace {

struct Rng2 {
  uint64_t s;
  Rng2(uint64_t seed, uint64_t stream) {
    s = seed * 0x9E3779B97F4A7C15ull + (stream + 1) * 0xD6E8FEB86659FD93ull;
    next();
    next();
  }
  inline uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
  inline double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

inline int64_t want(int64_t avail, int64_t fanout, bool replace) {
  if (avail == 0 || fanout == 0) return 0;
  if (fanout < 0) return avail;
  return replace ? fanout : std::min(fanout, avail);
}

// pick `cnt` of the candidate positions `cand` (weights `w`, empty = uniform) and append them to `out`
inline void pick_from(const std::vector<int64_t>& cand, const std::vector<double>& w, int64_t cnt, bool replace, Rng2& rng,
                      std::vector<std::pair<double, int64_t>>& keys, std::vector<int64_t>& out) {
  const int64_t m = (int64_t)cand.size();
  if (cnt == 0) return;
  if (w.empty()) {
    if (!replace && cnt >= m) {
      out.insert(out.end(), cand.begin(), cand.end());
    } else if (replace) {
      for (int64_t k = 0; k < cnt; ++k) out.push_back(cand[rng.below(m)]);
    } else {
      // partial Fisher-Yates over a scratch copy
      keys.clear();
      for (int64_t k = 0; k < m; ++k) keys.emplace_back(0.0, cand[k]);
      for (int64_t k = 0; k < cnt; ++k) {
        const int64_t j = k + (int64_t)rng.below(m - k);
        std::swap(keys[k], keys[j]);
        out.push_back(keys[k].second);
      }
    }
    return;
  }
  if (replace) {
    keys.clear();
    double tot = 0;
    for (int64_t k = 0; k < m; ++k) {
      tot += w[k];
      keys.emplace_back(tot, cand[k]);
    }
    for (int64_t k = 0; k < cnt; ++k) {
      const double r = rng.uniform() * tot;
      auto it = std::upper_bound(keys.begin(), keys.end(), std::make_pair(r, std::numeric_limits<int64_t>::max()));
      if (it == keys.end()) --it;
      out.push_back(it->second);
    }
  } else {
    keys.clear();
    for (int64_t k = 0; k < m; ++k) keys.emplace_back(-std::log(std::max(rng.uniform(), 1e-300)) / w[k], cand[k]);
    if (m > cnt) std::partial_sort(keys.begin(), keys.begin() + cnt, keys.end());
    for (int64_t k = 0; k < cnt; ++k) out.push_back(keys[k].second);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- per-edge-type fan-outs
// etype of an edge = index of the [etype_offset[t], etype_offset[t+1]) range its edge id falls in.
template <typename I>
void etype_pick(const I* indptr, const I* indices, const I* eids, const I* seeds, int64_t n, const int64_t* etype_offset,
                int64_t num_etypes, const int64_t* fanouts, bool replace, const float* prob, const int64_t* offsets,
                uint64_t seed, int64_t* counts, I* out_nb, I* out_seed, I* out_eid) {
  const bool fill = offsets != nullptr;
  parallel_for_dynamic(0, n, 128, [&](int64_t b, int64_t e_) {
    std::vector<std::vector<int64_t>> cand(num_etypes);
    std::vector<std::vector<double>> wts(num_etypes);
    std::vector<std::pair<double, int64_t>> keys;
    std::vector<int64_t> picked;
    for (int64_t i = b; i < e_; ++i) {
      const int64_t v = seeds[i];
      const int64_t s = indptr[v], deg = indptr[v + 1] - s;
      for (auto& c : cand) c.clear();
      for (auto& w : wts) w.clear();
      for (int64_t k = 0; k < deg; ++k) {
        const int64_t eid = eids ? (int64_t)eids[s + k] : s + k;
        const int64_t t = (std::upper_bound(etype_offset, etype_offset + num_etypes + 1, eid) - etype_offset) - 1;
        if (t < 0 || t >= num_etypes) continue;
        if (prob) {
          const double w = prob[eid];
          if (!(w > 0)) continue;
          wts[t].push_back(w);
        }
        cand[t].push_back(k);
      }
      int64_t total = 0;
      for (int64_t t = 0; t < num_etypes; ++t) total += want((int64_t)cand[t].size(), fanouts[t], replace);
      if (!fill) {
        counts[i] = total;
        continue;
      }
      if (total == 0) continue;
      Rng2 rng(seed, (uint64_t)i);
      picked.clear();
      for (int64_t t = 0; t < num_etypes; ++t)
        pick_from(cand[t], wts[t], want((int64_t)cand[t].size(), fanouts[t], replace), replace, rng, keys, picked);
      int64_t o = offsets[i];
      for (int64_t k : picked) {
        out_nb[o] = indices[s + k];
        out_seed[o] = I(v);
        out_eid[o] = eids ? eids[s + k] : I(s + k);
        ++o;
      }
    }
  });
}

// ------------------------------------------------------------------------------------------------- tag-biased sampling
// Neighbours of row v are sorted by tag; tag_offset[v * (T + 1) + t] is the first position of tag t inside the row.
template <typename I>
void biased_pick(const I* indptr, const I* indices, const I* eids, const I* seeds, int64_t n, const int64_t* tag_offset,
                 int64_t num_tags, const double* bias, int64_t fanout, bool replace, const int64_t* offsets, uint64_t seed,
                 int64_t* counts, I* out_nb, I* out_seed, I* out_eid) {
  const bool fill = offsets != nullptr;
  parallel_for_dynamic(0, n, 256, [&](int64_t b, int64_t e_) {
    std::vector<double> cum(num_tags);
    std::vector<std::pair<double, int64_t>> keys;
    for (int64_t i = b; i < e_; ++i) {
      const int64_t v = seeds[i];
      const int64_t s = indptr[v];
      const int64_t* off = tag_offset + v * (num_tags + 1);
      int64_t avail = 0;
      double tot = 0;
      for (int64_t t = 0; t < num_tags; ++t) {
        const int64_t len = off[t + 1] - off[t];
        if (bias[t] > 0) {
          avail += len;
          tot += bias[t] * (double)len;
        }
        cum[t] = tot;
      }
      const int64_t cnt = want(avail, fanout, replace);
      if (!fill) {
        counts[i] = cnt;
        continue;
      }
      if (cnt == 0) continue;
      Rng2 rng(seed, (uint64_t)i);
      int64_t o = offsets[i];
      auto emit = [&](int64_t k) {
        out_nb[o] = indices[s + k];
        out_seed[o] = I(v);
        out_eid[o] = eids ? eids[s + k] : I(s + k);
        ++o;
      };
      if (replace) {
        // two-level draw: tag proportional to bias * population, then uniform inside the tag's position range
        for (int64_t k = 0; k < cnt; ++k) {
          const double r = rng.uniform() * tot;
          int64_t t = std::upper_bound(cum.begin(), cum.end(), r) - cum.begin();
          if (t >= num_tags) t = num_tags - 1;
          while (t > 0 && !(bias[t] > 0 && off[t + 1] > off[t])) --t;
          emit(off[t] + (int64_t)rng.below(off[t + 1] - off[t]));
        }
      } else if (cnt == avail) {
        for (int64_t t = 0; t < num_tags; ++t)
          if (bias[t] > 0)
            for (int64_t k = off[t]; k < off[t + 1]; ++k) emit(k);
      } else {
        keys.clear();
        for (int64_t t = 0; t < num_tags; ++t) {
          if (!(bias[t] > 0)) continue;
          for (int64_t k = off[t]; k < off[t + 1]; ++k)
            keys.emplace_back(-std::log(std::max(rng.uniform(), 1e-300)) / bias[t], k);
        }
        std::partial_sort(keys.begin(), keys.begin() + cnt, keys.end());
        for (int64_t k = 0; k < cnt; ++k) emit(keys[k].second);
      }
    }
  });
}

// ------------------------------------------------------------------------------------------------------------- top-k
template <typename I, typename W>
void topk_pick(const I* indptr, const I* indices, const I* eids, const I* seeds, int64_t n, const W* weight, int64_t k,
               bool ascending, const int64_t* offsets, I* out_nb, I* out_seed, I* out_eid) {
  parallel_for_dynamic(0, n, 256, [&](int64_t b, int64_t e_) {
    std::vector<int64_t> pos;
    for (int64_t i = b; i < e_; ++i) {
      const int64_t v = seeds[i];
      const int64_t s = indptr[v], deg = indptr[v + 1] - s;
      const int64_t cnt = offsets[i + 1] - offsets[i];
      if (cnt == 0) continue;
      pos.resize(deg);
      for (int64_t j = 0; j < deg; ++j) pos[j] = j;
      auto wt = [&](int64_t j) { return weight[eids ? (int64_t)eids[s + j] : s + j]; };
      // ties keep the earlier position first, like a stable sort of the row
      auto less = [&](int64_t a, int64_t c) {
        const W wa = wt(a), wc = wt(c);
        if (wa != wc) return ascending ? wa < wc : wa > wc;
        return a < c;
      };
      if (cnt < deg) std::partial_sort(pos.begin(), pos.begin() + cnt, pos.end(), less);
      else std::sort(pos.begin(), pos.end(), less);
      int64_t o = offsets[i];
      for (int64_t j = 0; j < cnt; ++j, ++o) {
        out_nb[o] = indices[s + pos[j]];
        out_seed[o] = I(v);
        out_eid[o] = eids ? eids[s + pos[j]] : I(s + pos[j]);
      }
    }
  });
}

}  // namespace dglb

using namespace dglb;

// offsets == NULL: count phase (writes counts[n]); otherwise fill phase.
DGLB_API int dglb200_cpu_etype_pick(int idtype, const void* indptr, const void* indices, const void* eids, const void* seeds,
                                    int64_t n, const void* etype_offset, int64_t num_etypes, const void* fanouts,
                                    int replace, const void* prob, const void* offsets, int64_t seed, void* counts,
                                    void* out_nb, void* out_seed, void* out_eid) {
  DGLB_TRY
  DGLB_CHECK(num_etypes > 0, "etype_pick: no edge types");
  DGLB_ID_SWITCH(idtype, I, {
    etype_pick<I>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, (const int64_t*)etype_offset,
                  num_etypes, (const int64_t*)fanouts, replace != 0, (const float*)prob, (const int64_t*)offsets,
                  (uint64_t)seed, (int64_t*)counts, (I*)out_nb, (I*)out_seed, (I*)out_eid);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_biased_pick(int idtype, const void* indptr, const void* indices, const void* eids, const void* seeds,
                                     int64_t n, const void* tag_offset, int64_t num_tags, const void* bias, int64_t fanout,
                                     int replace, const void* offsets, int64_t seed, void* counts, void* out_nb,
                                     void* out_seed, void* out_eid) {
  DGLB_TRY
  DGLB_CHECK(num_tags > 0, "biased_pick: no tags");
  DGLB_ID_SWITCH(idtype, I, {
    biased_pick<I>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, (const int64_t*)tag_offset,
                   num_tags, (const double*)bias, fanout, replace != 0, (const int64_t*)offsets, (uint64_t)seed,
                   (int64_t*)counts, (I*)out_nb, (I*)out_seed, (I*)out_eid);
  });
  DGLB_CATCH
}

// wdtype: 0 = float32, 1 = float64, 2 = int64 weights (indexed by edge id)
DGLB_API int dglb200_cpu_topk_pick(int idtype, int wdtype, const void* indptr, const void* indices, const void* eids,
                                   const void* seeds, int64_t n, const void* weight, int64_t k, int ascending,
                                   const void* offsets, void* out_nb, void* out_seed, void* out_eid) {
  DGLB_TRY
  DGLB_CHECK(wdtype >= 0 && wdtype <= 2, "topk_pick: weight dtype must be float32, float64 or int64");
  DGLB_ID_SWITCH(idtype, I, {
    if (wdtype == 0)
      topk_pick<I, float>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, (const float*)weight, k,
                          ascending != 0, (const int64_t*)offsets, (I*)out_nb, (I*)out_seed, (I*)out_eid);
    else if (wdtype == 1)
      topk_pick<I, double>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, (const double*)weight,
                           k, ascending != 0, (const int64_t*)offsets, (I*)out_nb, (I*)out_seed, (I*)out_eid);
    else
      topk_pick<I, int64_t>((const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n,
                            (const int64_t*)weight, k, ascending != 0, (const int64_t*)offsets, (I*)out_nb, (I*)out_seed,
                            (I*)out_eid);
  });
  DGLB_CATCH
}
