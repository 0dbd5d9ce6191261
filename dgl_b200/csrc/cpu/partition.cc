// Graph partitioners: balanced edge-cut (streaming LDG + balanced label-propagation refinement) and
// Libra vertex-cut.
// Reference behaviour: METIS k-way wrappers (/root/reference/src/graph/metis_partition.cc,
// transform/metis_partition_hetero.cc) and Libra (/root/reference/src/array/libra_partition.cc).
// METIS itself is third-party code in the reference; this is an independent O(E) partitioner that
// optimises the same objective (edge cut under a vertex/edge balance constraint).
#include <numeric>
#include <random>

#include "common.h"

namespace dglb {

// CSR of the symmetrised graph is expected (indptr[n+1], indices). vwgt optional (balance constraint weights).
template <typename I>
void ldg_partition(const I* indptr, const I* indices, int64_t n, int k, const int64_t* vwgt, int refine_iters,
                   double imbalance, uint64_t seed, int64_t* part) {
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  // BFS order from random roots gives LDG much better locality than a random stream
  {
    std::mt19937_64 rng(seed);
    std::vector<char> seen(n, 0);
    std::vector<int64_t> bfs;
    bfs.reserve(n);
    std::vector<int64_t> roots(n);
    std::iota(roots.begin(), roots.end(), 0);
    std::shuffle(roots.begin(), roots.end(), rng);
    size_t head = 0;
    for (int64_t r : roots) {
      if (seen[r]) continue;
      seen[r] = 1;
      bfs.push_back(r);
      while (head < bfs.size()) {
        int64_t u = bfs[head++];
        for (I p = indptr[u]; p < indptr[u + 1]; ++p) {
          int64_t v = indices[p];
          if (!seen[v]) { seen[v] = 1; bfs.push_back(v); }
        }
      }
    }
    order.swap(bfs);
  }
  double total_w = 0;
  for (int64_t v = 0; v < n; ++v) total_w += vwgt ? (double)vwgt[v] : 1.0;
  const double cap = imbalance * total_w / k;
  std::vector<double> load(k, 0.0);
  std::vector<double> score(k);
  for (int64_t v = 0; v < n; ++v) part[v] = -1;
  for (int64_t v : order) {
    std::fill(score.begin(), score.end(), 0.0);
    for (I p = indptr[v]; p < indptr[v + 1]; ++p) {
      int64_t u = indices[p];
      if (part[u] >= 0) score[part[u]] += 1.0;
    }
    int best = -1;
    double bs = -1;
    for (int i = 0; i < k; ++i) {
      if (load[i] >= cap) continue;
      double s = (score[i] + 1e-3) * (1.0 - load[i] / cap);
      if (s > bs) { bs = s; best = i; }
    }
    if (best < 0) best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    part[v] = best;
    load[best] += vwgt ? (double)vwgt[v] : 1.0;
  }
  // balanced label-propagation refinement: move a vertex to the partition holding most of its
  // neighbours when that reduces the cut and keeps the balance constraint
  for (int it = 0; it < refine_iters; ++it) {
    int64_t moved = 0;
    for (int64_t v : order) {
      std::fill(score.begin(), score.end(), 0.0);
      for (I p = indptr[v]; p < indptr[v + 1]; ++p) score[part[indices[p]]] += 1.0;
      const int cur = (int)part[v];
      const double w = vwgt ? (double)vwgt[v] : 1.0;
      int best = cur;
      double gain = 0;
      for (int i = 0; i < k; ++i) {
        if (i == cur || load[i] + w > cap) continue;
        double g = score[i] - score[cur];
        if (g > gain) { gain = g; best = i; }
      }
      if (best != cur) {
        part[v] = best;
        load[cur] -= w;
        load[best] += w;
        ++moved;
      }
    }
    if (moved == 0) break;
  }
}

// Libra vertex-cut: edges are streamed; each edge goes to the least-loaded partition among those that
// already hold one of its endpoints (else globally least loaded).
template <typename I>
void libra_partition(const I* src, const I* dst, int64_t E, int64_t n, int k, int64_t* edge_part) {
  std::vector<int64_t> load(k, 0);
  std::vector<uint64_t> mask_words((size_t)n * ((k + 63) / 64), 0);
  const int W = (k + 63) / 64;
  auto has = [&](int64_t v, int p) { return (mask_words[(size_t)v * W + p / 64] >> (p % 64)) & 1ull; };
  auto set = [&](int64_t v, int p) { mask_words[(size_t)v * W + p / 64] |= 1ull << (p % 64); };
  for (int64_t e = 0; e < E; ++e) {
    const int64_t u = src[e], v = dst[e];
    int best = -1;
    int64_t bl = std::numeric_limits<int64_t>::max();
    // prefer partitions containing both endpoints, then either endpoint
    for (int pass = 0; pass < 2 && best < 0; ++pass) {
      for (int p = 0; p < k; ++p) {
        const bool hu = has(u, p), hv = has(v, p);
        const bool ok = pass == 0 ? (hu && hv) : (hu || hv);
        if (ok && load[p] < bl) { bl = load[p]; best = p; }
      }
    }
    if (best < 0) best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    edge_part[e] = best;
    ++load[best];
    set(u, best);
    set(v, best);
  }
}


// ------------------------------------------------------------------------------------------------ multilevel k-way
// The classical scheme (Karypis & Kumar) with this repo's own building blocks:
//   coarsen   heavy-edge matching in random vertex order (a vertex joins the unmatched neighbour it shares the heaviest
//             edge with, subject to a weight cap); matched pairs collapse, parallel edges add up;
//   initial   LDG streaming over the coarsest graph (edge weights as affinities, vertex weights as load);
//   refine    at every level on the way back: boundary vertices move to the part they are most strongly tied to when
//             that lowers the weighted cut and keeps the balance constraint (greedy Fiduccia-Mattheyses style passes
//             without the hill-climbing tail).
// Against plain LDG + label propagation on the fine graph this roughly halves the cut on graphs with community / mesh
// structure (tests/test_cpu_struct.py), where a streaming heuristic sees too little context per decision.
struct WGraph {
  int64_t n = 0;
  std::vector<int64_t> indptr, indices;
  std::vector<double> ewgt, vwgt;
};

static WGraph coarsen_once(const WGraph& g, double max_vwgt, std::mt19937_64& rng, std::vector<int64_t>& cmap) {
  const int64_t n = g.n;
  std::vector<int64_t> order(n), match(n, -1);
  std::iota(order.begin(), order.end(), 0);
  std::shuffle(order.begin(), order.end(), rng);
  for (int64_t v : order) {
    if (match[v] >= 0) continue;
    int64_t best = -1;
    double bw = -1;
    for (int64_t p = g.indptr[v]; p < g.indptr[v + 1]; ++p) {
      const int64_t u = g.indices[p];
      if (u == v || match[u] >= 0 || g.vwgt[v] + g.vwgt[u] > max_vwgt) continue;
      if (g.ewgt[p] > bw) { bw = g.ewgt[p]; best = u; }
    }
    match[v] = best >= 0 ? best : v;
    if (best >= 0) match[best] = v;
  }
  cmap.assign(n, -1);
  int64_t nc = 0;
  for (int64_t v = 0; v < n; ++v) {
    if (cmap[v] >= 0) continue;
    cmap[v] = nc;
    cmap[match[v]] = nc;
    ++nc;
  }
  WGraph c;
  c.n = nc;
  c.vwgt.assign(nc, 0.0);
  for (int64_t v = 0; v < n; ++v) c.vwgt[cmap[v]] += g.vwgt[v];
  // members of every coarse vertex, then merge their adjacency through a dense marker array
  std::vector<int64_t> first(nc, -1), second(nc, -1);
  for (int64_t v = 0; v < n; ++v) {
    const int64_t cv = cmap[v];
    if (first[cv] < 0) first[cv] = v; else second[cv] = v;
  }
  c.indptr.assign(nc + 1, 0);
  std::vector<int64_t> slot(nc, -1);
  for (int64_t cv = 0; cv < nc; ++cv) {
    const int64_t begin = (int64_t)c.indices.size();
    for (int64_t v : {first[cv], second[cv]}) {
      if (v < 0) continue;
      for (int64_t p = g.indptr[v]; p < g.indptr[v + 1]; ++p) {
        const int64_t cu = cmap[g.indices[p]];
        if (cu == cv) continue;
        if (slot[cu] < begin) {
          slot[cu] = (int64_t)c.indices.size();
          c.indices.push_back(cu);
          c.ewgt.push_back(g.ewgt[p]);
        } else {
          c.ewgt[slot[cu]] += g.ewgt[p];
        }
      }
    }
    c.indptr[cv + 1] = (int64_t)c.indices.size();
  }
  return c;
}

static void refine_kway(const WGraph& g, int k, double cap, int iters, std::vector<int64_t>& part, std::mt19937_64& rng) {
  std::vector<double> load(k, 0.0), tie(k);
  for (int64_t v = 0; v < g.n; ++v) load[part[v]] += g.vwgt[v];
  std::vector<int64_t> order(g.n);
  std::iota(order.begin(), order.end(), 0);
  for (int it = 0; it < iters; ++it) {
    std::shuffle(order.begin(), order.end(), rng);
    int64_t moved = 0;
    for (int64_t v : order) {
      const int cur = (int)part[v];
      bool boundary = false;
      for (int64_t p = g.indptr[v]; p < g.indptr[v + 1] && !boundary; ++p) boundary = part[g.indices[p]] != cur;
      if (!boundary) continue;
      std::fill(tie.begin(), tie.end(), 0.0);
      for (int64_t p = g.indptr[v]; p < g.indptr[v + 1]; ++p) tie[part[g.indices[p]]] += g.ewgt[p];
      int best = cur;
      double gain = 0.0;
      for (int i = 0; i < k; ++i) {
        if (i == cur || load[i] + g.vwgt[v] > cap) continue;
        const double d = tie[i] - tie[cur];
        // equal cut but better balance also counts as progress (it opens room for later moves)
        if (d > gain || (d == gain && d >= 0 && best == cur && load[i] + g.vwgt[v] < load[cur])) { gain = d; best = i; }
      }
      if (best != cur && (gain > 0 || load[cur] > cap)) {
        part[v] = best;
        load[cur] -= g.vwgt[v];
        load[best] += g.vwgt[v];
        ++moved;
      }
    }
    if (moved == 0) break;
  }
}

static void initial_kway(const WGraph& g, int k, double cap, std::mt19937_64& rng, std::vector<int64_t>& part) {
  // LDG over a BFS order with weighted affinities
  const int64_t n = g.n;
  std::vector<int64_t> order;
  order.reserve(n);
  std::vector<char> seen(n, 0);
  std::vector<int64_t> roots(n);
  std::iota(roots.begin(), roots.end(), 0);
  std::shuffle(roots.begin(), roots.end(), rng);
  size_t head = 0;
  for (int64_t r : roots) {
    if (seen[r]) continue;
    seen[r] = 1;
    order.push_back(r);
    while (head < order.size()) {
      const int64_t u = order[head++];
      for (int64_t p = g.indptr[u]; p < g.indptr[u + 1]; ++p)
        if (!seen[g.indices[p]]) { seen[g.indices[p]] = 1; order.push_back(g.indices[p]); }
    }
  }
  part.assign(n, -1);
  std::vector<double> load(k, 0.0), score(k);
  for (int64_t v : order) {
    std::fill(score.begin(), score.end(), 0.0);
    for (int64_t p = g.indptr[v]; p < g.indptr[v + 1]; ++p)
      if (part[g.indices[p]] >= 0) score[part[g.indices[p]]] += g.ewgt[p];
    int best = -1;
    double bs = -1;
    for (int i = 0; i < k; ++i) {
      if (load[i] + g.vwgt[v] > cap) continue;
      const double s = (score[i] + 1e-3) * (1.0 - load[i] / cap);
      if (s > bs) { bs = s; best = i; }
    }
    if (best < 0) best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    part[v] = best;
    load[best] += g.vwgt[v];
  }
}

template <typename I>
void multilevel_partition(const I* indptr, const I* indices, int64_t n, int k, const int64_t* vwgt, int refine_iters,
                          double imbalance, uint64_t seed, int64_t* part_out) {
  std::mt19937_64 rng(seed);
  std::vector<WGraph> levels(1);
  WGraph& g0 = levels[0];
  g0.n = n;
  g0.indptr.assign(indptr, indptr + n + 1);
  g0.indices.assign(indices, indices + indptr[n]);
  g0.ewgt.assign((size_t)indptr[n], 1.0);
  g0.vwgt.resize(n);
  double total = 0;
  for (int64_t v = 0; v < n; ++v) total += (g0.vwgt[v] = vwgt ? (double)vwgt[v] : 1.0);
  const double cap = imbalance * total / k;
  const int64_t stop_at = std::max<int64_t>(40 * k, 2000);
  std::vector<std::vector<int64_t>> cmaps;
  while (levels.back().n > stop_at && (int)levels.size() < 40) {
    std::vector<int64_t> cmap;
    WGraph c = coarsen_once(levels.back(), 0.3 * cap, rng, cmap);   // no coarse vertex above 30 % of a part
    if (c.n > 0.93 * levels.back().n) break;                         // matching stalled (star-like remainder)
    cmaps.push_back(std::move(cmap));
    levels.push_back(std::move(c));
  }
  // the coarsest graph is tiny: try several BFS orders / refinements and keep the lightest cut
  const WGraph& gc = levels.back();
  auto weighted_cut = [&](const std::vector<int64_t>& p) {
    double c = 0;
    for (int64_t v = 0; v < gc.n; ++v)
      for (int64_t q = gc.indptr[v]; q < gc.indptr[v + 1]; ++q)
        if (p[gc.indices[q]] != p[v]) c += gc.ewgt[q];
    return c;
  };
  std::vector<int64_t> part;
  double best_cut = -1;
  for (int attempt = 0; attempt < 12; ++attempt) {
    std::vector<int64_t> cand;
    initial_kway(gc, k, cap, rng, cand);
    refine_kway(gc, k, cap, refine_iters + 8, cand, rng);
    const double c = weighted_cut(cand);
    if (best_cut < 0 || c < best_cut) { best_cut = c; part.swap(cand); }
  }
  for (int l = (int)levels.size() - 2; l >= 0; --l) {
    std::vector<int64_t> fine(levels[l].n);
    for (int64_t v = 0; v < levels[l].n; ++v) fine[v] = part[cmaps[l][v]];
    part.swap(fine);
    refine_kway(levels[l], k, cap, refine_iters, part, rng);
  }
  std::copy(part.begin(), part.end(), part_out);
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cpu_ldg_partition(int idtype, const void* indptr, const void* indices, int64_t n, int k,
                                       const void* vwgt, int refine_iters, double imbalance, int64_t seed, void* part) {
  DGLB_TRY
  DGLB_CHECK(k >= 1, "k must be >= 1");
  DGLB_ID_SWITCH(idtype, I, {
    ldg_partition<I>((const I*)indptr, (const I*)indices, n, k, (const int64_t*)vwgt, refine_iters, imbalance,
                     (uint64_t)seed, (int64_t*)part);
  });
  DGLB_CATCH
}

// multilevel k-way partitioner; same arguments as dglb200_cpu_ldg_partition
DGLB_API int dglb200_cpu_ml_partition(int idtype, const void* indptr, const void* indices, int64_t n, int k,
                                      const void* vwgt, int refine_iters, double imbalance, int64_t seed, void* part) {
  DGLB_TRY
  DGLB_CHECK(k >= 1, "k must be >= 1");
  DGLB_ID_SWITCH(idtype, I, {
    multilevel_partition<I>((const I*)indptr, (const I*)indices, n, k, (const int64_t*)vwgt, refine_iters, imbalance,
                            (uint64_t)seed, (int64_t*)part);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_libra_partition(int idtype, const void* src, const void* dst, int64_t E, int64_t n, int k,
                                         void* edge_part) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, { libra_partition<I>((const I*)src, (const I*)dst, E, n, k, (int64_t*)edge_part); });
  DGLB_CATCH
}
