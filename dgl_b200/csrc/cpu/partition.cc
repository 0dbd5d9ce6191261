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

DGLB_API int dglb200_cpu_libra_partition(int idtype, const void* src, const void* dst, int64_t E, int64_t n, int k,
                                         void* edge_part) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, { libra_partition<I>((const I*)src, (const I*)dst, E, n, k, (int64_t*)edge_part); });
  DGLB_CATCH
}
