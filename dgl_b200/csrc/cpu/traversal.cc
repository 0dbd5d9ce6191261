// Frontier generators over a CSR adjacency: level-synchronous BFS (node or tree-edge frontiers),
// Kahn topological levels and labelled DFS edge streams.
//
// Behavioural reference: src/graph/traversal.{h,cc} (BFSNodes/BFSEdges/TopologicalNodes/
// DFSLabeledEdges).  Output convention shared by all three: a flat id array plus the length of each
// frontier ("sections"); the Python side splits.  Everything here is a single pass over the edges
// that the traversal reaches, so these are serial by design (frontier order is part of the
// contract: discovery order, ties by adjacency order).
#include <algorithm>
#include <vector>

#include "common.h"

namespace dglb {
namespace {

template <typename I>
void bfs(const I* indptr, const I* indices, const I* eids, int64_t n, const I* src, int64_t n_src, bool want_edges,
         I* out_ids, int64_t* sections, int64_t* n_sections, int64_t* n_ids) {
  std::vector<uint8_t> seen(n, 0);
  std::vector<I> cur, next;
  int64_t w = 0, ns = 0;
  for (int64_t i = 0; i < n_src; ++i) {
    DGLB_CHECK(src[i] >= 0 && src[i] < n, "BFS source out of range");
    if (!seen[src[i]]) {
      seen[src[i]] = 1;
      cur.push_back(src[i]);
    }
  }
  if (!want_edges && !cur.empty()) {
    for (I v : cur) out_ids[w++] = v;
    sections[ns++] = (int64_t)cur.size();
  }
  while (!cur.empty()) {
    next.clear();
    const int64_t w0 = w;
    for (I u : cur) {
      for (I p = indptr[u]; p < indptr[u + 1]; ++p) {
        const I v = indices[p];
        if (seen[v]) continue;
        seen[v] = 1;
        next.push_back(v);
        out_ids[w++] = want_edges ? (eids ? eids[p] : p) : v;
      }
    }
    if (w > w0) sections[ns++] = w - w0;
    cur.swap(next);
  }
  *n_sections = ns;
  *n_ids = w;
}

template <typename I>
void topo(const I* indptr, const I* indices, int64_t n, I* out_ids, int64_t* sections, int64_t* n_sections,
          int64_t* n_ids) {
  // indptr/indices give successors; in-degree is recomputed here so callers need no second format
  std::vector<int64_t> deg(n, 0);
  for (int64_t p = 0; p < indptr[n]; ++p) ++deg[indices[p]];
  std::vector<I> cur, next;
  for (int64_t v = 0; v < n; ++v)
    if (deg[v] == 0) cur.push_back((I)v);
  int64_t w = 0, ns = 0;
  while (!cur.empty()) {
    for (I v : cur) out_ids[w++] = v;
    sections[ns++] = (int64_t)cur.size();
    next.clear();
    for (I u : cur)
      for (I p = indptr[u]; p < indptr[u + 1]; ++p)
        if (--deg[indices[p]] == 0) next.push_back(indices[p]);
    cur.swap(next);
  }
  *n_sections = ns;
  *n_ids = w;  // < n  <=>  the graph has a cycle
}

enum { kForward = 0, kReverse = 1, kNonTree = 2 };

template <typename I>
int64_t dfs(const I* indptr, const I* indices, const I* eids, int64_t n, I source, bool has_reverse, bool has_nontree,
            I* out_edges, int64_t* out_labels, int64_t capacity) {
  DGLB_CHECK(source >= 0 && source < n, "DFS source out of range");
  std::vector<uint8_t> seen(n, 0);
  std::vector<I> node_stack, cursor;  // cursor[d] = next adjacency slot of node_stack[d]
  seen[source] = 1;
  node_stack.push_back(source);
  cursor.push_back(indptr[source]);
  int64_t w = 0;
  auto emit = [&](I p, int label) {
    DGLB_CHECK(w < capacity, "DFS output overflow");
    out_edges[w] = eids ? eids[p] : p;
    if (out_labels) out_labels[w] = label;
    ++w;
  };
  while (!node_stack.empty()) {
    const I u = node_stack.back();
    I& p = cursor.back();
    if (p == indptr[u + 1]) {
      // u is exhausted: the tree edge that led here is walked back
      node_stack.pop_back();
      cursor.pop_back();
      if (has_reverse && !node_stack.empty()) emit(cursor.back() - 1, kReverse);
      continue;
    }
    const I slot = p++;
    const I v = indices[slot];
    if (seen[v]) {
      if (has_nontree) emit(slot, kNonTree);
      continue;
    }
    seen[v] = 1;
    emit(slot, kForward);
    node_stack.push_back(v);
    cursor.push_back(indptr[v]);
  }
  return w;
}

// Reverse Cuthill-McKee on a symmetric adjacency: per connected component start from a
// pseudo-peripheral vertex (George-Liu: repeat BFS from a minimum-degree vertex of the last level
// until the eccentricity stops growing), number vertices breadth-first visiting neighbours by
// increasing degree, and finally reverse the whole numbering.
template <typename I>
void rcmk(const I* indptr, const I* indices, int64_t n, int64_t* perm) {
  std::vector<int64_t> deg(n);
  for (int64_t v = 0; v < n; ++v) deg[v] = indptr[v + 1] - indptr[v];
  std::vector<int64_t> by_degree(n);
  for (int64_t v = 0; v < n; ++v) by_degree[v] = v;
  std::stable_sort(by_degree.begin(), by_degree.end(), [&](int64_t a, int64_t b) { return deg[a] < deg[b]; });
  std::vector<uint8_t> placed(n, 0);
  std::vector<int32_t> level(n, -1);
  std::vector<I> queue, touched, nbrs;
  int64_t w = 0;

  // BFS over the not-yet-placed component of `root`; returns (eccentricity, min-degree vertex of last level)
  auto probe = [&](int64_t root) {
    for (I v : touched) level[v] = -1;
    touched.clear();
    queue.clear();
    queue.push_back((I)root);
    level[root] = 0;
    touched.push_back((I)root);
    int64_t last = root, ecc = 0;
    for (size_t h = 0; h < queue.size(); ++h) {
      const int64_t u = queue[h];
      if (level[u] > ecc || (level[u] == ecc && deg[u] < deg[last])) {
        if (level[u] > ecc) last = u;
        ecc = level[u];
        if (deg[u] < deg[last]) last = u;
      }
      for (I p = indptr[u]; p < indptr[u + 1]; ++p) {
        const int64_t v = indices[p];
        if (placed[v] || level[v] >= 0) continue;
        level[v] = level[u] + 1;
        touched.push_back((I)v);
        queue.push_back((I)v);
      }
    }
    return std::make_pair(ecc, last);
  };

  for (int64_t seed : by_degree) {
    if (placed[seed]) continue;
    int64_t root = seed;
    auto cur = probe(root);
    for (int iter = 0; iter < 8; ++iter) {
      auto nxt = probe(cur.second);
      if (nxt.first <= cur.first) break;
      root = cur.second;
      cur = nxt;
    }
    // Cuthill-McKee numbering of this component
    const int64_t first = w;
    perm[w++] = root;
    placed[root] = 1;
    for (int64_t h = first; h < w; ++h) {
      const int64_t u = perm[h];
      nbrs.clear();
      for (I p = indptr[u]; p < indptr[u + 1]; ++p)
        if (!placed[indices[p]]) {
          placed[indices[p]] = 1;
          nbrs.push_back(indices[p]);
        }
      std::stable_sort(nbrs.begin(), nbrs.end(), [&](I a, I b) { return deg[a] < deg[b]; });
      for (I v : nbrs) perm[w++] = v;
    }
  }
  std::reverse(perm, perm + n);
}

}  // namespace
}  // namespace dglb

using namespace dglb;

// RCM straight from an edge list: the symmetrised adjacency is built here with one counting sort (O(E), no
// comparison sort, self loops dropped) instead of materialising it in Python first.
template <typename I, typename J>
void rcmk_coo_impl(const I* u, const I* v, int64_t E, int64_t n, int64_t* perm) {
  // J = index type of the temporary adjacency (int32 whenever ids and edge counts fit: half the memory traffic of
  // the four or five BFS sweeps that dominate the run time)
  std::vector<J> indptr(n + 1, 0);
  for (int64_t e = 0; e < E; ++e)
    if (u[e] != v[e]) {
      ++indptr[u[e] + 1];
      ++indptr[v[e] + 1];
    }
  for (int64_t i = 0; i < n; ++i) indptr[i + 1] += indptr[i];
  std::vector<J> indices(indptr[n]);
  std::vector<J> cursor(indptr.begin(), indptr.end() - 1);
  for (int64_t e = 0; e < E; ++e)
    if (u[e] != v[e]) {
      indices[cursor[u[e]]++] = (J)v[e];
      indices[cursor[v[e]]++] = (J)u[e];
    }
  rcmk<J>(indptr.data(), indices.data(), n, perm);
}

template <typename I>
void rcmk_coo(const I* u, const I* v, int64_t E, int64_t n, int64_t* perm) {
  if (n < (int64_t(1) << 31) - 1 && 2 * E < (int64_t(1) << 31) - 1)
    rcmk_coo_impl<I, int32_t>(u, v, E, n, perm);
  else
    rcmk_coo_impl<I, int64_t>(u, v, E, n, perm);
}

DGLB_API int dglb200_cpu_rcmk_coo(int idtype, const void* u, const void* v, int64_t E, int64_t n, int64_t* perm) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, rcmk_coo<I>((const I*)u, (const I*)v, E, n, perm));
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_rcmk(int idtype, const void* indptr, const void* indices, int64_t n, int64_t* perm) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, rcmk<I>((const I*)indptr, (const I*)indices, n, perm));
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_bfs(int idtype, const void* indptr, const void* indices, const void* eids, int64_t n,
                             const void* src, int64_t n_src, int want_edges, void* out_ids, int64_t* sections,
                             int64_t* n_sections, int64_t* n_ids) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, bfs<I>((const I*)indptr, (const I*)indices, (const I*)eids, n, (const I*)src, n_src,
                                   want_edges != 0, (I*)out_ids, sections, n_sections, n_ids));
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_topological(int idtype, const void* indptr, const void* indices, int64_t n, void* out_ids,
                                     int64_t* sections, int64_t* n_sections, int64_t* n_ids) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, topo<I>((const I*)indptr, (const I*)indices, n, (I*)out_ids, sections, n_sections, n_ids));
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_dfs(int idtype, const void* indptr, const void* indices, const void* eids, int64_t n,
                             int64_t source, int has_reverse, int has_nontree, void* out_edges, int64_t* out_labels,
                             int64_t capacity, int64_t* n_out) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, *n_out = dfs<I>((const I*)indptr, (const I*)indices, (const I*)eids, n, (I)source,
                                            has_reverse != 0, has_nontree != 0, (I*)out_edges, out_labels, capacity));
  DGLB_CATCH
}
