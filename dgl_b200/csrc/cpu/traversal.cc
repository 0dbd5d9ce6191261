// Frontier generators over a CSR adjacency: level-synchronous BFS (node or tree-edge frontiers),
// Kahn topological levels and labelled DFS edge streams.
//
// Behavioural reference: src/graph/traversal.{h,cc} (BFSNodes/BFSEdges/TopologicalNodes/
// DFSLabeledEdges).  Output convention shared by all three: a flat id array plus the length of each
// frontier ("sections"); the Python side splits.  Everything here is a single pass over the edges
// that the traversal reaches, so these are serial by design (frontier order is part of the
// contract: discovery order, ties by adjacency order).
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

}  // namespace
}  // namespace dglb

using namespace dglb;

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
