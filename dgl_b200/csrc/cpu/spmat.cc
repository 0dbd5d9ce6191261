// Host sparse-structure operations on COO / CSR adjacency: format conversion, transpose, coalesce, row / sub-matrix
// slicing, line graph. Counterpart of the reference's aten structure ops
// (/root/reference/src/array/cpu/spmat_op_impl_coo.cc: COOToCSR :330-520, COOCoalesce :700-760, COOLineGraph;
//  spmat_op_impl_csr.cc: CSRTranspose :420-470, CSRSliceRows :480-560, CSRSliceMatrix :570-634;
//  src/array/cpu/coo_sort.cc, coo_coalesce.cc, coo_line_graph.cc).
// Design: everything is a stable counting sort or a two-pass (count, exclusive scan, fill) over rows, parallel over
// contiguous chunks, so edge order inside a row is the input order and results do not depend on the thread count.
// Outputs are allocated by the caller (torch); variable-size results use a count entry point first.
#include "common.h"

namespace dglb {

// Stable counting sort of E keys in [0, n): indptr[n+1], perm[E] (perm[k] = input position of the k-th sorted element).
template <typename I>
void counting_sort(const I* key, int64_t E, int64_t n, I* indptr, int64_t* perm) {
  const int T = std::max(1, std::min<int>(omp_get_max_threads(), (int)((E + 65535) / 65536)));
  if (T == 1 || n > (int64_t)8 << 20) {
    // serial histogram (also used when a per-thread histogram would not fit comfortably)
    std::vector<int64_t> cnt(n + 1, 0);
    for (int64_t e = 0; e < E; ++e) ++cnt[key[e] + 1];
    for (int64_t r = 0; r < n; ++r) cnt[r + 1] += cnt[r];
    for (int64_t r = 0; r <= n; ++r) indptr[r] = I(cnt[r]);
    for (int64_t e = 0; e < E; ++e) perm[cnt[key[e]]++] = e;
    return;
  }
  // per-chunk histograms -> per-(row, chunk) start offsets -> scatter; chunk order preserves stability
  const int64_t chunk = (E + T - 1) / T;
  std::vector<std::vector<int64_t>> hist(T, std::vector<int64_t>(n, 0));
  parallel_for(0, T, 1, [&](int64_t tb, int64_t te) {
    for (int64_t t = tb; t < te; ++t) {
      const int64_t b = t * chunk, e = std::min(E, b + chunk);
      auto& h = hist[t];
      for (int64_t k = b; k < e; ++k) ++h[key[k]];
    }
  });
  int64_t run = 0;
  for (int64_t r = 0; r < n; ++r) {
    indptr[r] = I(run);
    for (int t = 0; t < T; ++t) {
      const int64_t c = hist[t][r];
      hist[t][r] = run;
      run += c;
    }
  }
  indptr[n] = I(run);
  parallel_for(0, T, 1, [&](int64_t tb, int64_t te) {
    for (int64_t t = tb; t < te; ++t) {
      const int64_t b = t * chunk, e = std::min(E, b + chunk);
      auto& h = hist[t];
      for (int64_t k = b; k < e; ++k) perm[h[key[k]]++] = k;
    }
  });
}

template <typename I>
void coo_to_csr(const I* row, const I* col, int64_t E, int64_t n_rows, I* indptr, I* indices, I* eids) {
  std::vector<int64_t> perm(E);
  counting_sort<I>(row, E, n_rows, indptr, perm.data());
  parallel_for(0, E, 1 << 16, [&](int64_t b, int64_t e) {
    for (int64_t k = b; k < e; ++k) {
      indices[k] = col[perm[k]];
      eids[k] = I(perm[k]);
    }
  });
}

// CSR (n_rows x n_cols, optional data) -> CSR of the transpose; t_data carries the original edge ids
template <typename I>
void csr_transpose(const I* indptr, const I* indices, const I* data, int64_t n_rows, int64_t n_cols, I* t_indptr,
                   I* t_indices, I* t_data) {
  const int64_t E = indptr[n_rows];
  std::vector<int64_t> perm(E);
  counting_sort<I>(indices, E, n_cols, t_indptr, perm.data());
  std::vector<I> rowof(E);
  parallel_for(0, n_rows, 4096, [&](int64_t b, int64_t e) {
    for (int64_t r = b; r < e; ++r)
      for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) rowof[k] = I(r);
  });
  parallel_for(0, E, 1 << 16, [&](int64_t b, int64_t e) {
    for (int64_t k = b; k < e; ++k) {
      const int64_t p = perm[k];
      t_indices[k] = rowof[p];
      t_data[k] = data ? data[p] : I(p);
    }
  });
}

// Coalesce: order[E] = edge positions sorted by (row, col) (stable); group[E] = id of the distinct (row, col) pair of
// every ORIGINAL edge; returns the number of distinct pairs. first[g] / count[g] are filled when non-null.
template <typename I>
int64_t coalesce(const I* row, const I* col, int64_t E, int64_t n_rows, int64_t n_cols, int64_t* group, I* out_row,
                 I* out_col, int64_t* count, int64_t* first) {
  if (E == 0) return 0;
  // two stable counting passes: by column, then by row (LSD radix on the pair)
  std::vector<I> ptr(std::max(n_rows, n_cols) + 1);
  std::vector<int64_t> p1(E), p2(E);
  counting_sort<I>(col, E, n_cols, ptr.data(), p1.data());
  std::vector<I> rkey(E);
  parallel_for(0, E, 1 << 16, [&](int64_t b, int64_t e) {
    for (int64_t k = b; k < e; ++k) rkey[k] = row[p1[k]];
  });
  counting_sort<I>(rkey.data(), E, n_rows, ptr.data(), p2.data());
  // order[k] = p1[p2[k]]
  int64_t g = -1;
  I pr = I(-1), pc = I(-1);
  for (int64_t k = 0; k < E; ++k) {
    const int64_t e = p1[p2[k]];
    if (row[e] != pr || col[e] != pc) {
      ++g;
      pr = row[e];
      pc = col[e];
      out_row[g] = pr;
      out_col[g] = pc;
      if (count) count[g] = 0;
      if (first) first[g] = e;
    }
    group[e] = g;
    if (count) ++count[g];
    if (first && e < first[g]) first[g] = e;
  }
  return g + 1;
}

// rows[] subset of a CSR: sub_indptr[nr+1] (count pass when sub_indices == NULL), then indices / data of those rows.
template <typename I>
void csr_slice_rows(const I* indptr, const I* indices, const I* data, const I* rows, int64_t nr, I* sub_indptr,
                    I* sub_indices, I* sub_data) {
  if (!sub_indices) {
    sub_indptr[0] = 0;
    for (int64_t i = 0; i < nr; ++i) sub_indptr[i + 1] = sub_indptr[i] + (indptr[rows[i] + 1] - indptr[rows[i]]);
    return;
  }
  parallel_for(0, nr, 1024, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      const int64_t s = indptr[rows[i]], len = indptr[rows[i] + 1] - s, o = sub_indptr[i];
      for (int64_t k = 0; k < len; ++k) {
        sub_indices[o + k] = indices[s + k];
        sub_data[o + k] = data ? data[s + k] : I(s + k);
      }
    }
  });
}

// rows[] x cols[] sub-matrix with column relabelling through col_map (size n_cols, -1 = dropped).
template <typename I>
void csr_slice_matrix(const I* indptr, const I* indices, const I* data, const I* rows, int64_t nr, const I* col_map,
                      I* sub_indptr, I* sub_indices, I* sub_data) {
  if (!sub_indices) {
    std::vector<int64_t> cnt(nr, 0);
    parallel_for(0, nr, 1024, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        int64_t c = 0;
        for (int64_t k = indptr[rows[i]]; k < indptr[rows[i] + 1]; ++k) c += col_map[indices[k]] >= 0;
        cnt[i] = c;
      }
    });
    sub_indptr[0] = 0;
    for (int64_t i = 0; i < nr; ++i) sub_indptr[i + 1] = sub_indptr[i] + I(cnt[i]);
    return;
  }
  parallel_for(0, nr, 1024, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      int64_t o = sub_indptr[i];
      for (int64_t k = indptr[rows[i]]; k < indptr[rows[i] + 1]; ++k) {
        const I c = col_map[indices[k]];
        if (c < 0) continue;
        sub_indices[o] = c;
        sub_data[o] = data ? data[k] : I(k);
        ++o;
      }
    }
  });
}

// Line graph of a graph given as COO (u, v by edge id) + out-CSR (indptr, eids): edge i -> every out-edge j of v_i.
// Count pass (lg_src == NULL) fills lg_indptr[E + 1]; fill pass writes (i, j) pairs grouped by i.
template <typename I>
void line_graph(const I* u, const I* v, int64_t E, const I* indptr, const I* csr_eids, bool backtracking, I* lg_indptr,
                I* lg_src, I* lg_dst) {
  auto out_edge = [&](int64_t pos) { return csr_eids ? (int64_t)csr_eids[pos] : pos; };
  if (!lg_src) {
    std::vector<int64_t> cnt(E);
    parallel_for(0, E, 4096, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        const int64_t s = indptr[v[i]], t = indptr[v[i] + 1];
        int64_t c = t - s;
        if (!backtracking)
          for (int64_t k = s; k < t; ++k) c -= (v[out_edge(k)] == u[i]);  // j = (v_i -> u_i) reverses i
        cnt[i] = c;
      }
    });
    lg_indptr[0] = 0;
    for (int64_t i = 0; i < E; ++i) lg_indptr[i + 1] = lg_indptr[i] + I(cnt[i]);
    return;
  }
  parallel_for(0, E, 4096, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      int64_t o = lg_indptr[i];
      for (int64_t k = indptr[v[i]]; k < indptr[v[i] + 1]; ++k) {
        const int64_t j = out_edge(k);
        if (!backtracking && v[j] == u[i]) continue;
        lg_src[o] = I(i);
        lg_dst[o] = I(j);
        ++o;
      }
    }
  });
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cpu_coo_to_csr(int idtype, const void* row, const void* col, int64_t E, int64_t n_rows, void* indptr,
                                    void* indices, void* eids) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, { coo_to_csr<I>((const I*)row, (const I*)col, E, n_rows, (I*)indptr, (I*)indices, (I*)eids); });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_transpose(int idtype, const void* indptr, const void* indices, const void* data, int64_t n_rows,
                                       int64_t n_cols, void* t_indptr, void* t_indices, void* t_data) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    csr_transpose<I>((const I*)indptr, (const I*)indices, (const I*)data, n_rows, n_cols, (I*)t_indptr, (I*)t_indices,
                     (I*)t_data);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_coalesce(int idtype, const void* row, const void* col, int64_t E, int64_t n_rows, int64_t n_cols,
                                  void* group, void* out_row, void* out_col, void* count, void* first, void* n_out) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    *(int64_t*)n_out = coalesce<I>((const I*)row, (const I*)col, E, n_rows, n_cols, (int64_t*)group, (I*)out_row,
                                   (I*)out_col, (int64_t*)count, (int64_t*)first);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_slice_rows(int idtype, const void* indptr, const void* indices, const void* data,
                                        const void* rows, int64_t nr, void* sub_indptr, void* sub_indices, void* sub_data) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    csr_slice_rows<I>((const I*)indptr, (const I*)indices, (const I*)data, (const I*)rows, nr, (I*)sub_indptr,
                      (I*)sub_indices, (I*)sub_data);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_slice_matrix(int idtype, const void* indptr, const void* indices, const void* data,
                                          const void* rows, int64_t nr, const void* col_map, void* sub_indptr,
                                          void* sub_indices, void* sub_data) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    csr_slice_matrix<I>((const I*)indptr, (const I*)indices, (const I*)data, (const I*)rows, nr, (const I*)col_map,
                        (I*)sub_indptr, (I*)sub_indices, (I*)sub_data);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_line_graph(int idtype, const void* u, const void* v, int64_t E, const void* indptr,
                                    const void* csr_eids, int backtracking, void* lg_indptr, void* lg_src, void* lg_dst) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    line_graph<I>((const I*)u, (const I*)v, E, (const I*)indptr, (const I*)csr_eids, backtracking != 0, (I*)lg_indptr,
                  (I*)lg_src, (I*)lg_dst);
  });
  DGLB_CATCH
}
