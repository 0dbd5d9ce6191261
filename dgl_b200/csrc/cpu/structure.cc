// Sparse-structure arithmetic on the host: CSR x CSR (Gustavson, two passes so the output is
// allocated exactly once), weighted CSR + CSR, and the masked row-dot that the SpGEMM backward
// needs.  All outputs have column-sorted rows, which is what lets the row-dot be a linear merge.
//
// Behavioural reference: src/array/cpu/csr_mm.cc, src/array/cpu/csr_sum.cc (reference tree);
// the design here (symbolic pass + numeric pass over thread-private dense accumulators,
// dynamic row chunks for skewed graphs) is our own.
#include <algorithm>
#include <vector>

#include "common.h"

namespace dglb {
namespace {

// ---- C = A @ B, symbolic: nnz of every output row ------------------------------------------
template <typename I>
void csr_mm_count(int64_t n_rows, const I* ap, const I* ai, const I* bp, const I* bi, int64_t n_cols,
                  int64_t* row_nnz) {
  parallel_for_dynamic(0, n_rows, 256, [&](int64_t b, int64_t e) {
    // marker[c] == tag of the current row means column c was already seen
    thread_local std::vector<int64_t> marker;
    if ((int64_t)marker.size() < n_cols) marker.assign(n_cols, 0);
    // every processed row draws a fresh tag from a per-thread counter, so marks left by earlier
    // rows, chunks or calls can never be mistaken for the current row
    thread_local int64_t epoch = 0;
    for (int64_t r = b; r < e; ++r) {
      const int64_t tag = ++epoch;
      int64_t cnt = 0;
      for (I x = ap[r]; x < ap[r + 1]; ++x) {
        const I k = ai[x];
        for (I y = bp[k]; y < bp[k + 1]; ++y) {
          const I c = bi[y];
          if (marker[c] != tag) {
            marker[c] = tag;
            ++cnt;
          }
        }
      }
      row_nnz[r] = cnt;
    }
  });
}

// ---- numeric pass -----------------------------------------------------------------------------
template <typename I, typename T>
void csr_mm_fill(int64_t n_rows, const I* ap, const I* ai, const T* av, const I* bp, const I* bi, const T* bv,
                 int64_t n_cols, const int64_t* cp, I* ci, T* cv) {
  using A = typename AccT<T>::type;
  parallel_for_dynamic(0, n_rows, 256, [&](int64_t b, int64_t e) {
    thread_local std::vector<int64_t> marker;
    thread_local std::vector<double> acc;  // wide enough for every AccT
    thread_local int64_t epoch = 0;
    if ((int64_t)marker.size() < n_cols) {
      marker.assign(n_cols, 0);
      acc.assign(n_cols, 0.0);
    }
    for (int64_t r = b; r < e; ++r) {
      const int64_t tag = ++epoch;
      I* cols = ci + cp[r];
      int64_t cnt = 0;
      for (I x = ap[r]; x < ap[r + 1]; ++x) {
        const I k = ai[x];
        const A a = av ? (A)(float)av[x] : (A)1;
        for (I y = bp[k]; y < bp[k + 1]; ++y) {
          const I c = bi[y];
          const A prod = a * (bv ? (A)(float)bv[y] : (A)1);
          if (marker[c] != tag) {
            marker[c] = tag;
            acc[c] = prod;
            cols[cnt++] = c;
          } else {
            acc[c] += prod;
          }
        }
      }
      std::sort(cols, cols + cnt);
      T* vals = cv + cp[r];
      for (int64_t j = 0; j < cnt; ++j) vals[j] = T((float)acc[cols[j]]);
    }
  });
}

template <typename I>
void csr_mm_fill_f64(int64_t n_rows, const I* ap, const I* ai, const double* av, const I* bp, const I* bi,
                     const double* bv, int64_t n_cols, const int64_t* cp, I* ci, double* cv) {
  parallel_for_dynamic(0, n_rows, 256, [&](int64_t b, int64_t e) {
    thread_local std::vector<int64_t> marker;
    thread_local std::vector<double> acc;
    thread_local int64_t epoch = 0;
    if ((int64_t)marker.size() < n_cols) {
      marker.assign(n_cols, 0);
      acc.assign(n_cols, 0.0);
    }
    for (int64_t r = b; r < e; ++r) {
      const int64_t tag = ++epoch;
      I* cols = ci + cp[r];
      int64_t cnt = 0;
      for (I x = ap[r]; x < ap[r + 1]; ++x) {
        const I k = ai[x];
        const double a = av ? av[x] : 1.0;
        for (I y = bp[k]; y < bp[k + 1]; ++y) {
          const I c = bi[y];
          const double prod = a * (bv ? bv[y] : 1.0);
          if (marker[c] != tag) {
            marker[c] = tag;
            acc[c] = prod;
            cols[cnt++] = c;
          } else {
            acc[c] += prod;
          }
        }
      }
      std::sort(cols, cols + cnt);
      double* vals = cv + cp[r];
      for (int64_t j = 0; j < cnt; ++j) vals[j] = acc[cols[j]];
    }
  });
}

// ---- C = alpha*A + beta*B on sorted rows: a two-pointer merge -------------------------------
template <typename I>
void csr_sum_count(int64_t n_rows, const I* ap, const I* ai, const I* bp, const I* bi, int64_t* row_nnz) {
  parallel_for(0, n_rows, 1024, [&](int64_t b, int64_t e) {
    for (int64_t r = b; r < e; ++r) {
      I x = ap[r], xe = ap[r + 1], y = bp[r], ye = bp[r + 1];
      int64_t cnt = 0;
      while (x < xe && y < ye) {
        const I ca = ai[x], cb = bi[y];
        x += ca <= cb;
        y += cb <= ca;
        ++cnt;
      }
      row_nnz[r] = cnt + (xe - x) + (ye - y);
    }
  });
}

template <typename I, typename T>
void csr_sum_fill(int64_t n_rows, const I* ap, const I* ai, const T* av, double alpha, const I* bp, const I* bi,
                  const T* bv, double beta, const int64_t* cp, I* ci, T* cv, I* src_a, I* src_b) {
  // src_a / src_b (optional): for every output nnz the position of the contributing input nnz or
  // -1; the autograd wrapper turns those into the two value gradients with one gather each.
  parallel_for(0, n_rows, 1024, [&](int64_t b, int64_t e) {
    for (int64_t r = b; r < e; ++r) {
      I x = ap[r], xe = ap[r + 1], y = bp[r], ye = bp[r + 1];
      int64_t o = cp[r];
      while (x < xe || y < ye) {
        const bool take_a = y >= ye || (x < xe && ai[x] <= bi[y]);
        const bool take_b = x >= xe || (y < ye && bi[y] <= ai[x]);
        double v = 0;
        if (take_a) v += alpha * (av ? (double)(float)av[x] : 1.0);
        if (take_b) v += beta * (bv ? (double)(float)bv[y] : 1.0);
        ci[o] = take_a ? ai[x] : bi[y];
        cv[o] = T((float)v);
        if (src_a) src_a[o] = take_a ? x : (I)-1;
        if (src_b) src_b[o] = take_b ? y : (I)-1;
        x += take_a;
        y += take_b;
        ++o;
      }
    }
  });
}

template <typename I>
void csr_sum_fill_f64(int64_t n_rows, const I* ap, const I* ai, const double* av, double alpha, const I* bp,
                      const I* bi, const double* bv, double beta, const int64_t* cp, I* ci, double* cv, I* src_a,
                      I* src_b) {
  parallel_for(0, n_rows, 1024, [&](int64_t b, int64_t e) {
    for (int64_t r = b; r < e; ++r) {
      I x = ap[r], xe = ap[r + 1], y = bp[r], ye = bp[r + 1];
      int64_t o = cp[r];
      while (x < xe || y < ye) {
        const bool take_a = y >= ye || (x < xe && ai[x] <= bi[y]);
        const bool take_b = x >= xe || (y < ye && bi[y] <= ai[x]);
        double v = 0;
        if (take_a) v += alpha * (av ? av[x] : 1.0);
        if (take_b) v += beta * (bv ? bv[y] : 1.0);
        ci[o] = take_a ? ai[x] : bi[y];
        cv[o] = v;
        if (src_a) src_a[o] = take_a ? x : (I)-1;
        if (src_b) src_b[o] = take_b ? y : (I)-1;
        x += take_a;
        y += take_b;
        ++o;
      }
    }
  });
}

// ---- out[e] = <X[rows[e], :], Y[cols[e], :]> on sorted sparse rows ---------------------------
// Both value gradients of C = A @ B are this shape:
//   dA[i,k] = <dC[i,:], B[k,:]>          dB[k,j] = <A^T[k,:], dC^T[j,:]>
template <typename I, typename T, typename Acc>
void csr_rowdot(int64_t nnz, const I* rows, const I* cols, const I* xp, const I* xi, const T* xv, const I* yp,
                const I* yi, const T* yv, T* out) {
  parallel_for_dynamic(0, nnz, 4096, [&](int64_t b, int64_t e) {
    for (int64_t t = b; t < e; ++t) {
      I x = xp[rows[t]], xe = xp[rows[t] + 1], y = yp[cols[t]], ye = yp[cols[t] + 1];
      Acc s = 0;
      // gallop when the two rows have very different lengths
      const int64_t lx = xe - x, ly = ye - y;
      if (lx * 16 < ly || ly * 16 < lx) {
        const bool x_short = lx < ly;
        const I* si = x_short ? xi : yi;
        const I* li = x_short ? yi : xi;
        I s0 = x_short ? x : y, s1 = x_short ? xe : ye, l0 = x_short ? y : x, l1 = x_short ? ye : xe;
        for (I p = s0; p < s1; ++p) {
          const I* hit = std::lower_bound(li + l0, li + l1, si[p]);
          l0 = (I)(hit - li);
          if (l0 < l1 && *hit == si[p]) {
            const I px = x_short ? p : l0, py = x_short ? l0 : p;
            s += (Acc)(float)(xv ? xv[px] : T(1.f)) * (Acc)(float)(yv ? yv[py] : T(1.f));
          }
        }
      } else {
        while (x < xe && y < ye) {
          const I cx = xi[x], cy = yi[y];
          if (cx == cy) s += (Acc)(float)(xv ? xv[x] : T(1.f)) * (Acc)(float)(yv ? yv[y] : T(1.f));
          x += cx <= cy;
          y += cy <= cx;
        }
      }
      out[t] = T((float)s);
    }
  });
}

template <typename I>
void csr_rowdot_f64(int64_t nnz, const I* rows, const I* cols, const I* xp, const I* xi, const double* xv,
                    const I* yp, const I* yi, const double* yv, double* out) {
  parallel_for_dynamic(0, nnz, 4096, [&](int64_t b, int64_t e) {
    for (int64_t t = b; t < e; ++t) {
      I x = xp[rows[t]], xe = xp[rows[t] + 1], y = yp[cols[t]], ye = yp[cols[t] + 1];
      double s = 0;
      while (x < xe && y < ye) {
        const I cx = xi[x], cy = yi[y];
        if (cx == cy) s += (xv ? xv[x] : 1.0) * (yv ? yv[y] : 1.0);
        x += cx <= cy;
        y += cy <= cx;
      }
      out[t] = s;
    }
  });
}

// ---- sort every CSR row by column, carrying a permutation ------------------------------------
template <typename I>
void csr_sort_rows(int64_t n_rows, const I* indptr, I* indices, int64_t* perm) {
  parallel_for_dynamic(0, n_rows, 512, [&](int64_t b, int64_t e) {
    std::vector<std::pair<I, int64_t>> tmp;
    for (int64_t r = b; r < e; ++r) {
      const int64_t lo = indptr[r], hi = indptr[r + 1];
      if (std::is_sorted(indices + lo, indices + hi)) continue;
      tmp.resize(hi - lo);
      for (int64_t j = lo; j < hi; ++j) tmp[j - lo] = {indices[j], perm[j]};
      std::stable_sort(tmp.begin(), tmp.end(),
                       [](const std::pair<I, int64_t>& a, const std::pair<I, int64_t>& b) { return a.first < b.first; });
      for (int64_t j = lo; j < hi; ++j) {
        indices[j] = tmp[j - lo].first;
        perm[j] = tmp[j - lo].second;
      }
    }
  });
}

}  // namespace
}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cpu_csr_mm_count(int idtype, int64_t n_rows, const void* ap, const void* ai, const void* bp,
                                      const void* bi, int64_t n_cols, int64_t* row_nnz) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, csr_mm_count<I>(n_rows, (const I*)ap, (const I*)ai, (const I*)bp, (const I*)bi, n_cols,
                                            row_nnz));
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_mm_fill(int idtype, int dtype, int64_t n_rows, const void* ap, const void* ai,
                                     const void* av, const void* bp, const void* bi, const void* bv, int64_t n_cols,
                                     const int64_t* cp, void* ci, void* cv) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kF64) {
      csr_mm_fill_f64<I>(n_rows, (const I*)ap, (const I*)ai, (const double*)av, (const I*)bp, (const I*)bi,
                         (const double*)bv, n_cols, cp, (I*)ci, (double*)cv);
    } else {
      DGLB_DTYPE_SWITCH(dtype, T,
                        csr_mm_fill<I, T>(n_rows, (const I*)ap, (const I*)ai, (const T*)av, (const I*)bp,
                                          (const I*)bi, (const T*)bv, n_cols, cp, (I*)ci, (T*)cv));
    }
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_sum_count(int idtype, int64_t n_rows, const void* ap, const void* ai, const void* bp,
                                       const void* bi, int64_t* row_nnz) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, csr_sum_count<I>(n_rows, (const I*)ap, (const I*)ai, (const I*)bp, (const I*)bi, row_nnz));
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_sum_fill(int idtype, int dtype, int64_t n_rows, const void* ap, const void* ai,
                                      const void* av, double alpha, const void* bp, const void* bi, const void* bv,
                                      double beta, const int64_t* cp, void* ci, void* cv, void* src_a, void* src_b) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kF64) {
      csr_sum_fill_f64<I>(n_rows, (const I*)ap, (const I*)ai, (const double*)av, alpha, (const I*)bp, (const I*)bi,
                          (const double*)bv, beta, cp, (I*)ci, (double*)cv, (I*)src_a, (I*)src_b);
    } else {
      DGLB_DTYPE_SWITCH(dtype, T,
                        csr_sum_fill<I, T>(n_rows, (const I*)ap, (const I*)ai, (const T*)av, alpha, (const I*)bp,
                                           (const I*)bi, (const T*)bv, beta, cp, (I*)ci, (T*)cv, (I*)src_a,
                                           (I*)src_b));
    }
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_rowdot(int idtype, int dtype, int64_t nnz, const void* rows, const void* cols,
                                    const void* xp, const void* xi, const void* xv, const void* yp, const void* yi,
                                    const void* yv, void* out) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kF64) {
      csr_rowdot_f64<I>(nnz, (const I*)rows, (const I*)cols, (const I*)xp, (const I*)xi, (const double*)xv,
                        (const I*)yp, (const I*)yi, (const double*)yv, (double*)out);
    } else {
      DGLB_DTYPE_SWITCH(dtype, T,
                        (csr_rowdot<I, T, float>(nnz, (const I*)rows, (const I*)cols, (const I*)xp, (const I*)xi,
                                                 (const T*)xv, (const I*)yp, (const I*)yi, (const T*)yv, (T*)out)));
    }
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_csr_sort_rows(int idtype, int64_t n_rows, const void* indptr, void* indices, int64_t* perm) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, csr_sort_rows<I>(n_rows, (const I*)indptr, (I*)indices, perm));
  DGLB_CATCH
}
