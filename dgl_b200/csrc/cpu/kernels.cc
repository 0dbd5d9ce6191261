// CPU g-SpMM / g-SDDMM / segment-reduce / fused edge-softmax.
// Semantics follow SURVEY.md Appendix A (spec extracted from
// /root/reference/python/dgl/_sparse_ops.py:156-566 and
// /root/reference/src/array/cpu/{spmm,sddmm,segment_reduce}.h) but the
// implementation is a fresh row-parallel design: operand broadcasting is
// expressed as idx = (f / div) % mod per operand (mod == operand length)
// instead of the reference's materialised offset tables (src/bcast.cc).
#include "common.h"

namespace dglb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

struct Bcast {
  int64_t ldiv, lmod, rdiv, rmod;
};

template <typename A, int OP>
inline A apply_op(A l, A r) {
  if (OP == kAdd) return l + r;
  if (OP == kSub) return l - r;
  if (OP == kMul) return l * r;
  if (OP == kDiv) return l / r;
  if (OP == kCopyLhs) return l;
  if (OP == kCopyRhs) return r;
  return l * r;
}

template <typename A>
inline A apply_op_rt(int op, A l, A r) {
  switch (op) {
    case kAdd: return l + r;
    case kSub: return l - r;
    case kMul: return l * r;
    case kDiv: return l / r;
    case kCopyLhs: return l;
    case kCopyRhs: return r;
    default: return l * r;
  }
}

template <typename T, typename A>
inline A ld(const T* p, int64_t i) {
  return A(float(p[i]));
}
template <>
inline double ld<double, double>(const double* p, int64_t i) {
  return p[i];
}
template <typename T, typename A>
inline T st(A v) {
  return T(float(v));
}
template <>
inline double st<double, double>(double v) {
  return v;
}

// out[v, f] = reduce_{e=(u->v)} op(ufeat[u, lidx(f)], efeat[eid, ridx(f)])
template <typename T, typename I, int OP, int RED>
void gspmm_csr(const I* indptr, const I* indices, const I* eids, int64_t num_rows, const T* ufeat,
               const T* efeat, T* out, I* arg_u, I* arg_e, int64_t F, Bcast bc) {
  using A = typename AccT<T>::type;
  constexpr bool use_l = OP != kCopyRhs;
  constexpr bool use_r = OP != kCopyLhs;
  const bool simple = (!use_l || (bc.ldiv == 1 && bc.lmod == F)) && (!use_r || (bc.rdiv == 1 && bc.rmod == F));
  parallel_for_dynamic(0, num_rows, 128, [&](int64_t rb, int64_t re) {
    std::vector<A> acc(F);
    std::vector<I> au(RED == kSum ? 0 : F), ae(RED == kSum ? 0 : F);
    for (int64_t v = rb; v < re; ++v) {
      const I s = indptr[v], t = indptr[v + 1];
      const A init = RED == kSum ? A(0)
                                 : (RED == kMax ? -std::numeric_limits<A>::infinity()
                                                : std::numeric_limits<A>::infinity());
      for (int64_t f = 0; f < F; ++f) acc[f] = init;
      if (RED != kSum)
        for (int64_t f = 0; f < F; ++f) au[f] = ae[f] = I(-1);
      for (I p = s; p < t; ++p) {
        const I u = indices[p];
        const I e = eids ? eids[p] : p;
        const T* lp = use_l ? ufeat + int64_t(u) * bc.lmod : nullptr;
        const T* rp = use_r ? efeat + int64_t(e) * bc.rmod : nullptr;
        for (int64_t f = 0; f < F; ++f) {
          const A l = use_l ? ld<T, A>(lp, simple ? f : (f / bc.ldiv) % bc.lmod) : A(0);
          const A r = use_r ? ld<T, A>(rp, simple ? f : (f / bc.rdiv) % bc.rmod) : A(0);
          const A val = apply_op<A, OP>(l, r);
          if (RED == kSum) {
            acc[f] += val;
          } else if (RED == kMax ? (val > acc[f]) : (val < acc[f])) {
            acc[f] = val;
            au[f] = u;
            ae[f] = e;
          }
        }
      }
      T* op = out + v * F;
      if (RED == kSum) {
        for (int64_t f = 0; f < F; ++f) op[f] = st<T, A>(acc[f]);
      } else {
        for (int64_t f = 0; f < F; ++f) {
          // zero-degree rows produce 0 (reference: python/dgl/ops/spmm.py:119-133)
          op[f] = (s == t) ? st<T, A>(A(0)) : st<T, A>(acc[f]);
          if (arg_u) arg_u[v * F + f] = au[f];
          if (arg_e) arg_e[v * F + f] = ae[f];
        }
      }
    }
  });
}

template <typename T, typename I>
void gspmm_dispatch(int op, int red, const I* indptr, const I* indices, const I* eids, int64_t n,
                    const T* u, const T* e, T* out, I* au, I* ae, int64_t F, Bcast bc) {
#define CASE(OPC, REDC)                                                             \
  if (op == OPC && red == REDC) {                                                   \
    gspmm_csr<T, I, OPC, REDC>(indptr, indices, eids, n, u, e, out, au, ae, F, bc); \
    return;                                                                         \
  }
  CASE(kAdd, kSum) CASE(kAdd, kMax) CASE(kAdd, kMin)
  CASE(kSub, kSum) CASE(kSub, kMax) CASE(kSub, kMin)
  CASE(kMul, kSum) CASE(kMul, kMax) CASE(kMul, kMin)
  CASE(kDiv, kSum) CASE(kDiv, kMax) CASE(kDiv, kMin)
  CASE(kCopyLhs, kSum) CASE(kCopyLhs, kMax) CASE(kCopyLhs, kMin)
  CASE(kCopyRhs, kSum) CASE(kCopyRhs, kMax) CASE(kCopyRhs, kMin)
#undef CASE
  throw std::runtime_error("gspmm: bad op/reduce");
}

// out[e, f] = op(lhs[sel_l(e)], rhs[sel_r(e)]); `dot` reduces the trailing rsz elements.
// target codes: 0 = u (row/src), 1 = e (edge id), 2 = v (col/dst).
template <typename T, typename I>
void gsddmm_coo(int op, const I* row, const I* col, int64_t E, const T* lhs, const T* rhs, T* out,
                int lt, int rt, int64_t F, int64_t rsz, Bcast bc) {
  using A = typename AccT<T>::type;
  const bool use_l = op != kCopyRhs, use_r = op != kCopyLhs;
  parallel_for(0, E, 512, [&](int64_t b, int64_t e_) {
    for (int64_t e = b; e < e_; ++e) {
      const int64_t li = lt == 0 ? int64_t(row[e]) : (lt == 1 ? e : int64_t(col[e]));
      const int64_t ri = rt == 0 ? int64_t(row[e]) : (rt == 1 ? e : int64_t(col[e]));
      const T* lp = use_l ? lhs + li * bc.lmod * rsz : nullptr;
      const T* rp = use_r ? rhs + ri * bc.rmod * rsz : nullptr;
      T* op_ = out + e * F;
      for (int64_t f = 0; f < F; ++f) {
        const int64_t lf = use_l ? (f / bc.ldiv) % bc.lmod : 0;
        const int64_t rf = use_r ? (f / bc.rdiv) % bc.rmod : 0;
        if (op == kDot) {
          A acc = 0;
          for (int64_t k = 0; k < rsz; ++k) acc += ld<T, A>(lp, lf * rsz + k) * ld<T, A>(rp, rf * rsz + k);
          op_[f] = st<T, A>(acc);
        } else {
          const A l = use_l ? ld<T, A>(lp, lf) : A(0);
          const A r = use_r ? ld<T, A>(rp, rf) : A(0);
          op_[f] = st<T, A>(apply_op_rt<A>(op, l, r));
        }
      }
    }
  });
}

// segment reduce over contiguous segments: offsets[nseg+1]
template <typename T, typename I>
void segment_reduce(int red, const I* offsets, int64_t nseg, const T* x, T* out, I* arg, int64_t F) {
  using A = typename AccT<T>::type;
  parallel_for_dynamic(0, nseg, 64, [&](int64_t b, int64_t e_) {
    std::vector<A> acc(F);
    std::vector<I> ar(F);
    for (int64_t s = b; s < e_; ++s) {
      const int64_t lo = offsets[s], hi = offsets[s + 1];
      const A init = red == kSum ? A(0)
                                 : (red == kMax ? -std::numeric_limits<A>::infinity()
                                                : std::numeric_limits<A>::infinity());
      for (int64_t f = 0; f < F; ++f) { acc[f] = init; ar[f] = I(-1); }
      for (int64_t r = lo; r < hi; ++r) {
        const T* xp = x + r * F;
        for (int64_t f = 0; f < F; ++f) {
          const A v = ld<T, A>(xp, f);
          if (red == kSum) acc[f] += v;
          else if (red == kMax ? v > acc[f] : v < acc[f]) { acc[f] = v; ar[f] = I(r); }
        }
      }
      for (int64_t f = 0; f < F; ++f) {
        out[s * F + f] = (red != kSum && lo == hi) ? st<T, A>(A(0)) : st<T, A>(acc[f]);
        if (arg && red != kSum) arg[s * F + f] = ar[f];
      }
    }
  });
}

// Fused edge-softmax over the in-edges of every dst row (CSC order), F = #heads.
// score/out are indexed by edge id (eids[p] or p).
template <typename T, typename I>
void edge_softmax_fwd(const I* indptr, const I* eids, int64_t n, const T* score, T* out, int64_t F) {
  using A = typename AccT<T>::type;
  parallel_for_dynamic(0, n, 128, [&](int64_t b, int64_t e_) {
    for (int64_t v = b; v < e_; ++v) {
      const I s = indptr[v], t = indptr[v + 1];
      for (int64_t f = 0; f < F; ++f) {
        A mx = -std::numeric_limits<A>::infinity();
        for (I p = s; p < t; ++p) mx = std::max(mx, ld<T, A>(score, int64_t(eids ? eids[p] : p) * F + f));
        A sum = 0;
        for (I p = s; p < t; ++p) sum += std::exp(ld<T, A>(score, int64_t(eids ? eids[p] : p) * F + f) - mx);
        const A inv = sum > 0 ? A(1) / sum : A(0);
        for (I p = s; p < t; ++p) {
          const int64_t e = int64_t(eids ? eids[p] : p) * F + f;
          out[e] = st<T, A>(std::exp(ld<T, A>(score, e) - mx) * inv);
        }
      }
    }
  });
}

// grad_score = out * (grad_out - sum_in_edges(out * grad_out))
template <typename T, typename I>
void edge_softmax_bwd(const I* indptr, const I* eids, int64_t n, const T* out, const T* gout, T* gscore, int64_t F) {
  using A = typename AccT<T>::type;
  parallel_for_dynamic(0, n, 128, [&](int64_t b, int64_t e_) {
    for (int64_t v = b; v < e_; ++v) {
      const I s = indptr[v], t = indptr[v + 1];
      for (int64_t f = 0; f < F; ++f) {
        A acc = 0;
        for (I p = s; p < t; ++p) {
          const int64_t e = int64_t(eids ? eids[p] : p) * F + f;
          acc += ld<T, A>(out, e) * ld<T, A>(gout, e);
        }
        for (I p = s; p < t; ++p) {
          const int64_t e = int64_t(eids ? eids[p] : p) * F + f;
          const A o = ld<T, A>(out, e);
          gscore[e] = st<T, A>(o * (ld<T, A>(gout, e) - acc));
        }
      }
    }
  });
}

}  // namespace dglb

using namespace dglb;

DGLB_API const char* dglb200_last_error() { return dglb::get_last_error(); }

DGLB_API int dglb200_cpu_num_threads() {
  const int t = dglb::team_size_global().load(std::memory_order_relaxed);
  return t > 0 ? t : omp_get_max_threads();
}
// process-wide: every entry point re-applies it to its calling thread (common.h sync_team)
DGLB_API int dglb200_cpu_set_num_threads(int n) {
  if (n < 1) n = 1;
  dglb::team_size_global().store(n, std::memory_order_relaxed);
  omp_set_num_threads(n);
  return 0;
}

// calling-thread override of the team size (0 = follow the process-wide value)
DGLB_API int dglb200_cpu_set_thread_team(int n) {
  dglb::team_size_thread() = n < 0 ? 0 : n;
  return 0;
}

DGLB_API int dglb200_cpu_gspmm(int op, int red, int dtype, int idtype, const void* indptr, const void* indices,
                               const void* eids, int64_t num_rows, const void* ufeat, const void* efeat, void* out,
                               void* arg_u, void* arg_e, int64_t F, int64_t ldiv, int64_t lmod, int64_t rdiv,
                               int64_t rmod) {
  DGLB_TRY
  Bcast bc{ldiv, lmod, rdiv, rmod};
  DGLB_DTYPE_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, {
      gspmm_dispatch<T, I>(op, red, (const I*)indptr, (const I*)indices, (const I*)eids, num_rows, (const T*)ufeat,
                           (const T*)efeat, (T*)out, (I*)arg_u, (I*)arg_e, F, bc);
    });
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_gsddmm(int op, int dtype, int idtype, const void* row, const void* col, int64_t E,
                                const void* lhs, const void* rhs, void* out, int lhs_target, int rhs_target,
                                int64_t F, int64_t rsz, int64_t ldiv, int64_t lmod, int64_t rdiv, int64_t rmod) {
  DGLB_TRY
  Bcast bc{ldiv, lmod, rdiv, rmod};
  DGLB_DTYPE_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, {
      gsddmm_coo<T, I>(op, (const I*)row, (const I*)col, E, (const T*)lhs, (const T*)rhs, (T*)out, lhs_target,
                       rhs_target, F, rsz, bc);
    });
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_segment_reduce(int red, int dtype, int idtype, const void* offsets, int64_t nseg,
                                        const void* x, void* out, void* arg, int64_t F) {
  DGLB_TRY
  DGLB_DTYPE_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, { segment_reduce<T, I>(red, (const I*)offsets, nseg, (const T*)x, (T*)out, (I*)arg, F); });
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_edge_softmax_fwd(int dtype, int idtype, const void* indptr, const void* eids, int64_t n,
                                          const void* score, void* out, int64_t F) {
  DGLB_TRY
  DGLB_DTYPE_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, { edge_softmax_fwd<T, I>((const I*)indptr, (const I*)eids, n, (const T*)score, (T*)out, F); });
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_edge_softmax_bwd(int dtype, int idtype, const void* indptr, const void* eids, int64_t n,
                                          const void* out, const void* gout, void* gscore, int64_t F) {
  DGLB_TRY
  DGLB_DTYPE_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, {
      edge_softmax_bwd<T, I>((const I*)indptr, (const I*)eids, n, (const T*)out, (const T*)gout, (T*)gscore, F);
    });
  });
  DGLB_CATCH
}
