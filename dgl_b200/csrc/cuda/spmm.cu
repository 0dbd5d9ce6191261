// g-SpMM for sm_100a:  out[v, f] = reduce_{e=(u->v)} op(ufeat[u, lidx(f)], efeat[eid, ridx(f)])
//
// Two kernel families (both CSC row-parallel, fp32 accumulation for 16-bit inputs):
//  * spmm_vec     : the hot path (copy_u / u_mul_e with a per-edge scalar, sum) — a sub-warp per
//                   destination row, 4/8/16-byte vector gathers of whole feature rows, 4 neighbour
//                   rows in flight per lane (MLP), column tiles on grid.y for wide features.
//  * spmm_generic : every (op, reducer, broadcast) combination incl. arg-max/min outputs.
// Replaces the reference's SpMMCsrKernel / cuSPARSE / COO-atomic paths
// (/root/reference/src/array/cuda/spmm.cuh:413-732, spmm.cu:27-58) with one CSC family; no
// descriptor creation, no workspace allocation, no host synchronisation per call.
#include <type_traits>

#include "common.cuh"
#include "spmm_bulk.cuh"

namespace dglb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

template <int VB>
struct VecT;
template <>
struct VecT<16> {
  using type = uint4;
};
template <>
struct VecT<8> {
  using type = uint2;
};
template <>
struct VecT<4> {
  using type = uint32_t;
};

template <typename T, int VB>
struct VecOps {
  static constexpr int VE = VB / sizeof(T);
  using V = typename VecT<VB>::type;
  __device__ static __forceinline__ void fma(float* acc, const V& v, float w) {
    const T* p = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[i] = fmaf(to_f<T>(p[i]), w, acc[i]);
  }
  __device__ static __forceinline__ V pack(const float* acc, float scale) {
    V v;
    T* p = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < VE; ++i) p[i] = from_f<T>(acc[i] * scale);
    return v;
  }
};

template <typename V>
__device__ __forceinline__ V ldg_vec(const V* p) {
  return __ldg(p);
}

// sub lanes cooperate on one row; each lane owns NV vectors per column tile.
template <typename T, typename I, int VB, int NV, int HAS_W>  // HAS_W: 0 none, 1 one scalar per edge, 2 one per head
__global__ void __launch_bounds__(256) spmm_vec_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                       const I* __restrict__ eids, int64_t num_rows,
                                                       const T* __restrict__ x, const T* __restrict__ w,
                                                       T* __restrict__ out, int64_t F, int nvec, int sub,
                                                       int64_t max_deg, int wH, int vph, int zero_long) {
  using VO = VecOps<T, VB>;
  using V = typename VO::V;
  constexpr int VE = VO::VE;
  constexpr int U = 4;
  const int rows_per_block = blockDim.x / sub;
  const int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub;
  if (row >= num_rows) return;
  const int lane = threadIdx.x % sub;
  const int vbase = blockIdx.y * sub * NV + lane;  // first vector owned by this lane
  int hd[NV];  // weight column (attention head) of each owned vector; wH == 1: one scalar per edge
#pragma unroll
  for (int j = 0; j < NV; ++j) hd[j] = wH > 1 ? min(vbase + j * sub, nvec - 1) / vph : 0;
  float acc[NV][VE];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[j][i] = 0.f;
  const int64_t s = indptr[row], t = indptr[row + 1];
  if (max_deg > 0 && t - s > max_deg) {
    // long rows are handled by the edge-balanced split kernel; when it accumulates straight into `out` (fp32) the row is
    // cleared here, which saves the scratch memset and the finalize pass
    if (zero_long) {
      V* zr = reinterpret_cast<V*>(out + row * F);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = vbase + j * sub;
        if (c < nvec) zr[c] = VO::pack(acc[j], 0.f);
      }
    }
    return;
  }
  const int64_t rowvecs = F / VE;  // == nvec
  int64_t p = s;
  for (; p + U <= t; p += U) {
    int64_t u[U];
    constexpr int NW = HAS_W == 2 ? NV : 1;  // a scalar weight is shared by all vectors of the lane
    float wt[U][NW];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      u[k] = indices[p + k];
      const int64_t eidx = (eids ? (int64_t)eids[p + k] : p + k) * wH;
#pragma unroll
      for (int j = 0; j < NW; ++j) wt[k][j] = HAS_W ? to_f<T>(w[eidx + (HAS_W == 2 ? hd[j] : 0)]) : 1.f;
    }
    V v[U][NV];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const V* xr = reinterpret_cast<const V*>(x + u[k] * F);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = vbase + j * sub;
        if (c < nvec) v[k][j] = ldg_vec(xr + c);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k)
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (vbase + j * sub < nvec) VO::fma(acc[j], v[k][j], wt[k][HAS_W == 2 ? j : 0]);
  }
  for (; p < t; ++p) {
    const int64_t u = indices[p];
    const int64_t eidx = (eids ? (int64_t)eids[p] : p) * wH;
    const V* xr = reinterpret_cast<const V*>(x + u * F);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = vbase + j * sub;
      if (c < nvec) {
        V v = ldg_vec(xr + c);
        VO::fma(acc[j], v, HAS_W ? to_f<T>(w[eidx + (HAS_W == 2 ? hd[j] : 0)]) : 1.f);
      }
    }
  }
  V* orow = reinterpret_cast<V*>(out + row * F);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = vbase + j * sub;
    if (c < nvec) orow[c] = VO::pack(acc[j], 1.f);
  }
  (void)rowvecs;
}

template <typename T, typename I, int OP, int RED>
__global__ void __launch_bounds__(256) spmm_generic_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                           const I* __restrict__ eids, int64_t num_rows,
                                                           const T* __restrict__ ufeat, const T* __restrict__ efeat,
                                                           T* __restrict__ out, I* __restrict__ arg_u,
                                                           I* __restrict__ arg_e, int64_t F, Bcast bc, int sub) {
  using A = typename AccT<T>::type;
  constexpr bool use_l = OP != kCopyRhs;
  constexpr bool use_r = OP != kCopyLhs;
  const int rows_per_block = blockDim.x / sub;
  const int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub;
  if (row >= num_rows) return;
  const int lane = threadIdx.x % sub;
  const int64_t s = indptr[row], t = indptr[row + 1];
  for (int64_t f = lane; f < F; f += sub) {
    const int64_t lf = use_l ? (f / bc.ldiv) % bc.lmod : 0;
    const int64_t rf = use_r ? (f / bc.rdiv) % bc.rmod : 0;
    A acc = RED == kSum ? A(0) : (RED == kMax ? -INFINITY : INFINITY);
    I au = -1, ae = -1;
    for (int64_t p = s; p < t; ++p) {
      const I u = indices[p];
      const I e = eids ? eids[p] : (I)p;
      const A l = use_l ? ld_acc<T, A>(ufeat + (int64_t)u * bc.lmod + lf) : A(0);
      const A r = use_r ? ld_acc<T, A>(efeat + (int64_t)e * bc.rmod + rf) : A(0);
      const A val = apply_op<A, OP>(l, r);
      if (RED == kSum) {
        acc += val;
      } else if (RED == kMax ? (val > acc) : (val < acc)) {
        acc = val;
        au = u;
        ae = e;
      }
    }
    if (RED != kSum && s == t) acc = A(0);
    out[row * F + f] = st_acc<T, A>(acc);
    if (RED != kSum) {
      if (arg_u) arg_u[row * F + f] = au;
      if (arg_e) arg_e[row * F + f] = ae;
    }
  }
}

static int pow2_ge(int64_t x, int cap) {
  int p = 1;
  while (p < x && p < cap) p <<= 1;
  return p;
}

template <typename T, typename I, int VB, int HAS_W>
void launch_vec(const I* indptr, const I* indices, const I* eids, int64_t n, const T* x, const T* w, T* out, int64_t F,
                cudaStream_t st, int64_t max_deg = 0, int wH = 1, int zero_long = 0) {
  constexpr int VE = VB / sizeof(T);
  const int nvec = (int)(F / VE);
  const int vph = wH > 1 ? nvec / wH : nvec;  // vectors per head
  const int sub = pow2_ge(nvec, 32);
  const int rows_per_block = 256 / sub;
  const unsigned gx = (unsigned)((n + rows_per_block - 1) / rows_per_block);
  const int per_lane = (nvec + sub - 1) / sub;
  if (per_lane <= 1) {
    spmm_vec_kernel<T, I, VB, 1, HAS_W><<<dim3(gx, 1), 256, 0, st>>>(indptr, indices, eids, n, x, w, out, F, nvec, sub, max_deg, wH, vph, zero_long);
  } else if (per_lane <= 2) {
    spmm_vec_kernel<T, I, VB, 2, HAS_W><<<dim3(gx, 1), 256, 0, st>>>(indptr, indices, eids, n, x, w, out, F, nvec, sub, max_deg, wH, vph, zero_long);
  } else {
    const unsigned gy = (unsigned)((nvec + sub * 4 - 1) / (sub * 4));
    spmm_vec_kernel<T, I, VB, 4, HAS_W><<<dim3(gx, gy), 256, 0, st>>>(indptr, indices, eids, n, x, w, out, F, nvec, sub, max_deg, wH, vph, zero_long);
  }
}


// ---- edge-balanced path for long rows --------------------------------------------------------
// Work item = (long-row ordinal, first edge, #edges <= chunk). One sub-warp per item accumulates its
// chunk in registers and adds the partial row into an fp32 scratch row with vector red.global.add;
// spmm_split_finalize converts scratch -> out. Hub rows of power-law graphs are thereby spread over
// the whole GPU instead of serialising on one warp.
template <typename T, typename I, int VB, int NV, int HAS_W>  // HAS_W: 0 none, 1 one scalar per edge, 2 one per head
__global__ void __launch_bounds__(256) spmm_split_kernel(const I* __restrict__ indices, const I* __restrict__ eids,
                                                         const int32_t* __restrict__ item_ord,
                                                         const int64_t* __restrict__ item_start,
                                                         const int32_t* __restrict__ item_len, int64_t num_items,
                                                         const T* __restrict__ x, const T* __restrict__ w,
                                                         float* __restrict__ scratch, int64_t F, int nvec, int sub,
                                                         int wH, int vph, const int64_t* __restrict__ row_map) {
  using VO = VecOps<T, VB>;
  using V = typename VO::V;
  constexpr int VE = VO::VE;
  constexpr int U = 4;
  const int items_per_block = blockDim.x / sub;
  const int64_t item = (int64_t)blockIdx.x * items_per_block + threadIdx.x / sub;
  if (item >= num_items) return;
  const int lane = threadIdx.x % sub;
  const int vbase = blockIdx.y * sub * NV + lane;
  int hd[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) hd[j] = wH > 1 ? min(vbase + j * sub, nvec - 1) / vph : 0;
  float acc[NV][VE];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[j][i] = 0.f;
  const int64_t s = item_start[item], t = s + item_len[item];
  int64_t p = s;
  for (; p + U <= t; p += U) {
    int64_t u[U];
    constexpr int NW = HAS_W == 2 ? NV : 1;  // a scalar weight is shared by all vectors of the lane
    float wt[U][NW];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      u[k] = indices[p + k];
      const int64_t eidx = (eids ? (int64_t)eids[p + k] : p + k) * wH;
#pragma unroll
      for (int j = 0; j < NW; ++j) wt[k][j] = HAS_W ? to_f<T>(w[eidx + (HAS_W == 2 ? hd[j] : 0)]) : 1.f;
    }
    V v[U][NV];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const V* xr = reinterpret_cast<const V*>(x + u[k] * F);
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (vbase + j * sub < nvec) v[k][j] = ldg_vec(xr + vbase + j * sub);
    }
#pragma unroll
    for (int k = 0; k < U; ++k)
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (vbase + j * sub < nvec) VO::fma(acc[j], v[k][j], wt[k][HAS_W == 2 ? j : 0]);
  }
  for (; p < t; ++p) {
    const int64_t u = indices[p];
    const int64_t eidx = (eids ? (int64_t)eids[p] : p) * wH;
    const V* xr = reinterpret_cast<const V*>(x + u * F);
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (vbase + j * sub < nvec) {
        V v = ldg_vec(xr + vbase + j * sub);
        VO::fma(acc[j], v, HAS_W ? to_f<T>(w[eidx + (HAS_W == 2 ? hd[j] : 0)]) : 1.f);
      }
  }
  // row_map != nullptr: `scratch` is the fp32 output itself and row_map the long-row ids (direct accumulation)
  float* srow = scratch + (row_map ? row_map[item_ord[item]] : (int64_t)item_ord[item]) * F;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = vbase + j * sub;
    if (c < nvec) {
#pragma unroll
      for (int i = 0; i < VE; ++i) atomicAdd(srow + c * VE + i, acc[j][i]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) spmm_split_finalize_kernel(const float* __restrict__ scratch,
                                                                  const int64_t* __restrict__ long_rows,
                                                                  int64_t n_long, int64_t F, T* __restrict__ out) {
  const int64_t total = n_long * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / F, f = i - o * F;
    out[long_rows[o] * F + f] = from_f<T>(scratch[i]);
  }
}

template <typename T, typename I, int VB, int HAS_W>
void launch_split(const I* indices, const I* eids, const int32_t* item_ord, const int64_t* item_start,
                  const int32_t* item_len, int64_t num_items, const T* x, const T* w, float* scratch, int64_t F,
                  cudaStream_t st, int wH = 1, const int64_t* row_map = nullptr) {
  constexpr int VE = VB / sizeof(T);
  const int nvec = (int)(F / VE);
  const int vph = wH > 1 ? nvec / wH : nvec;
  const int sub = pow2_ge(nvec, 32);
  const int ipb = 256 / sub;
  const unsigned gx = (unsigned)((num_items + ipb - 1) / ipb);
  const int per_lane = (nvec + sub - 1) / sub;
  if (per_lane <= 1) {
    spmm_split_kernel<T, I, VB, 1, HAS_W><<<dim3(gx, 1), 256, 0, st>>>(indices, eids, item_ord, item_start, item_len, num_items, x, w, scratch, F, nvec, sub, wH, vph, row_map);
  } else if (per_lane <= 2) {
    spmm_split_kernel<T, I, VB, 2, HAS_W><<<dim3(gx, 1), 256, 0, st>>>(indices, eids, item_ord, item_start, item_len, num_items, x, w, scratch, F, nvec, sub, wH, vph, row_map);
  } else {
    const unsigned gy = (unsigned)((nvec + sub * 4 - 1) / (sub * 4));
    spmm_split_kernel<T, I, VB, 4, HAS_W><<<dim3(gx, gy), 256, 0, st>>>(indices, eids, item_ord, item_start, item_len, num_items, x, w, scratch, F, nvec, sub, wH, vph, row_map);
  }
}

static int pick_vb_spmm(int64_t row_bytes, const void* x, const void* out) {
  const uintptr_t xa = (uintptr_t)x, oa = (uintptr_t)out;
  if (row_bytes % 16 == 0 && xa % 16 == 0 && oa % 16 == 0) return 16;
  if (row_bytes % 8 == 0 && xa % 8 == 0 && oa % 8 == 0) return 8;
  if (row_bytes % 4 == 0 && xa % 4 == 0 && oa % 4 == 0) return 4;
  return 0;
}

template <typename T, typename I>
bool try_vec(int op, int red, const I* indptr, const I* indices, const I* eids, int64_t n, const T* x, const T* w,
             T* out, int64_t F, Bcast bc, cudaStream_t st) {
  if (red != kSum) return false;
  const bool copy = op == kCopyLhs;
  const bool mul_scalar = op == kMul && bc.rmod == 1 && bc.ldiv == 1 && bc.lmod == F;
  if (!copy && !mul_scalar) return false;
  if (sizeof(T) == 8) return false;
  const int64_t row_bytes = F * (int64_t)sizeof(T);
  const uintptr_t xa = (uintptr_t)x, oa = (uintptr_t)out;
  int vb = 0;
  if (row_bytes % 16 == 0 && xa % 16 == 0 && oa % 16 == 0) vb = 16;
  else if (row_bytes % 8 == 0 && xa % 8 == 0 && oa % 8 == 0) vb = 8;
  else if (row_bytes % 4 == 0 && xa % 4 == 0 && oa % 4 == 0) vb = 4;
  if (vb == 0) return false;
#define LV(VBV)                                                                          \
  if (vb == VBV) {                                                                       \
    if (copy) launch_vec<T, I, VBV, false>(indptr, indices, eids, n, x, w, out, F, st);  \
    else launch_vec<T, I, VBV, true>(indptr, indices, eids, n, x, w, out, F, st);        \
    return true;                                                                         \
  }
  LV(16) LV(8) LV(4)
#undef LV
  return false;
}

template <typename T, typename I>
void spmm_dispatch(int op, int red, const I* indptr, const I* indices, const I* eids, int64_t n, const T* u, const T* e,
                   T* out, I* au, I* ae, int64_t F, Bcast bc, cudaStream_t st) {
  if constexpr (!std::is_same<T, double>::value) {
    if (try_vec<T, I>(op, red, indptr, indices, eids, n, u, e, out, F, bc, st)) return;
  }
  const int sub = pow2_ge(F, 32);
  const int rows_per_block = 256 / sub;
  const unsigned gx = (unsigned)((n + rows_per_block - 1) / rows_per_block);
#define CASE(OPC, REDC)                                                                                              \
  if (op == OPC && red == REDC) {                                                                                    \
    spmm_generic_kernel<T, I, OPC, REDC><<<gx, 256, 0, st>>>(indptr, indices, eids, n, u, e, out, au, ae, F, bc, sub); \
    return;                                                                                                          \
  }
  CASE(kAdd, kSum) CASE(kAdd, kMax) CASE(kAdd, kMin)
  CASE(kMul, kSum) CASE(kMul, kMax) CASE(kMul, kMin)
  CASE(kSub, kSum) CASE(kSub, kMax) CASE(kSub, kMin)
  CASE(kDiv, kSum) CASE(kDiv, kMax) CASE(kDiv, kMin)
  CASE(kCopyLhs, kSum) CASE(kCopyLhs, kMax) CASE(kCopyLhs, kMin)
  CASE(kCopyRhs, kSum) CASE(kCopyRhs, kMax) CASE(kCopyRhs, kMin)
#undef CASE
  throw std::runtime_error("gspmm: bad op/reduce");
}

}  // namespace dglb

using namespace dglb;

DGLB_API const char* dglb200_last_error() { return dglb::get_last_error(); }

DGLB_API int dglb200_cuda_gspmm(int op, int red, int dtype, int idtype, const void* indptr, const void* indices,
                                const void* eids, int64_t num_rows, const void* ufeat, const void* efeat, void* out,
                                void* arg_u, void* arg_e, int64_t F, int64_t ldiv, int64_t lmod, int64_t rdiv,
                                int64_t rmod, void* stream) {
  DGLB_TRY
  Bcast bc{ldiv, lmod, rdiv, rmod};
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_FLOAT_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, {
      spmm_dispatch<T, I>(op, red, (const I*)indptr, (const I*)indices, (const I*)eids, num_rows, (const T*)ufeat,
                          (const T*)efeat, (T*)out, (I*)arg_u, (I*)arg_e, F, bc, st);
    });
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// Load-balanced g-SpMM (copy_u / u_mul_e with per-edge scalar, sum): rows with degree <= max_deg go through the
// row-parallel kernel, longer rows through edge-balanced work items (see spmm_split_kernel).
DGLB_API int dglb200_cuda_gspmm_balanced(int has_w, int dtype, int idtype, const void* indptr, const void* indices,
                                         const void* eids, int64_t num_rows, const void* x, const void* w, void* out,
                                         int64_t F, int64_t max_deg, const void* item_ord, const void* item_start,
                                         const void* item_len, int64_t num_items, const void* long_rows,
                                         int64_t n_long, void* scratch, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  auto run = [&](auto tag, auto itag) {
    using T = decltype(tag);
    using I = decltype(itag);
    // has_w = number of weight columns per edge (0: copy_u, 1: scalar, H > 1: one per head of F/H features);
    // a vector must not straddle two heads, so the vector width has to divide the per-head byte count
    const int wH = has_w > 1 ? has_w : 1;
    DGLB_CHECK(F % wH == 0, "feature size must be a multiple of the number of weight columns");
    const int vb = pick_vb_spmm((F / wH) * (int64_t)sizeof(T), x, out);
    DGLB_CHECK(vb != 0 && (F * (int64_t)sizeof(T)) % vb == 0, "feature rows must be 4-byte aligned for the balanced g-SpMM");
#define RUN(VBV, HW)                                                                                              \
    /* fp32 outputs: the split kernel adds straight into `out` (rows cleared by the row kernel) - 2 launches instead */ \
    /* of 4; 16-bit outputs accumulate in an fp32 scratch that a finalize pass converts                            */ \
    const bool direct = std::is_same<T, float>::value && num_items > 0;                                             \
    if (direct || !(wH == 1 && bulk::launch<T, I, (HW != 0)>((const I*)indptr, (const I*)indices, (const I*)eids,    \
                                                             num_rows, (const T*)x, (const T*)w, (T*)out, F, max_deg, st))) \
      launch_vec<T, I, VBV, HW>((const I*)indptr, (const I*)indices, (const I*)eids, num_rows, (const T*)x,         \
                                (const T*)w, (T*)out, F, st, max_deg, wH, direct ? 1 : 0);                          \
    if (direct) {                                                                                                   \
      launch_split<T, I, VBV, HW>((const I*)indices, (const I*)eids, (const int32_t*)item_ord,                      \
                                  (const int64_t*)item_start, (const int32_t*)item_len, num_items, (const T*)x,     \
                                  (const T*)w, (float*)out, F, st, wH, (const int64_t*)long_rows);                  \
    } else if (num_items > 0) {                                                                                     \
      DGLB_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)n_long * F * sizeof(float), st));                         \
      launch_split<T, I, VBV, HW>((const I*)indices, (const I*)eids, (const int32_t*)item_ord,                      \
                                  (const int64_t*)item_start, (const int32_t*)item_len, num_items, (const T*)x,     \
                                  (const T*)w, (float*)scratch, F, st, wH);                                         \
      int64_t blocks = (n_long * F + 255) / 256;                                                                    \
      if (blocks > 148 * 16) blocks = 148 * 16;                                                                     \
      spmm_split_finalize_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const float*)scratch,                        \
                                                                      (const int64_t*)long_rows, n_long, F, (T*)out); \
    }
    const int hw = has_w == 0 ? 0 : (wH > 1 ? 2 : 1);
    if (vb == 16) { if (hw == 2) { RUN(16, 2) } else if (hw == 1) { RUN(16, 1) } else { RUN(16, 0) } }
    else if (vb == 8) { if (hw == 2) { RUN(8, 2) } else if (hw == 1) { RUN(8, 1) } else { RUN(8, 0) } }
    else { if (hw == 2) { RUN(4, 2) } else if (hw == 1) { RUN(4, 1) } else { RUN(4, 0) } }
#undef RUN
  };
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kF32) run(float(), I());
    else if (dtype == kBF16) run(__nv_bfloat16(), I());
    else if (dtype == kF16) run(__half(), I());
    else throw std::runtime_error("balanced gspmm: unsupported dtype");
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
