// g-SDDMM for sm_100a:  out[e, f] = op(lhs[sel_l(e)], rhs[sel_r(e)]), `dot` reduces the trailing
// rsz elements. COO, edge-parallel. Three kernels:
//  * sddmm_dot_kernel  : sub-warp per (edge, head) with vectorised loads + shuffle tree
//                        (u_dot_v attention scores; also dY of u_mul_e_sum).
//  * sddmm_elem_kernel : thread per (edge, f) for add/sub/mul/div/copy, coalesced over f.
// Replaces /root/reference/src/array/cuda/sddmm.cuh:100-318 (SDDMMCoo*, SDDMMCsr binary search).
#include "common.cuh"

namespace dglb {

template <typename I>
__device__ __forceinline__ int64_t sel(int target, const I* row, const I* col, int64_t e) {
  return target == 0 ? (int64_t)row[e] : (target == 1 ? e : (int64_t)col[e]);
}

template <typename T, typename I>
__global__ void __launch_bounds__(256) sddmm_elem_kernel(int op, const I* __restrict__ row, const I* __restrict__ col,
                                                         int64_t E, const T* __restrict__ lhs,
                                                         const T* __restrict__ rhs, T* __restrict__ out, int lt, int rt,
                                                         int64_t F, Bcast bc) {
  using A = typename AccT<T>::type;
  const int64_t total = E * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / F, f = i - e * F;
    A l = 0, r = 0;
    if (op != kCopyRhs) l = ld_acc<T, A>(lhs + sel<I>(lt, row, col, e) * bc.lmod + (f / bc.ldiv) % bc.lmod);
    if (op != kCopyLhs) r = ld_acc<T, A>(rhs + sel<I>(rt, row, col, e) * bc.rmod + (f / bc.rdiv) % bc.rmod);
    A v;
    switch (op) {
      case kAdd: v = l + r; break;
      case kSub: v = l - r; break;
      case kMul: v = l * r; break;
      case kDiv: v = l / r; break;
      case kCopyLhs: v = l; break;
      default: v = r; break;
    }
    out[i] = st_acc<T, A>(v);
  }
}

// one sub-warp (sub lanes) per (edge, out-feature); lanes stride over the reduce dim
template <typename T, typename I>
__global__ void __launch_bounds__(256) sddmm_dot_kernel(const I* __restrict__ row, const I* __restrict__ col, int64_t E,
                                                        const T* __restrict__ lhs, const T* __restrict__ rhs,
                                                        T* __restrict__ out, int lt, int rt, int64_t F, int64_t rsz,
                                                        Bcast bc, int sub) {
  using A = typename AccT<T>::type;
  const int groups_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int64_t total = E * F;
  for (int64_t i = (int64_t)blockIdx.x * groups_per_block + threadIdx.x / sub; i < total;
       i += (int64_t)gridDim.x * groups_per_block) {
    const int64_t e = i / F, f = i - e * F;
    const T* lp = lhs + (sel<I>(lt, row, col, e) * bc.lmod + (f / bc.ldiv) % bc.lmod) * rsz;
    const T* rp = rhs + (sel<I>(rt, row, col, e) * bc.rmod + (f / bc.rdiv) % bc.rmod) * rsz;
    A acc = 0;
    for (int64_t k = lane; k < rsz; k += sub) acc += ld_acc<T, A>(lp + k) * ld_acc<T, A>(rp + k);
    for (int o = sub >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[i] = st_acc<T, A>(acc);
  }
}

// Vectorised dot: `sub` lanes per (edge, head), 16-byte loads, four independent items in flight per lane group so
// the two dependent round trips (edge endpoints -> feature rows) of different items overlap.
template <typename T, typename I>
__global__ void __launch_bounds__(256) sddmm_dot_vec_kernel(const I* __restrict__ row, const I* __restrict__ col,
                                                            int64_t E, const T* __restrict__ lhs,
                                                            const T* __restrict__ rhs, T* __restrict__ out, int lt,
                                                            int rt, int64_t F, int64_t rsz, Bcast bc, int sub) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int U = 4;
  const int gpb = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int64_t total = E * F;
  const int64_t ngroups = (int64_t)gridDim.x * gpb;
  for (int64_t i0 = (int64_t)blockIdx.x * gpb + threadIdx.x / sub; i0 < total; i0 += ngroups * U) {
    const T* lp[U];
    const T* rp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * ngroups;
      lp[u] = nullptr;
      if (i < total) {
        int64_t e = i, f = 0;
        if (F != 1) { e = i / F; f = i - e * F; }
        lp[u] = lhs + (sel<I>(lt, row, col, e) * bc.lmod + (f / bc.ldiv) % bc.lmod) * rsz;
        rp[u] = rhs + (sel<I>(rt, row, col, e) * bc.rmod + (f / bc.rdiv) % bc.rmod) * rsz;
      }
    }
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.f;
    for (int64_t k = (int64_t)lane * VEC; k < rsz; k += (int64_t)sub * VEC) {
      uint4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lp[u]) {
          a[u] = __ldg(reinterpret_cast<const uint4*>(lp[u] + k));
          b[u] = __ldg(reinterpret_cast<const uint4*>(rp[u] + k));
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lp[u]) {
          const T* av = reinterpret_cast<const T*>(&a[u]);
          const T* bv = reinterpret_cast<const T*>(&b[u]);
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[u] += to_f<T>(av[q]) * to_f<T>(bv[q]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      for (int o = sub >> 1; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
      if (lane == 0 && lp[u]) out[i0 + u * ngroups] = from_f<T>(acc[u]);
    }
  }
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cuda_gsddmm(int op, int dtype, int idtype, const void* row, const void* col, int64_t E,
                                 const void* lhs, const void* rhs, void* out, int lt, int rt, int64_t F, int64_t rsz,
                                 int64_t ldiv, int64_t lmod, int64_t rdiv, int64_t rmod, void* stream) {
  DGLB_TRY
  Bcast bc{ldiv, lmod, rdiv, rmod};
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_FLOAT_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, {
      constexpr int kVec = 16 / (int)sizeof(T);
      if (op == kDot && sizeof(T) <= 4 && rsz % kVec == 0 && ((uintptr_t)lhs % 16 == 0) && ((uintptr_t)rhs % 16 == 0)) {
        int sub = 1;
        while (sub * kVec < rsz && sub < 32) sub <<= 1;
        const int gpb = 256 / sub;
        int64_t blocks = (E * F + (int64_t)gpb * 4 - 1) / ((int64_t)gpb * 4);
        if (blocks > 148 * 32) blocks = 148 * 32;
        sddmm_dot_vec_kernel<T, I><<<(unsigned)blocks, 256, 0, st>>>((const I*)row, (const I*)col, E, (const T*)lhs,
                                                                     (const T*)rhs, (T*)out, lt, rt, F, rsz, bc, sub);
      } else if (op == kDot) {
        int sub = 1;
        while (sub < rsz && sub < 32) sub <<= 1;
        const int gpb = 256 / sub;
        int64_t blocks = (E * F + gpb - 1) / gpb;
        if (blocks > 148 * 64) blocks = 148 * 64;
        sddmm_dot_kernel<T, I><<<(unsigned)blocks, 256, 0, st>>>((const I*)row, (const I*)col, E, (const T*)lhs,
                                                                 (const T*)rhs, (T*)out, lt, rt, F, rsz, bc, sub);
      } else {
        int64_t blocks = (E * F + 255) / 256;
        if (blocks > 148 * 64) blocks = 148 * 64;
        sddmm_elem_kernel<T, I><<<(unsigned)blocks, 256, 0, st>>>(op, (const I*)row, (const I*)col, E, (const T*)lhs,
                                                                  (const T*)rhs, (T*)out, lt, rt, F, bc);
      }
    });
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
