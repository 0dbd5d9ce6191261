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

// Edge-major vectorised dot (multi-head attention scores / dE of u_mul_e_sum): ONE sub-warp per edge covers all heads.
// Each lane owns 16-byte pieces of the [F heads x rsz] rows of both endpoints (coalesced 16 B x sub loads, the edge's
// endpoints are read once, no per-(edge, head) index arithmetic), partial products are reduced inside the lph lanes of
// a head with xor shuffles and the F results of an edge are written by the head-leader lanes. U edges are in flight
// per sub-warp so the dependent endpoint -> row round trips of different edges overlap.
// Requires: rsz * sizeof(T) = lph * 16 with lph a power of two <= 32, no broadcasting (lmod == rmod == F).
template <typename T, typename I, int U>
__global__ void __launch_bounds__(256) sddmm_dot_edge_kernel(const I* __restrict__ row, const I* __restrict__ col,
                                                             int64_t E, const T* __restrict__ lhs,
                                                             const T* __restrict__ rhs, T* __restrict__ out, int lt,
                                                             int rt, int F, int rsz, int sub, int lph) {
  constexpr int VEC = 16 / sizeof(T);
  const int gpb = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int nvec = F * lph;                 // 16-byte pieces per row
  const int64_t row_elems = (int64_t)F * rsz;
  const int64_t ngroups = (int64_t)gridDim.x * gpb;
  for (int64_t e0 = (int64_t)blockIdx.x * gpb + threadIdx.x / sub; e0 < E; e0 += ngroups * U) {
    const T* lp[U];
    const T* rp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * ngroups;
      lp[u] = nullptr;
      if (e < E) {
        lp[u] = lhs + sel<I>(lt, row, col, e) * row_elems;
        rp[u] = rhs + sel<I>(rt, row, col, e) * row_elems;
      }
    }
    for (int j = lane; j < nvec; j += sub) {  // sub >= lph, so a head never straddles two iterations
      float acc[U];
      uint4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lp[u]) {
          a[u] = __ldg(reinterpret_cast<const uint4*>(lp[u]) + j);
          b[u] = __ldg(reinterpret_cast<const uint4*>(rp[u]) + j);
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[u] = 0.f;
        if (lp[u]) {
          const T* av = reinterpret_cast<const T*>(&a[u]);
          const T* bv = reinterpret_cast<const T*>(&b[u]);
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[u] = fmaf(to_f<T>(av[q]), to_f<T>(bv[q]), acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        for (int o = lph >> 1; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
        if (lp[u] && (j & (lph - 1)) == 0) out[(e0 + u * ngroups) * F + j / lph] = from_f<T>(acc[u]);
      }
    }
  }
}

// Single-head dot with an arbitrary reduce length (rows not 16-byte sized): sub-warp per edge, scalar coalesced loads,
// U edges in flight.
template <typename T, typename I, int U>
__global__ void __launch_bounds__(256) sddmm_dot1_kernel(const I* __restrict__ row, const I* __restrict__ col, int64_t E,
                                                         const T* __restrict__ lhs, const T* __restrict__ rhs,
                                                         T* __restrict__ out, int lt, int rt, int rsz, int sub) {
  const int gpb = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int64_t ngroups = (int64_t)gridDim.x * gpb;
  for (int64_t e0 = (int64_t)blockIdx.x * gpb + threadIdx.x / sub; e0 < E; e0 += ngroups * U) {
    const T* lp[U];
    const T* rp[U];
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * ngroups;
      lp[u] = nullptr;
      acc[u] = 0.f;
      if (e < E) {
        lp[u] = lhs + sel<I>(lt, row, col, e) * rsz;
        rp[u] = rhs + sel<I>(rt, row, col, e) * rsz;
      }
    }
    for (int k = lane; k < rsz; k += sub) {
      float a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lp[u]) {
          a[u] = to_f<T>(__ldg(lp[u] + k));
          b[u] = to_f<T>(__ldg(rp[u] + k));
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lp[u]) acc[u] = fmaf(a[u], b[u], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      for (int o = sub >> 1; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
      if (lane == 0 && lp[u]) out[e0 + u * ngroups] = from_f<T>(acc[u]);
    }
  }
}

// Single-head dot, rows a multiple of 16 bytes (any length): 16-byte loads, sub lanes stride over the row.
template <typename T, typename I, int U>
__global__ void __launch_bounds__(256) sddmm_dot1_vec_kernel(const I* __restrict__ row, const I* __restrict__ col, int64_t E,
                                                             const T* __restrict__ lhs, const T* __restrict__ rhs,
                                                             T* __restrict__ out, int lt, int rt, int rsz, int sub) {
  constexpr int VEC = 16 / sizeof(T);
  const int gpb = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int nvec = rsz / VEC;
  const int64_t ngroups = (int64_t)gridDim.x * gpb;
  for (int64_t e0 = (int64_t)blockIdx.x * gpb + threadIdx.x / sub; e0 < E; e0 += ngroups * U) {
    const uint4* lp[U];
    const uint4* rp[U];
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * ngroups;
      lp[u] = nullptr;
      acc[u] = 0.f;
      if (e < E) {
        lp[u] = reinterpret_cast<const uint4*>(lhs + sel<I>(lt, row, col, e) * rsz);
        rp[u] = reinterpret_cast<const uint4*>(rhs + sel<I>(rt, row, col, e) * rsz);
      }
    }
    for (int j = lane; j < nvec; j += sub) {
      uint4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lp[u]) {
          a[u] = __ldg(lp[u] + j);
          b[u] = __ldg(rp[u] + j);
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lp[u]) {
          const T* av = reinterpret_cast<const T*>(&a[u]);
          const T* bv = reinterpret_cast<const T*>(&b[u]);
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[u] = fmaf(to_f<T>(av[q]), to_f<T>(bv[q]), acc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      for (int o = sub >> 1; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
      if (lane == 0 && lp[u]) out[e0 + u * ngroups] = from_f<T>(acc[u]);
    }
  }
}


// Dot product per edge with the edge endpoints staged through the warp.
//
// A warp owns chunks of 32 consecutive edges. Each lane loads the endpoints of ONE edge of the chunk (two coalesced
// 128-byte index loads per 32 edges, the runtime target selection runs once per lane) and the sub-groups of the warp then
// walk the chunk U edges at a time, pulling "their" endpoints out of the staging registers with a shuffle. The feature
// rows are read with 16-byte loads (sub lanes per edge, U edges in flight per group) exactly like sddmm_dot_edge_kernel,
// but nothing in the inner loop waits on an index load any more - in that kernel the 2 x U dependent endpoint loads
// (branchy runtime `sel`) serialised in front of the row loads and it ran at 36 % occupancy / 34 % SM throughput; ncu
// shows the reference's one-edge-per-warp kernel (src/array/cuda/sddmm.cuh:146-196) at 84 %.
//  * SINGLE (one output per edge): the row may be any multiple of 16 bytes; results are transposed back into lane order
//    and leave as ONE coalesced store per 32 edges (or a scatter through `perm`).
//  * multi-head: lph lanes per head (power of two), heads written by their leader lanes.
// `perm` (optional) maps the processing order to the output position: the caller passes the edges in destination-sorted
// (CSC) order so that consecutive edges re-read the same destination row out of L1, and scatters to edge-id order.
template <int VB> struct VecOf;
template <> struct VecOf<16> { using type = uint4; };
template <> struct VecOf<8> { using type = uint2; };
template <> struct VecOf<4> { using type = uint32_t; };

template <typename T, typename I, int U, bool SINGLE, int VB>
__global__ void __launch_bounds__(256) sddmm_dot_staged_kernel(const I* __restrict__ row, const I* __restrict__ col,
                                                               const I* __restrict__ perm, int64_t E,
                                                               const T* __restrict__ lhs, const T* __restrict__ rhs,
                                                               T* __restrict__ out, int lt, int rt, int H, int lph,
                                                               int sub, int nvec) {
  using V = typename VecOf<VB>::type;
  constexpr int VEC = VB / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int gps = 32 / sub;                    // edge groups per warp
  const int g = lane / sub, gl = lane % sub;   // my group, my lane inside it
  const int64_t row_elems = (int64_t)nvec * VEC;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < E; base += nwarps * 32) {
    const int64_t my_e = base + lane;
    const bool mine = my_e < E;
    const int64_t li = mine ? sel<I>(lt, row, col, my_e) : 0;
    const int64_t ri = mine ? sel<I>(rt, row, col, my_e) : 0;
    float res = 0.f;
    for (int k0 = 0; k0 < 32; k0 += gps * U) {
      if (base + k0 >= E) break;               // warp-uniform
      const V* lp[U];
      const V* rp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int slot = (k0 + u * gps + g) & 31;
        lp[u] = reinterpret_cast<const V*>(lhs + __shfl_sync(0xffffffffu, li, slot) * row_elems);
        rp[u] = reinterpret_cast<const V*>(rhs + __shfl_sync(0xffffffffu, ri, slot) * row_elems);
      }
      float acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = 0.f;
      // multi-head: every lane of the warp runs the same number of iterations (the head reduction shuffles inside)
      const int jend = SINGLE ? nvec : (nvec + sub - 1) / sub * sub;
      for (int j = gl; j < jend; j += sub) {
        V a[U], b[U];
        const bool live = SINGLE || j < nvec;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          a[u] = live ? __ldg(lp[u] + j) : V{};
          b[u] = live ? __ldg(rp[u] + j) : V{};
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const T* av = reinterpret_cast<const T*>(&a[u]);
          const T* bv = reinterpret_cast<const T*>(&b[u]);
          float part = 0.f;
#pragma unroll
          for (int q = 0; q < VEC; ++q) part = fmaf(to_f<T>(av[q]), to_f<T>(bv[q]), part);
          if (SINGLE) {
            acc[u] += part;
          } else {
            for (int o = lph >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            const int64_t e = base + k0 + u * gps + g;
            if (live && (j & (lph - 1)) == 0 && e < E) out[(perm ? (int64_t)perm[e] : e) * H + j / lph] = from_f<T>(part);
          }
        }
      }
      if (SINGLE) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          for (int o = sub >> 1; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
          // hand the result of slot s to lane s: lane L listens to group (L % gps), which worked on slot k0 + u*gps + L % gps
          const float v = __shfl_sync(0xffffffffu, acc[u], (lane % gps) * sub);
          if (k0 + u * gps + (lane % gps) == lane) res = v;
        }
      }
    }
    if (SINGLE && mine) out[perm ? (int64_t)perm[my_e] : my_e] = from_f<T>(res);
  }
}

}  // namespace dglb

using namespace dglb;

static int sddmm_dot_unroll() {
  static const int u = [] {
    const char* e = std::getenv("DGLB200_SDDMM_U");
    const int v = e ? std::atoi(e) : 4;
    return (v == 2 || v == 4) ? v : 4;
  }();
  return u;
}

template <typename T, typename I>
static void launch_dot_staged(const void* row, const void* col, const void* perm, int64_t E, const void* lhs, const void* rhs,
                              void* out, int lt, int rt, int64_t H, int64_t rsz, cudaStream_t st) {
  DGLB_CHECK(sizeof(T) <= 4, "sddmm_dot: fp64 goes through the generic kernel");
  const int64_t row_bytes = H * rsz * (int64_t)sizeof(T);
  const uintptr_t mis = (uintptr_t)lhs | (uintptr_t)rhs;
  // widest vector that divides a row (and, for heads, a head) and matches the table alignment
  const int64_t unit = (H > 1 ? rsz : H * rsz) * (int64_t)sizeof(T);
  int vb = 16;
  while (vb > 4 && (unit % vb || mis % vb)) vb >>= 1;
  DGLB_CHECK(unit % vb == 0 && mis % vb == 0 && vb >= (int)sizeof(T), "sddmm_dot: rows must be a multiple of 4 bytes");
  const int64_t lph64 = rsz * (int64_t)sizeof(T) / vb;   // vectors per head
  const int64_t nvec = row_bytes / vb;
  DGLB_CHECK(H == 1 || (vb == 16 && lph64 <= 32 && (lph64 & (lph64 - 1)) == 0), "sddmm_dot: per-head size must be 16 B x 2^k");
  int sub = 1;
  while (sub < nvec && sub < 32) sub <<= 1;
  if (H > 1 && sub < lph64) sub = (int)lph64;
  const int64_t chunks = (E + 31) / 32;
  int64_t blocks = (chunks + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  auto go = [&](auto kern) {
    kern<<<(unsigned)blocks, 256, 0, st>>>((const I*)row, (const I*)col, (const I*)perm, E, (const T*)lhs, (const T*)rhs,
                                          (T*)out, lt, rt, (int)H, (int)lph64, sub, (int)nvec);
  };
  const bool wide = sddmm_dot_unroll() >= 4;
  if (H > 1) {
    if (wide) go(sddmm_dot_staged_kernel<T, I, 4, false, 16>); else go(sddmm_dot_staged_kernel<T, I, 2, false, 16>);
  } else if (vb == 16) {
    if (wide) go(sddmm_dot_staged_kernel<T, I, 4, true, 16>); else go(sddmm_dot_staged_kernel<T, I, 2, true, 16>);
  } else if (vb == 8) {
    if (wide) go(sddmm_dot_staged_kernel<T, I, 4, true, 8>); else go(sddmm_dot_staged_kernel<T, I, 2, true, 8>);
  } else {
    if (wide) go(sddmm_dot_staged_kernel<T, I, 4, true, 4>); else go(sddmm_dot_staged_kernel<T, I, 2, true, 4>);
  }
}

// out[perm[e]] (or out[e]) = <lhs[sel_l(e)], rhs[sel_r(e)]> per head. H heads of rsz elements, rows of H * rsz elements
// H == 1: any row that is a multiple of 4 bytes (16 / 8 / 4-byte loads are picked from the row size and the table
// alignment); H > 1: rsz * sizeof(T) / 16 a power of two <= 32.
DGLB_API int dglb200_cuda_sddmm_dot(int dtype, int idtype, const void* row, const void* col, const void* perm, int64_t E,
                                    const void* lhs, const void* rhs, void* out, int lt, int rt, int64_t H, int64_t rsz,
                                    void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  if (E == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kF32) launch_dot_staged<float, I>(row, col, perm, E, lhs, rhs, out, lt, rt, H, rsz, st);
    else if (dtype == kBF16) launch_dot_staged<__nv_bfloat16, I>(row, col, perm, E, lhs, rhs, out, lt, rt, H, rsz, st);
    else if (dtype == kF16) launch_dot_staged<__half, I>(row, col, perm, E, lhs, rhs, out, lt, rt, H, rsz, st);
    else throw std::runtime_error("sddmm_dot: unsupported dtype");
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_gsddmm(int op, int dtype, int idtype, const void* row, const void* col, int64_t E,
                                 const void* lhs, const void* rhs, void* out, int lt, int rt, int64_t F, int64_t rsz,
                                 int64_t ldiv, int64_t lmod, int64_t rdiv, int64_t rmod, void* stream) {
  DGLB_TRY
  Bcast bc{ldiv, lmod, rdiv, rmod};
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_FLOAT_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, {
      constexpr int kVec = 16 / (int)sizeof(T);
      const int64_t lph64 = rsz / kVec;
      const bool plain = lmod == F && rmod == F && ldiv == 1 && rdiv == 1;  // no broadcasting between the operands
      if (op == kDot && sizeof(T) <= 4 && plain && rsz % kVec == 0 && lph64 >= 1 && lph64 <= 32 && (lph64 & (lph64 - 1)) == 0 &&
          F * lph64 <= 1024 && ((uintptr_t)lhs % 16 == 0) && ((uintptr_t)rhs % 16 == 0)) {
        const int lph = (int)lph64;
        int sub = lph;
        while (sub < (int)(F * lph) && sub < 32) sub <<= 1;
        const int gpb = 256 / sub;
        int64_t blocks = (E + (int64_t)gpb * 4 - 1) / ((int64_t)gpb * 4);
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (blocks < 1) blocks = 1;
        sddmm_dot_edge_kernel<T, I, 4><<<(unsigned)blocks, 256, 0, st>>>((const I*)row, (const I*)col, E, (const T*)lhs,
                                                                         (const T*)rhs, (T*)out, lt, rt, (int)F, (int)rsz, sub, lph);
      } else if (op == kDot && sizeof(T) <= 4 && F == 1 && lmod == 1 && rmod == 1 && rsz % kVec == 0 && rsz < (1 << 30) &&
                 ((uintptr_t)lhs % 16 == 0) && ((uintptr_t)rhs % 16 == 0)) {
        int sub = 1;
        while (sub * kVec < rsz && sub < 32) sub <<= 1;
        const int gpb = 256 / sub;
        int64_t blocks = (E + (int64_t)gpb * 4 - 1) / ((int64_t)gpb * 4);
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (blocks < 1) blocks = 1;
        sddmm_dot1_vec_kernel<T, I, 4><<<(unsigned)blocks, 256, 0, st>>>((const I*)row, (const I*)col, E, (const T*)lhs,
                                                                         (const T*)rhs, (T*)out, lt, rt, (int)rsz, sub);
      } else if (op == kDot && sizeof(T) <= 4 && F == 1 && lmod == 1 && rmod == 1 && rsz < (1 << 30)) {
        int sub = 1;
        while (sub < rsz && sub < 32) sub <<= 1;
        const int gpb = 256 / sub;
        int64_t blocks = (E + (int64_t)gpb * 4 - 1) / ((int64_t)gpb * 4);
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (blocks < 1) blocks = 1;
        sddmm_dot1_kernel<T, I, 4><<<(unsigned)blocks, 256, 0, st>>>((const I*)row, (const I*)col, E, (const T*)lhs,
                                                                     (const T*)rhs, (T*)out, lt, rt, (int)rsz, sub);
      } else if (op == kDot && sizeof(T) <= 4 && rsz % kVec == 0 && ((uintptr_t)lhs % 16 == 0) && ((uintptr_t)rhs % 16 == 0)) {
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
