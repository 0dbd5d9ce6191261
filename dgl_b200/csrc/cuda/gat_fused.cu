// Fused GAT attention + aggregation (FlashAttention-style online softmax over the in-edges of a destination).
//
//   forward  : out[v,h,:] = sum_{u in N(v)} softmax_u( lrelu(el[u,h] + er[v,h]) ) * ft[u,h,:]
//              one pass over the CSC row: running max / running sum / rescaled accumulator in registers; the [E, H]
//              score, exp and attention tensors of the unfused formulation (g-SDDMM u_add_v -> LeakyReLU -> edge-softmax
//              -> g-SpMM u_mul_e_sum: 4 kernels, 3 [E, H] round trips through HBM) never exist. Saves per (v, h) the
//              max m and the normaliser l, from which the backward recomputes every attention weight.
//   backward : source-parallel over the CSR rows: a = exp(lrelu(el[u]+er[v]) - m[v]) / l[v] is recomputed per edge,
//              d_ft[u] and d_el[u] accumulate in registers (no atomics), d_er[v] is a red.add per (edge, head);
//              softmax backward uses  sum_u a_uv <ft_u, dOut_v> = <out_v, dOut_v>  (delta), so no second pass is needed.
//
// One sub-warp per row: every lane owns one 16-byte piece of the [H x D] row (lane -> head h = lane*VEC / D). Requires
// D % VEC == 0 and either H == 1 or D/VEC a power of two (head reductions are xor-shuffle butterflies).
// Reference formulation being replaced: python/dgl/nn/pytorch/conv/gatconv.py:300-340 + backend/pytorch/sparse.py:685-760.
#include "common.cuh"

namespace dglb {

constexpr float kNegBig = -3.0e38f;

template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
};

template <typename T>
__device__ __forceinline__ void unpack16(const uint4& q, float* f) {
  const T* p = reinterpret_cast<const T*>(&q);
#pragma unroll
  for (int i = 0; i < Vec16<T>::N; ++i) f[i] = to_f<T>(p[i]);
}
template <typename T>
__device__ __forceinline__ uint4 pack16(const float* f) {
  uint4 q;
  T* p = reinterpret_cast<T*>(&q);
#pragma unroll
  for (int i = 0; i < Vec16<T>::N; ++i) p[i] = from_f<T>(f[i]);
  return q;
}

template <typename T, typename I>
__global__ void __launch_bounds__(256) gat_fwd_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                      int64_t n_dst, const T* __restrict__ el, const T* __restrict__ er,
                                                      const T* __restrict__ ft, int H, int D, float slope,
                                                      T* __restrict__ out, float* __restrict__ mbuf,
                                                      float* __restrict__ lbuf, int sub) {
  constexpr int VEC = Vec16<T>::N;
  const int rpb = blockDim.x / sub;
  // rows wider than 32 vectors are covered by column tiles (blockIdx.y); every lane carries the softmax state of its own
  // head, so tiles are independent. `lane` is the vector index inside the row from here on.
  const int lane = blockIdx.y * sub + threadIdx.x % sub;
  const int nvec = H * D / VEC;
  const bool active = lane < nvec;
  const int h = active ? (lane * VEC) / D : 0;
  const int64_t row_elems = (int64_t)H * D;
  for (int64_t row = (int64_t)blockIdx.x * rpb + threadIdx.x / sub; row < n_dst; row += (int64_t)gridDim.x * rpb) {
    const int64_t s0 = indptr[row], s1 = indptr[row + 1];
    const float erv = active ? to_f<T>(er[row * H + h]) : 0.f;
    float m = kNegBig, l = 0.f, acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    for (int64_t p = s0; p < s1; p += 4) {
      float sc[4];
      uint4 f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sc[k] = kNegBig;
        if (p + k < s1 && active) {
          const int64_t u = indices[p + k];
          const float x = to_f<T>(__ldg(el + u * H + h)) + erv;
          sc[k] = x > 0.f ? x : x * slope;
          f[k] = __ldg(reinterpret_cast<const uint4*>(ft + u * row_elems) + lane);
        }
      }
      const float m_new = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), m);
      const float scale = __expf(m - m_new);
      l *= scale;
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] *= scale;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (p + k < s1 && active) {
          const float pk = __expf(sc[k] - m_new);
          l += pk;
          float v[VEC];
          unpack16<T>(f[k], v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(pk, v[i], acc[i]);
        }
      }
      m = m_new;
    }
    if (active) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] *= inv;
      reinterpret_cast<uint4*>(out + row * row_elems)[lane] = pack16<T>(acc);
      if ((lane * VEC) % D == 0) {
        mbuf[row * H + h] = l > 0.f ? m : 0.f;
        lbuf[row * H + h] = l > 0.f ? l : 1.f;
      }
    }
  }
}

// aux[v, h] = (er, m, 1/l, delta) as one float4. CSR rows = source vertices, indices = destination ids.
template <typename T, typename I>
__global__ void __launch_bounds__(256) gat_bwd_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                      int64_t n_src, const T* __restrict__ el, const T* __restrict__ ft,
                                                      const T* __restrict__ dout, const float4* __restrict__ aux, int H,
                                                      int D, float slope, T* __restrict__ d_ft, float* __restrict__ d_el,
                                                      float* __restrict__ d_er, int sub, int red_width) {
  constexpr int VEC = Vec16<T>::N;
  const int rpb = blockDim.x / sub;
  const int lane = blockIdx.y * sub + threadIdx.x % sub;   // vector index inside the row (column tiles in blockIdx.y)
  const int nvec = H * D / VEC;
  const bool active = lane < nvec;
  const int h = active ? (lane * VEC) / D : 0;
  const bool leader = active && (lane * VEC) % D == 0;
  const int64_t row_elems = (int64_t)H * D;
  // sub-warps of one warp walk rows of different lengths: shuffles are confined to the sub-warp's own lanes
  const unsigned gmask = sub == 32 ? 0xffffffffu : (((1u << sub) - 1u) << ((threadIdx.x & 31) & ~(sub - 1)));
  for (int64_t row = (int64_t)blockIdx.x * rpb + threadIdx.x / sub; row < n_src; row += (int64_t)gridDim.x * rpb) {
    const int64_t s0 = indptr[row], s1 = indptr[row + 1];
    float fu[VEC], acc[VEC], del = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) { fu[i] = 0.f; acc[i] = 0.f; }
    float elu = 0.f;
    if (active) {
      unpack16<T>(__ldg(reinterpret_cast<const uint4*>(ft + row * row_elems) + lane), fu);
      elu = to_f<T>(el[row * H + h]);
    }
    for (int64_t p = s0; p < s1; p += 4) {
      uint4 g[4];
      float4 ax[4];
      int64_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = -1;
        if (p + k < s1) {
          v[k] = indices[p + k];
          if (active) {
            g[k] = __ldg(reinterpret_cast<const uint4*>(dout + v[k] * row_elems) + lane);
            ax[k] = __ldg(aux + v[k] * H + h);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (v[k] < 0) continue;  // uniform across the sub-warp
        float gv[VEC], part = 0.f, a = 0.f, x = 0.f;
        if (active) {
          unpack16<T>(g[k], gv);
#pragma unroll
          for (int i = 0; i < VEC; ++i) part = fmaf(fu[i], gv[i], part);
          x = elu + ax[k].x;
          const float s = x > 0.f ? x : x * slope;
          a = __expf(s - ax[k].y) * ax[k].z;
        }
        for (int o = red_width >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o);
        if (active) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(a, gv[i], acc[i]);
          const float de = a * (part - ax[k].w) * (x > 0.f ? 1.f : slope);
          del += de;
          if (leader) atomicAdd(d_er + v[k] * H + h, de);
        }
      }
    }
    if (active) {
      reinterpret_cast<uint4*>(d_ft + row * row_elems)[lane] = pack16<T>(acc);
      if (leader) d_el[row * H + h] = del;
    }
  }
}

}  // namespace dglb

using namespace dglb;

static int gat_sub(int nvec) {
  int sub = 1;
  while (sub < nvec && sub < 32) sub <<= 1;
  return sub;
}

// el: [n_src, H], er: [n_dst, H], ft: [n_src, H, D] (all dtype T); out: [n_dst, H, D]; mbuf / lbuf: [n_dst, H] fp32
DGLB_API int dglb200_cuda_gat_fwd(int dtype, int idtype, const void* indptr, const void* indices, int64_t n_dst,
                                  const void* el, const void* er, const void* ft, int H, int D, float slope, void* out,
                                  void* mbuf, void* lbuf, void* stream) {
  DGLB_TRY
  if (n_dst == 0) return 0;
  auto run = [&](auto tag) {
    using T = decltype(tag);
    constexpr int VEC = 16 / (int)sizeof(T);
    const int nvec = H * D / VEC, lph = D / VEC;
    DGLB_CHECK(D % VEC == 0 && (nvec <= 32 || (H > 1 && lph <= 32 && (lph & (lph - 1)) == 0)),
               "gat_fwd: rows over 512 bytes need heads of 16 B x 2^k (<= 512 B each)");
    const int sub = gat_sub(nvec);
    const int rpb = 256 / sub;
    int64_t blocks = (n_dst + rpb - 1) / rpb;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const dim3 grid((unsigned)blocks, (unsigned)((nvec + sub - 1) / sub));
    DGLB_ID_SWITCH(idtype, I, {
      gat_fwd_kernel<T, I><<<grid, 256, 0, (cudaStream_t)stream>>>(
          (const I*)indptr, (const I*)indices, n_dst, (const T*)el, (const T*)er, (const T*)ft, H, D, slope, (T*)out,
          (float*)mbuf, (float*)lbuf, sub);
    });
  };
  if (dtype == kF32) run(float());
  else if (dtype == kBF16) run(__nv_bfloat16());
  else if (dtype == kF16) run(__half());
  else throw std::runtime_error("gat_fwd: fp32 / bf16 / fp16 only");
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// CSR of the same relation (rows = sources). aux: [n_dst, H, 4] fp32 = (er, m, 1/l, delta). d_er must be zeroed.
DGLB_API int dglb200_cuda_gat_bwd(int dtype, int idtype, const void* indptr, const void* indices, int64_t n_src,
                                  const void* el, const void* ft, const void* dout, const void* aux, int H, int D,
                                  float slope, void* d_ft, void* d_el, void* d_er, void* stream) {
  DGLB_TRY
  if (n_src == 0) return 0;
  auto run = [&](auto tag) {
    using T = decltype(tag);
    constexpr int VEC = 16 / (int)sizeof(T);
    const int lph = D / VEC;
    const int nvec = H * D / VEC;
    DGLB_CHECK(D % VEC == 0 && (nvec <= 32 || (H > 1 && lph <= 32)), "gat_bwd: rows over 512 bytes need heads <= 512 B");
    DGLB_CHECK(H == 1 || (lph & (lph - 1)) == 0, "gat_bwd: lanes per head must be a power of two when H > 1");
    const int sub = gat_sub(nvec);
    const int red_width = H == 1 ? sub : lph;
    const int rpb = 256 / sub;
    int64_t blocks = (n_src + rpb - 1) / rpb;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const dim3 grid((unsigned)blocks, (unsigned)((nvec + sub - 1) / sub));
    DGLB_ID_SWITCH(idtype, I, {
      gat_bwd_kernel<T, I><<<grid, 256, 0, (cudaStream_t)stream>>>(
          (const I*)indptr, (const I*)indices, n_src, (const T*)el, (const T*)ft, (const T*)dout, (const float4*)aux, H, D,
          slope, (T*)d_ft, (float*)d_el, (float*)d_er, sub, red_width);
    });
  };
  if (dtype == kF32) run(float());
  else if (dtype == kBF16) run(__nv_bfloat16());
  else if (dtype == kF16) run(__half());
  else throw std::runtime_error("gat_bwd: fp32 / bf16 / fp16 only");
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
