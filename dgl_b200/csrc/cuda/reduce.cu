// Segment reduce + fused edge-softmax (forward and backward) for sm_100a.
// The reference has no GPU edge-softmax kernel: it composes 5 launches forward / 3 backward
// (/root/reference/python/dgl/backend/pytorch/sparse.py:709-744). Here forward = one kernel
// (row max, exp-sum, normalise with the row kept in registers/L1), backward = one kernel.
#include "common.cuh"

namespace dglb {

template <typename T, typename I>
__global__ void __launch_bounds__(256) segment_reduce_kernel(int red, const I* __restrict__ offsets, int64_t nseg,
                                                             const T* __restrict__ x, T* __restrict__ out,
                                                             I* __restrict__ arg, int64_t F, int sub) {
  using A = typename AccT<T>::type;
  const int rows_per_block = blockDim.x / sub;
  const int64_t s = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub;
  if (s >= nseg) return;
  const int lane = threadIdx.x % sub;
  const int64_t lo = offsets[s], hi = offsets[s + 1];
  for (int64_t f = lane; f < F; f += sub) {
    A acc = red == kSum ? A(0) : (red == kMax ? -INFINITY : INFINITY);
    I ar = -1;
    for (int64_t r = lo; r < hi; ++r) {
      const A v = ld_acc<T, A>(x + r * F + f);
      if (red == kSum) acc += v;
      else if (red == kMax ? v > acc : v < acc) { acc = v; ar = (I)r; }
    }
    if (red != kSum && lo == hi) acc = A(0);
    out[s * F + f] = st_acc<T, A>(acc);
    if (arg && red != kSum) arg[s * F + f] = ar;
  }
}

// warp per destination row; lanes stride over the row's in-edges; heads looped.
template <typename T, typename I>
__global__ void __launch_bounds__(256) edge_softmax_fwd_kernel(const I* __restrict__ indptr,
                                                               const I* __restrict__ eids, int64_t n,
                                                               const T* __restrict__ score, T* __restrict__ out,
                                                               int64_t H) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t s = indptr[row], t = indptr[row + 1];
  for (int64_t h = 0; h < H; ++h) {
    float mx = -INFINITY;
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = eids ? (int64_t)eids[p] : p;
      mx = fmaxf(mx, to_f<T>(score[e * H + h]));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = eids ? (int64_t)eids[p] : p;
      sum += __expf(to_f<T>(score[e * H + h]) - mx);
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = eids ? (int64_t)eids[p] : p;
      out[e * H + h] = from_f<T>(__expf(to_f<T>(score[e * H + h]) - mx) * inv);
    }
  }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256) edge_softmax_bwd_kernel(const I* __restrict__ indptr,
                                                               const I* __restrict__ eids, int64_t n,
                                                               const T* __restrict__ out, const T* __restrict__ gout,
                                                               T* __restrict__ gscore, int64_t H) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t s = indptr[row], t = indptr[row + 1];
  for (int64_t h = 0; h < H; ++h) {
    float acc = 0.f;
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = (eids ? (int64_t)eids[p] : p) * H + h;
      acc += to_f<T>(out[e]) * to_f<T>(gout[e]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = (eids ? (int64_t)eids[p] : p) * H + h;
      const float o_ = to_f<T>(out[e]);
      gscore[e] = from_f<T>(o_ * (to_f<T>(gout[e]) - acc));
    }
  }
}

// double-precision variants reuse the float math above only for 16/32-bit; fp64 goes through a
// dedicated instantiation with double accumulators
template <typename I>
__global__ void __launch_bounds__(256) edge_softmax_fwd_f64(const I* __restrict__ indptr, const I* __restrict__ eids,
                                                            int64_t n, const double* __restrict__ score,
                                                            double* __restrict__ out, int64_t H) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t s = indptr[row], t = indptr[row + 1];
  for (int64_t h = 0; h < H; ++h) {
    double mx = -INFINITY;
    for (int64_t p = s + lane; p < t; p += 32) mx = fmax(mx, score[(eids ? (int64_t)eids[p] : p) * H + h]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double sum = 0;
    for (int64_t p = s + lane; p < t; p += 32) sum += exp(score[(eids ? (int64_t)eids[p] : p) * H + h] - mx);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double inv = sum > 0 ? 1.0 / sum : 0.0;
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = (eids ? (int64_t)eids[p] : p) * H + h;
      out[e] = exp(score[e] - mx) * inv;
    }
  }
}

template <typename I>
__global__ void __launch_bounds__(256) edge_softmax_bwd_f64(const I* __restrict__ indptr, const I* __restrict__ eids,
                                                            int64_t n, const double* __restrict__ out,
                                                            const double* __restrict__ gout,
                                                            double* __restrict__ gscore, int64_t H) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t s = indptr[row], t = indptr[row + 1];
  for (int64_t h = 0; h < H; ++h) {
    double acc = 0;
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = (eids ? (int64_t)eids[p] : p) * H + h;
      acc += out[e] * gout[e];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    for (int64_t p = s + lane; p < t; p += 32) {
      const int64_t e = (eids ? (int64_t)eids[p] : p) * H + h;
      gscore[e] = out[e] * (gout[e] - acc);
    }
  }
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cuda_segment_reduce(int red, int dtype, int idtype, const void* offsets, int64_t nseg,
                                         const void* x, void* out, void* arg, int64_t F, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  int sub = 1;
  while (sub < F && sub < 32) sub <<= 1;
  const int rpb = 256 / sub;
  DGLB_FLOAT_SWITCH(dtype, T, {
    DGLB_ID_SWITCH(idtype, I, {
      segment_reduce_kernel<T, I><<<(unsigned)((nseg + rpb - 1) / rpb), 256, 0, st>>>(
          red, (const I*)offsets, nseg, (const T*)x, (T*)out, (I*)arg, F, sub);
    });
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_edge_softmax_fwd(int dtype, int idtype, const void* indptr, const void* eids, int64_t n,
                                           const void* score, void* out, int64_t H, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 7) / 8);
  DGLB_ID_SWITCH(idtype, I, {
    switch (dtype) {
      case kF32: edge_softmax_fwd_kernel<float, I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const float*)score, (float*)out, H); break;
      case kF16: edge_softmax_fwd_kernel<__half, I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const __half*)score, (__half*)out, H); break;
      case kBF16: edge_softmax_fwd_kernel<__nv_bfloat16, I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const __nv_bfloat16*)score, (__nv_bfloat16*)out, H); break;
      case kF64: edge_softmax_fwd_f64<I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const double*)score, (double*)out, H); break;
      default: throw std::runtime_error("bad dtype");
    }
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_edge_softmax_bwd(int dtype, int idtype, const void* indptr, const void* eids, int64_t n,
                                           const void* out, const void* gout, void* gscore, int64_t H, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 7) / 8);
  DGLB_ID_SWITCH(idtype, I, {
    switch (dtype) {
      case kF32: edge_softmax_bwd_kernel<float, I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const float*)out, (const float*)gout, (float*)gscore, H); break;
      case kF16: edge_softmax_bwd_kernel<__half, I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const __half*)out, (const __half*)gout, (__half*)gscore, H); break;
      case kBF16: edge_softmax_bwd_kernel<__nv_bfloat16, I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const __nv_bfloat16*)out, (const __nv_bfloat16*)gout, (__nv_bfloat16*)gscore, H); break;
      case kF64: edge_softmax_bwd_f64<I><<<grid, 256, 0, st>>>((const I*)indptr, (const I*)eids, n, (const double*)out, (const double*)gout, (double*)gscore, H); break;
      default: throw std::runtime_error("bad dtype");
    }
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
