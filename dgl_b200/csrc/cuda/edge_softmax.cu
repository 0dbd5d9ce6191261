// Edge softmax over the in-edges of every destination node, forward and backward, 16/32-bit floats.
//
//   out[e, h] = exp(score[e, h] - max_{e' -> v} score[e', h]) / sum_{e' -> v} exp(...)        v = dst(e)
//   dscore   = out * (dout - sum_{e' -> v} out * dout)
//
// One kernel per direction for ordinary rows plus one for hub rows:
//  * lanes are laid out as (edge slot, head): the H heads of an edge are one coalesced sector, not H strided
//    passes (the reference gathers per head through g-SDDMM/g-SpMM launches: python/dgl/ops/edge_softmax.py,
//    src/array/cuda/spmm.cuh max / sum reducers + sddmm.cuh);
//  * rows with more than kLongRow in-edges are not processed by their warp: the warp appends the row to a device
//    work list and a second kernel gives each of them a whole 512-thread block (RMAT hubs have 1e5+ in-edges and
//    would otherwise serialise on one warp).
#include "common.cuh"

namespace dglb {

constexpr int kLongRow = 2048;
constexpr int kBlockLong = 512;

// reduce over the lanes that share (lane % hp): xor offsets hp, 2hp, ... 16
__device__ __forceinline__ float group_max(float v, int hp) {
  for (int o = 16; o >= hp; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group_sum(float v, int hp) {
  for (int o = 16; o >= hp; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide variant: combine the per-warp results of the lanes with equal (lane % hp) through shared memory
template <bool IS_MAX>
__device__ __forceinline__ float block_combine(float v, int hp, float* smem /* [32 warps][32] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = IS_MAX ? group_max(v, hp) : group_sum(v, hp);
  __syncthreads();  // smem reuse across calls
  smem[w * 32 + lane] = v;
  __syncthreads();
  float r = smem[lane];
  for (int k = 1; k < nw; ++k) r = IS_MAX ? fmaxf(r, smem[k * 32 + lane]) : r + smem[k * 32 + lane];
  return r;
}

template <typename T, typename I, bool BLOCK>
__device__ __forceinline__ void softmax_row_fwd(const I* __restrict__ eids, int64_t s, int64_t t,
                                                const T* __restrict__ score, T* __restrict__ out, int64_t H, int hp,
                                                float* smem) {
  const int nthr = BLOCK ? blockDim.x : 32;
  const int tid = BLOCK ? threadIdx.x : (threadIdx.x & 31);
  const int hl = tid % hp;          // head inside the chunk
  const int slot = tid / hp;        // edge slot
  const int epi = nthr / hp;        // edges per iteration
  for (int64_t hc = 0; hc < H; hc += hp) {
    const int64_t h = hc + hl;
    const bool hv = h < H;
    float mx = -INFINITY;
    for (int64_t p = s + slot; p < t; p += epi) {
      const int64_t e = eids ? (int64_t)eids[p] : p;
      if (hv) mx = fmaxf(mx, to_f<T>(score[e * H + h]));
    }
    mx = BLOCK ? block_combine<true>(mx, hp, smem) : group_max(mx, hp);
    float sum = 0.f;
    for (int64_t p = s + slot; p < t; p += epi) {
      const int64_t e = eids ? (int64_t)eids[p] : p;
      if (hv) sum += __expf(to_f<T>(score[e * H + h]) - mx);
    }
    sum = BLOCK ? block_combine<false>(sum, hp, smem) : group_sum(sum, hp);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    for (int64_t p = s + slot; p < t; p += epi) {
      const int64_t e = eids ? (int64_t)eids[p] : p;
      if (hv) out[e * H + h] = from_f<T>(__expf(to_f<T>(score[e * H + h]) - mx) * inv);
    }
  }
}

template <typename T, typename I, bool BLOCK>
__device__ __forceinline__ void softmax_row_bwd(const I* __restrict__ eids, int64_t s, int64_t t,
                                                const T* __restrict__ out, const T* __restrict__ gout,
                                                T* __restrict__ gscore, int64_t H, int hp, float* smem) {
  const int nthr = BLOCK ? blockDim.x : 32;
  const int tid = BLOCK ? threadIdx.x : (threadIdx.x & 31);
  const int hl = tid % hp, slot = tid / hp, epi = nthr / hp;
  for (int64_t hc = 0; hc < H; hc += hp) {
    const int64_t h = hc + hl;
    const bool hv = h < H;
    float acc = 0.f;
    for (int64_t p = s + slot; p < t; p += epi) {
      const int64_t e = (eids ? (int64_t)eids[p] : p) * H + h;
      if (hv) acc += to_f<T>(out[e]) * to_f<T>(gout[e]);
    }
    acc = BLOCK ? block_combine<false>(acc, hp, smem) : group_sum(acc, hp);
    for (int64_t p = s + slot; p < t; p += epi) {
      const int64_t e = (eids ? (int64_t)eids[p] : p) * H + h;
      if (hv) {
        const float o_ = to_f<T>(out[e]);
        gscore[e] = from_f<T>(o_ * (to_f<T>(gout[e]) - acc));
      }
    }
  }
}

// work: int32[0] = number of long rows found, int32[1..] = their ids
template <typename T, typename I, bool FWD>
__global__ void __launch_bounds__(256) edge_softmax_rows_kernel(const I* __restrict__ indptr, const I* __restrict__ eids,
                                                                int64_t n, const T* __restrict__ a,
                                                                const T* __restrict__ b, T* __restrict__ o, int64_t H,
                                                                int hp, int32_t* __restrict__ work, int64_t work_cap) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t s = indptr[row], t = indptr[row + 1];
  if (work && t - s > kLongRow) {
    if ((threadIdx.x & 31) == 0) {
      const int k = atomicAdd(work, 1);
      if (k + 1 < work_cap) work[1 + k] = (int32_t)row;
    }
    return;
  }
  if (FWD) softmax_row_fwd<T, I, false>(eids, s, t, a, o, H, hp, nullptr);
  else softmax_row_bwd<T, I, false>(eids, s, t, a, b, o, H, hp, nullptr);
}

template <typename T, typename I, bool FWD>
__global__ void __launch_bounds__(kBlockLong) edge_softmax_long_kernel(const I* __restrict__ indptr,
                                                                       const I* __restrict__ eids,
                                                                       const T* __restrict__ a, const T* __restrict__ b,
                                                                       T* __restrict__ o, int64_t H, int hp,
                                                                       const int32_t* __restrict__ work) {
  __shared__ float smem[(kBlockLong / 32) * 32];
  const int n_long = work[0];
  for (int i = blockIdx.x; i < n_long; i += gridDim.x) {
    const int64_t row = work[1 + i];
    const int64_t s = indptr[row], t = indptr[row + 1];
    if (FWD) softmax_row_fwd<T, I, true>(eids, s, t, a, o, H, hp, smem);
    else softmax_row_bwd<T, I, true>(eids, s, t, a, b, o, H, hp, smem);
  }
}

template <typename T, typename I, bool FWD>
static void launch_softmax(const void* indptr, const void* eids, int64_t n, const void* a, const void* b, void* o,
                           int64_t H, void* work, int64_t work_cap, cudaStream_t st) {
  int hp = 1;
  while (hp < H && hp < 32) hp <<= 1;
  if (work) DGLB_CUDA_CHECK(cudaMemsetAsync(work, 0, 4, st));
  edge_softmax_rows_kernel<T, I, FWD><<<(unsigned)((n + 7) / 8), 256, 0, st>>>(
      (const I*)indptr, (const I*)eids, n, (const T*)a, (const T*)b, (T*)o, H, hp, (int32_t*)work, work_cap);
  if (work)
    edge_softmax_long_kernel<T, I, FWD><<<148 * 2, kBlockLong, 0, st>>>((const I*)indptr, (const I*)eids, (const T*)a,
                                                                       (const T*)b, (T*)o, H, hp, (const int32_t*)work);
}

}  // namespace dglb

using namespace dglb;

// work: int32[work_cap] scratch (work_cap >= E / 2048 + 2), may be null (then hub rows stay on their warp)
DGLB_API int dglb200_cuda_edge_softmax_fwd2(int dtype, int idtype, const void* indptr, const void* eids, int64_t n,
                                            const void* score, void* out, int64_t H, void* work, int64_t work_cap,
                                            void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    switch (dtype) {
      case kF32: launch_softmax<float, I, true>(indptr, eids, n, score, nullptr, out, H, work, work_cap, st); break;
      case kF16: launch_softmax<__half, I, true>(indptr, eids, n, score, nullptr, out, H, work, work_cap, st); break;
      case kBF16: launch_softmax<__nv_bfloat16, I, true>(indptr, eids, n, score, nullptr, out, H, work, work_cap, st); break;
      default: throw std::runtime_error("edge_softmax_fwd2: 16/32-bit floats only");
    }
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_edge_softmax_bwd2(int dtype, int idtype, const void* indptr, const void* eids, int64_t n,
                                            const void* out, const void* gout, void* gscore, int64_t H, void* work,
                                            int64_t work_cap, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    switch (dtype) {
      case kF32: launch_softmax<float, I, false>(indptr, eids, n, out, gout, gscore, H, work, work_cap, st); break;
      case kF16: launch_softmax<__half, I, false>(indptr, eids, n, out, gout, gscore, H, work, work_cap, st); break;
      case kBF16: launch_softmax<__nv_bfloat16, I, false>(indptr, eids, n, out, gout, gscore, H, work, work_cap, st); break;
      default: throw std::runtime_error("edge_softmax_bwd2: 16/32-bit floats only");
    }
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
