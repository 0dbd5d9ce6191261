// g-SpMM (copy_u / u_mul_e, sum) with the neighbour rows staged through shared memory by the TMA engine.
//
// Every warp owns a run of destination rows and a ring of S feature-row slots in shared memory. One lane issues a
// row-granular bulk copy (cp.async.bulk.shared::cluster.global, completion on the slot's mbarrier - SASS: UBLKCP) per
// neighbour, S neighbours AHEAD of the row being reduced and across row boundaries, so the number of feature rows in
// flight per SM is bounded by shared memory (warps x S rows), not by registers: the SIMT formulation (spmm_vec_kernel)
// keeps 4 rows x NV vectors per lane in flight. The lanes then reduce the landed row from shared memory in fp32 and the
// finished destination row leaves with 16-byte stores. Rows longer than `max_deg` are skipped (the edge-balanced split
// kernel owns them), exactly like spmm_vec_kernel.
//
// Requirements: row bytes a multiple of 16 and <= kMaxRowBytes, 16-byte aligned feature table.
// Replaces the row gather of src/array/cuda/spmm.cuh:499-543 (one thread block per row, LDG per element).
#pragma once
#include "tc_common.cuh"

namespace dglb {
namespace bulk {

constexpr int kWarps = 8;
constexpr int kMaxRowBytes = 2048;
constexpr int kMaxSlots = 8;

template <typename T>
__device__ __forceinline__ void accum16(float* acc, const uint4& q, float w) {
  if constexpr (sizeof(T) == 4) {
    acc[0] = fmaf(__uint_as_float(q.x), w, acc[0]); acc[1] = fmaf(__uint_as_float(q.y), w, acc[1]);
    acc[2] = fmaf(__uint_as_float(q.z), w, acc[2]); acc[3] = fmaf(__uint_as_float(q.w), w, acc[3]);
  } else {
    const T* p = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(to_f<T>(p[i]), w, acc[i]);
  }
}

template <typename T>
__device__ __forceinline__ uint4 pack16(const float* acc) {
  uint4 q;
  if constexpr (sizeof(T) == 4) {
    q = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
  } else {
    T* p = reinterpret_cast<T*>(&q);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = from_f<T>(acc[i]);
  }
  return q;
}

// NV = 16-byte vectors per lane (row bytes <= NV * 512)
template <typename T, typename I, bool HAS_W, int NV>
__global__ void __launch_bounds__(kWarps * 32) spmm_bulk_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                                const I* __restrict__ eids, int64_t num_rows,
                                                                const T* __restrict__ x, const T* __restrict__ w,
                                                                T* __restrict__ out, int64_t F, int64_t max_deg,
                                                                int rows_per_warp, int S, uint32_t row_bytes,
                                                                uint32_t slot_stride) {
  extern __shared__ __align__(128) uint8_t smem_bulk[];
  constexpr int VE = 16 / (int)sizeof(T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem_bulk + (size_t)warp * S * slot_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bulk + (size_t)kWarps * S * slot_stride) + warp * kMaxSlots;
  if (lane < S) mbar_init(smem_u32(bars + lane), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t row0 = ((int64_t)blockIdx.x * kWarps + warp) * rows_per_warp;
  const int64_t row1 = min(num_rows, row0 + rows_per_warp);
  if (row0 >= row1) return;
  const int nvec = (int)(row_bytes >> 4);

  // ---- prefetch cursor (warp-uniform): next edge to request
  int64_t pr = row0, pp = 0, pe = 0;
  auto settle = [&]() -> bool {
    while (pr < row1) {
      const int64_t s = indptr[pr], e = indptr[pr + 1];
      if (e > s && (max_deg <= 0 || e - s <= max_deg)) {
        pp = s;
        pe = e;
        return true;
      }
      ++pr;
    }
    return false;
  };
  auto request = [&](int slot) {  // all lanes advance the cursor, lane 0 talks to the TMA engine
    if (lane == 0) {
      const uint32_t bar = smem_u32(bars + slot);
      mbar_arrive_expect_tx(bar, row_bytes);
      bulk_g2s(smem_u32(ring + (size_t)slot * slot_stride), x + (int64_t)indices[pp] * F, row_bytes, bar);
    }
    if (++pp == pe) ++pr;
  };
  bool more = settle();
  uint32_t issued = 0;
  while (more && (int)issued < S) {
    request((int)issued);
    ++issued;
    if (pp == pe) more = settle();
  }

  // ---- consume in row order
  uint32_t q = 0;
  for (int64_t r = row0; r < row1; ++r) {
    const int64_t s = indptr[r], e = indptr[r + 1];
    if (max_deg > 0 && e - s > max_deg) continue;  // written by the split kernel's finalize pass
    float acc[NV][VE];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[j][i] = 0.f;
    for (int64_t p = s; p < e; ++p, ++q) {
      const int slot = (int)(q % (uint32_t)S);
      mbar_wait(smem_u32(bars + slot), (q / (uint32_t)S) & 1);
      const float wt = HAS_W ? to_f<T>(w[eids ? (int64_t)eids[p] : p]) : 1.f;
      const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)slot * slot_stride);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = lane + 32 * j;
        if (c < nvec) accum16<T>(acc[j], src[c], wt);
      }
      fence_proxy_async();  // generic reads of the slot are ordered before the async-proxy refill
      __syncwarp();
      if (more) {
        request(slot);
        if (pp == pe) more = settle();
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + r * F);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < nvec) dst[c] = pack16<T>(acc[j]);
    }
  }
}

// true if the bulk kernel was launched for this shape
template <typename T, typename I, bool HAS_W>
bool launch(const I* indptr, const I* indices, const I* eids, int64_t num_rows, const T* x, const T* w, T* out, int64_t F,
            int64_t max_deg, cudaStream_t st) {
  const int64_t row_bytes = F * (int64_t)sizeof(T);
  if (row_bytes % 16 || row_bytes < 64 || row_bytes > kMaxRowBytes || ((uintptr_t)x & 15) || ((uintptr_t)out & 15)) return false;
  static const bool enabled = [] {
    // opt-in: on B200 the SIMT vector kernel wins (0.33-0.95x of it measured over arxiv/reddit/products x F 32..256,
    // profiles/kernels_bulk_vs_simt.md) - one elected lane issuing row copies and an mbarrier wait per neighbour row leave
    // the warp latency-bound, while the SIMT kernel keeps 4 rows x NV vectors per lane in flight with no waits.
    const char* e = std::getenv("DGLB200_SPMM_BULK");
    return e && e[0] == '1';
  }();
  if (!enabled) return false;
  const uint32_t stride = (uint32_t)((row_bytes + 127) / 128 * 128);
  int S = (int)std::min<int64_t>(kMaxSlots, (96 * 1024) / ((int64_t)kWarps * stride));
  if (S < 2) return false;
  const size_t smem = (size_t)kWarps * S * stride + (size_t)kWarps * kMaxSlots * 8;
  const int rows_per_warp = 16;
  const int64_t blocks = (num_rows + (int64_t)kWarps * rows_per_warp - 1) / ((int64_t)kWarps * rows_per_warp);
  const int nv = (int)((row_bytes / 16 + 31) / 32);
  auto go = [&](auto kern) {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, kWarps * 32, smem, st>>>(indptr, indices, eids, num_rows, x, w, out, F, max_deg, rows_per_warp, S,
                                                     (uint32_t)row_bytes, stride);
  };
  if (nv <= 1) go(spmm_bulk_kernel<T, I, HAS_W, 1>);
  else if (nv <= 2) go(spmm_bulk_kernel<T, I, HAS_W, 2>);
  else go(spmm_bulk_kernel<T, I, HAS_W, 4>);
  return true;
}

}  // namespace bulk
}  // namespace dglb
