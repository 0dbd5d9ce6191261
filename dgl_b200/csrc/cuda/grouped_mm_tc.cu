// Grouped (per-relation) GEMMs on 5th-gen tensor cores (sm_100a): segment_mm / gather_mm and their backward.
//
//   forward / dA :  out[i, :] = A[ia(i), :K] @ B[r(i)]^T        rows sorted by relation r, B[r] stored [Nout, K] (K-major)
//   dW           :  dW[r]    += A[seg_r, :K]^T @ dO[seg_r, :Nout]   (fp32, red.add; long segments are split into row chunks)
//
// The reference runs these as a HOST loop of R cuBLAS calls with the segment lengths copied to the CPU first
// (src/array/cuda/gather_mm.cu:201-285, a D2H sync in nn/pytorch/linear.py:205-207) or as a warp-per-row SIMT
// kernel for the unsorted case (gather_mm.cu:86-186). Here the segment table stays on the device: a persistent
// kernel walks a tile list derived from the prefix sums of the per-relation tile counts (binary search, no host
// round trip), so one launch covers all relations, and launches are CUDA-graph capturable.
//
// Forward kernel (warp-specialised, one CTA per SM, same skeleton as sage_tc.cu):
//   * contiguous A: one TMA thread streams the [128 x 64] A box and the [Np x 64] chunk of B[r] (3-D tensor map over
//     (K, Nout, R), OOB zero fill) of every K chunk through a 4-stage mbarrier ring;
//   * gathered A (idx_a): 16 producer warps GATHER the 128 A rows of a tile into the K-major SWIZZLE_128B smem image,
//     16 B per lane, up to 4 K chunks in flight, zero-filling rows past the segment end and columns past K;
//   * one thread issues tcgen05.mma (M128 x N<=256 x K16) into one of two TMEM accumulators;
//   * 8 epilogue warps drain the other accumulator with tcgen05.ld and write bf16/fp16/fp32 rows (masked to the segment).
// dW kernel: both operands are MN-major (the reduce dimension is the row index): producers copy 64-row chunks of A and
//   dO into [64 rows x 128 B] boxes (the same swizzled image), rows outside the segment are zeroed, the MMA uses
//   MN-major descriptors (a_major = b_major = 1), the epilogue adds the [128 x Nout] fp32 block into dW[r] with
//   vector reductions.
#include "tc_common.cuh"

namespace dglb {
namespace gmm {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;
constexpr int kStages = 4;      // dW kernel ring depth
constexpr int kMaxStages = 8;   // forward kernel: as many stages as fit (deeper prefetch for small K / N tiles)
constexpr int kProducerWarps = 16;
constexpr int kEpilogueWarps = 8;
constexpr int kTmaWarp = kProducerWarps + kEpilogueWarps;
constexpr int kMmaWarp = kTmaWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 832
constexpr int kDwRowChunk = 2048;              // rows of one dW work item (32 MMA chunks of 64 rows)

// tile t -> (relation r, first row, live rows). tile_off[r] = number of tiles of relations < r (tile_off[R] = total).
__device__ __forceinline__ void locate_tile(int64_t t, const int64_t* __restrict__ tile_off, int R,
                                            const int64_t* __restrict__ seg_off, int rows_per_tile, int& r,
                                            int64_t& row0, int& rows) {
  int lo = 0, hi = R;  // largest r with tile_off[r] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tile_off[mid] <= t) lo = mid; else hi = mid;
  }
  r = lo;
  const int64_t s = seg_off[r], e = seg_off[r + 1];
  row0 = s + (t - tile_off[r]) * rows_per_tile;
  const int64_t left = e - row0;
  rows = (int)(left < rows_per_tile ? left : rows_per_tile);
}

// out dtype codes: 0 = same 16-bit type as the inputs, 1 = fp32
template <bool IS_BF16>
__global__ void __launch_bounds__(kThreads, 1)
seg_mm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_o, const int tma_a, const int tma_out,
                 const uint16_t* __restrict__ A, int64_t lda,
                 const int64_t* __restrict__ idx_a, const int64_t* __restrict__ seg_off,
                 const int64_t* __restrict__ tile_off, int R, int K, int Nout, int NB, void* __restrict__ out,
                 int out_f32, const int nst) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t n_row_tiles = tile_off[R];
  const int64_t n_tiles = n_row_tiles * NB;
  if ((int64_t)blockIdx.x >= n_tiles) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = (K + kChunkK - 1) / kChunkK;
  const int Np_full = Nout < 256 ? ((Nout + 15) & ~15) : 256;  // MMA N of a full column block
  const uint32_t a_bytes = kTileM * 128;
  const uint32_t b_bytes = (uint32_t)Np_full * 128;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  // output staging for the TMA-store epilogue: Np/64 panels of [128 rows x 128 B], SWIZZLE_128B (1024-byte aligned)
  uint8_t* stage_out = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars) + 256 + 1023) & ~(uintptr_t)1023);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages);
  const uint32_t accf0 = smem_u32(bars + 2 * kMaxStages), acce0 = smem_u32(bars + 2 * kMaxStages + 2);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(full0 + 8 * s, tma_a ? 1 : kProducerWarps + 1);  // contiguous A: the TMA thread brings both operands
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf0 + 8 * a, 1);
      mbar_init(acce0 + 8 * a, kEpilogueWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t acc_cols = Np_full <= 32 ? 32 : (Np_full <= 64 ? 64 : (Np_full <= 128 ? 128 : 256));
  const uint32_t tmem_cols = 2 * acc_cols;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < kProducerWarps) {
    // ===================== producers: GATHERED A rows -> swizzled K-major smem tile (idle when A comes by TMA) =====
    const int tid = threadIdx.x;
    if (tma_a) goto done;
    uint32_t g = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int r, rows;
      int64_t row0;
      locate_tile(t / NB, tile_off, R, seg_off, kTileM, r, row0, rows);
      // each thread owns 2 (row, 16-byte chunk) slots of every K chunk: rows tid>>3 and 64 + (tid>>3), chunk tid&7
      const int c = tid & 7;
      const uint16_t* src[2];
      bool live[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rr = (tid >> 3) + h * 64;
        live[h] = rr < rows;
        const int64_t grow = live[h] ? (idx_a ? idx_a[row0 + rr] : row0 + rr) : 0;
        src[h] = A + grow * lda + c * 8;
      }
      // the loads of up to 4 K chunks (8 x 16 B per thread) are issued before the first smem slot is awaited, so a
      // tile's whole A block is in flight at once
      for (int kc0 = 0; kc0 < KC; kc0 += 4) {
        uint4 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int kc = kc0 + u;
          const bool col_ok = kc < KC && kc * kChunkK + c * 8 < K;  // K % 8 == 0: a chunk is fully inside or outside
#pragma unroll
          for (int h = 0; h < 2; ++h)
            v[u][h] = (live[h] && col_ok) ? __ldg(reinterpret_cast<const uint4*>(src[h] + kc * kChunkK)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (kc0 + u < KC) {
            const int stage = g % nst;
            mbar_wait(empty0 + 8 * stage, ((g / nst) & 1) ^ 1);
            uint8_t* sa = smem + stage * stage_bytes;
#pragma unroll
            for (int h = 0; h < 2; ++h) *reinterpret_cast<uint4*>(sa + sw128_off((tid >> 3) + h * 64, c)) = v[u][h];
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * stage);
            ++g;
          }
        }
      }
    }
  } else if (warp < kTmaWarp) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int ew = warp - kProducerWarps;
    const int quarter = ew & 3;
    const int row = quarter * 32 + lane;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      int r, rows;
      int64_t row0;
      locate_tile(t / NB, tile_off, R, seg_off, kTileM, r, row0, rows);
      const int nb = (int)(t % NB);
      const int ncols = min(256, Nout - nb * 256);  // live output columns of this block (multiple of 8)
      const int acc = it & 1;
      mbar_wait(accf0 + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const bool live = row < rows;
      const int64_t grow = row0 + row;
      // full tiles leave through shared memory + TMA store (coalesced 128-byte rows, no partial-sector stores);
      // partial tiles (segment tails) keep the masked direct stores below
      const bool via_tma = tma_out && rows == kTileM && (ncols & 63) == 0;
      if (via_tma) {
        if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging is free again
        asm volatile("bar.sync 3, %0;" ::"n"(kEpilogueWarps * 32) : "memory");
        for (int n0 = (ew >> 2) * 32; n0 < ncols; n0 += 32 * (kEpilogueWarps / 4)) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + n0), v);
          uint8_t* panel = stage_out + (n0 >> 6) * (kTileM * 128);
          const int c0 = (n0 & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x0 = __uint_as_float(v[j + 2 * q]), x1 = __uint_as_float(v[j + 2 * q + 1]);
              if (IS_BF16) {
                __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                pk[q] = *reinterpret_cast<uint32_t*>(&hh);
              } else {
                __half2 hh = __floats2half2_rn(x0, x1);
                pk[q] = *reinterpret_cast<uint32_t*>(&hh);
              }
            }
            *reinterpret_cast<uint4*>(panel + sw128_off(row, c0 + (j >> 3))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce0 + 8 * acc);  // TMEM accumulator is free as soon as it has been read
        fence_proxy_async();
        asm volatile("bar.sync 3, %0;" ::"n"(kEpilogueWarps * 32) : "memory");
        if (ew == 0 && lane == 0) {
          for (int pn = 0; pn < (ncols >> 6); ++pn)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap_o),
                         "r"(nb * 256 + pn * 64), "r"((int)row0), "r"(smem_u32(stage_out + pn * (kTileM * 128)))
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        continue;
      }
      for (int n0 = (ew >> 2) * 32; n0 < ncols; n0 += 32 * (kEpilogueWarps / 4)) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + n0), v);
        if (!live) continue;
        const int nlive = min(32, ncols - n0);
        if (out_f32) {
          float* dst = reinterpret_cast<float*>(out) + grow * Nout + nb * 256 + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < nlive)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
          uint16_t* dst = reinterpret_cast<uint16_t*>(out) + grow * Nout + nb * 256 + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < nlive) {
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float x0 = __uint_as_float(v[j + 2 * q]), x1 = __uint_as_float(v[j + 2 * q + 1]);
                if (IS_BF16) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                  pk[q] = *reinterpret_cast<uint32_t*>(&h);
                } else {
                  __half2 h = __floats2half2_rn(x0, x1);
                  pk[q] = *reinterpret_cast<uint32_t*>(&h);
                }
              }
              *reinterpret_cast<uint4*>(dst + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce0 + 8 * acc);
    }
    if (tma_out && ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp == kTmaWarp) {
    // ===================== TMA producer for the B[r] chunks =====================
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int r, rows;
        int64_t row0;
        locate_tile(t / NB, tile_off, R, seg_off, kTileM, r, row0, rows);
        const int nb = (int)(t % NB);
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int stage = g % nst;
          mbar_wait(empty0 + 8 * stage, ((g / nst) & 1) ^ 1);
          mbar_arrive_expect_tx(full0 + 8 * stage, tma_a ? a_bytes + b_bytes : b_bytes);
          if (tma_a) tma_load_2d(smem_u32(smem + stage * stage_bytes), &tmap_a, full0 + 8 * stage, kc * kChunkK, (int)row0);
          tma_load_3d(smem_u32(smem + stage * stage_bytes + a_bytes), &tmap_b, full0 + 8 * stage, kc * kChunkK, nb * 256, r);
        }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    const uint32_t fmt = IS_BF16 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(Np_full >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    uint32_t g = 0;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(acce0 + 8 * acc, ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
      for (int kc = 0; kc < KC; ++kc, ++g) {
        const int stage = g % nst;
        mbar_wait(full0 + 8 * stage, (g / nst) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(smem + stage * stage_bytes), b0 = a0 + a_bytes;
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k)
            umma_bf16(tmem_d, make_sw128_kmajor_desc(a0 + k * 32), make_sw128_kmajor_desc(b0 + k * 32), idesc,
                      (kc | k) != 0 ? 1u : 0u);
          umma_commit(empty0 + 8 * stage);
          if (kc == KC - 1) umma_commit(accf0 + 8 * acc);
        }
        __syncwarp();
      }
    }
  }
done:
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// dW[r][k, n] += sum_{i in seg_r} A[i, k] * dO[i, n].  Work item = (relation, chunk of kDwRowChunk rows, 128-wide k
// block, 256-wide n block); item_off[r] = number of row chunks of relations < r.
template <bool IS_BF16>
__global__ void __launch_bounds__(kThreads, 1)
seg_mm_dw_tc_kernel(const __grid_constant__ CUtensorMap tmap_do, const int tma_do,
                    const uint16_t* __restrict__ A, int64_t lda, const int64_t* __restrict__ idx_a,
                    const uint16_t* __restrict__ dO, int64_t ldo, const int64_t* __restrict__ seg_off,
                    const int64_t* __restrict__ item_off, int R, int K, int Nout, int KB, int NB,
                    float* __restrict__ dW) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t n_items = item_off[R] * KB * NB;
  if ((int64_t)blockIdx.x >= n_items) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Np_full = Nout < 256 ? ((Nout + 15) & ~15) : 256;
  const int n_boxes = (Np_full + 63) / 64;             // 64-feature boxes of the dO operand
  const uint32_t box_bytes = kChunkK * 128;            // 64 rows x 128 B
  const uint32_t a_bytes = 2 * box_bytes;              // 128 A features
  const uint32_t b_bytes = (uint32_t)n_boxes * box_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages);
  const uint32_t accf0 = smem_u32(bars + 2 * kStages), acce0 = smem_u32(bars + 2 * kStages + 2);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, kProducerWarps + (tma_do ? 1 : 0));  // + the TMA thread that brings the dO boxes
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf0 + 8 * a, 1);
      mbar_init(acce0 + 8 * a, kEpilogueWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t acc_cols = Np_full <= 32 ? 32 : (Np_full <= 64 ? 64 : (Np_full <= 128 ? 128 : 256));
  const uint32_t tmem_cols = 2 * acc_cols;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int per_chunk = KB * NB;

  if (warp < kProducerWarps) {
    // producers: 64-row chunks of A (128 features) and dO (n_boxes * 64 features) -> MN-major boxes
    const int tid = threadIdx.x;
    const int rr = tid >> 3, c = tid & 7;  // row of the chunk, 16-byte piece inside a 128-byte box row
    uint32_t g = 0;
    for (int64_t t = blockIdx.x; t < n_items; t += gridDim.x) {
      int r, rows;
      int64_t row0;
      locate_tile(t / per_chunk, item_off, R, seg_off, kDwRowChunk, r, row0, rows);
      const int kb = (int)((t % per_chunk) / NB), nb = (int)(t % NB);
      const int n_sub = (rows + kChunkK - 1) / kChunkK;
      for (int sc = 0; sc < n_sub; ++sc, ++g) {
        const int stage = g % kStages;
        mbar_wait(empty0 + 8 * stage, ((g / kStages) & 1) ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        const int lr = sc * kChunkK + rr;
        const bool live = lr < rows;
        const int64_t orow = row0 + lr;                               // dO is indexed by the sorted position
        const int64_t arow = live ? (idx_a ? idx_a[orow] : orow) : 0;  // A optionally through the gather index
        uint4 va[2], vb[4];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int col = kb * 128 + b * 64 + c * 8;
          va[b] = (live && col < K) ? __ldg(reinterpret_cast<const uint4*>(A + arow * lda + col)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int col = nb * 256 + b * 64 + c * 8;
          vb[b] = (!tma_do && b < n_boxes && live && col < Nout) ? __ldg(reinterpret_cast<const uint4*>(dO + orow * ldo + col))
                                                                 : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) *reinterpret_cast<uint4*>(sa + b * box_bytes + sw128_off(rr, c)) = va[b];
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (!tma_do && b < n_boxes) *reinterpret_cast<uint4*>(sb + b * box_bytes + sw128_off(rr, c)) = vb[b];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * stage);
      }
    }
  } else if (warp < kTmaWarp) {
    // epilogue: [128 k-features x Nout] fp32 block -> red.add into dW[r]
    const int ew = warp - kProducerWarps;
    const int quarter = ew & 3;
    const int krow = quarter * 32 + lane;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_items; t += gridDim.x, ++it) {
      int r, rows;
      int64_t row0;
      locate_tile(t / per_chunk, item_off, R, seg_off, kDwRowChunk, r, row0, rows);
      const int kb = (int)((t % per_chunk) / NB), nb = (int)(t % NB);
      const int ncols = min(256, Nout - nb * 256);
      const int acc = it & 1;
      mbar_wait(accf0 + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const int k = kb * 128 + krow;
      for (int n0 = (ew >> 2) * 32; n0 < ncols; n0 += 32 * (kEpilogueWarps / 4)) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + n0), v);
        if (k >= K) continue;
        const int nlive = min(32, ncols - n0);
        float* dst = dW + ((int64_t)r * K + k) * Nout + nb * 256 + n0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j < nlive) {
            if ((Nout & 3) == 0) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (j + q < nlive) atomicAdd(dst + j + q, __uint_as_float(v[j + q]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce0 + 8 * acc);
    }
  } else if (warp == kTmaWarp) {
    // dO boxes by TMA (contiguous rows, unmasked: rows outside the segment meet zeroed A rows, rows / columns outside the
    // tensor are zero-filled); the A operand stays with the producers because its out-of-segment rows must be zero
    if (tma_do && lane == 0) {
      uint32_t g = 0;
      for (int64_t t = blockIdx.x; t < n_items; t += gridDim.x) {
        int r, rows;
        int64_t row0;
        locate_tile(t / per_chunk, item_off, R, seg_off, kDwRowChunk, r, row0, rows);
        const int nb = (int)(t % NB);
        const int n_sub = (rows + kChunkK - 1) / kChunkK;
        for (int sc = 0; sc < n_sub; ++sc, ++g) {
          const int stage = g % kStages;
          mbar_wait(empty0 + 8 * stage, ((g / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(full0 + 8 * stage, b_bytes);
          const uint32_t sb = smem_u32(smem + stage * stage_bytes + a_bytes);
          for (int b = 0; b < n_boxes; ++b)
            tma_load_2d(sb + b * box_bytes, &tmap_do, full0 + 8 * stage, nb * 256 + b * 64, (int)(row0 + (int64_t)sc * kChunkK));
        }
      }
    }
  } else if (warp == kMmaWarp) {
    const uint32_t fmt = IS_BF16 ? 1u : 0u;
    // both operands MN-major (bits 15, 16)
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(Np_full >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    uint32_t g = 0;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_items; t += gridDim.x, ++it) {
      int r, rows;
      int64_t row0;
      locate_tile(t / per_chunk, item_off, R, seg_off, kDwRowChunk, r, row0, rows);
      const int n_sub = (rows + kChunkK - 1) / kChunkK;
      const int acc = it & 1;
      mbar_wait(acce0 + 8 * acc, ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
      for (int sc = 0; sc < n_sub; ++sc, ++g) {
        const int stage = g % kStages;
        mbar_wait(full0 + 8 * stage, (g / kStages) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(smem + stage * stage_bytes), b0 = a0 + a_bytes;
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k)  // 16 reduce rows per MMA = 2 groups of 8 rows = 2048 B
            umma_bf16(tmem_d, make_sw128_mnmajor_desc(a0 + k * 2048, box_bytes), make_sw128_mnmajor_desc(b0 + k * 2048, box_bytes),
                      idesc, (sc | k) != 0 ? 1u : 0u);
          umma_commit(empty0 + 8 * stage);
          if (sc == n_sub - 1) umma_commit(accf0 + 8 * acc);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace gmm
}  // namespace dglb

using namespace dglb;
using namespace dglb::gmm;

// A: [*, lda] 16-bit (bf16: dtype kBF16, fp16: kF16), rows addressed as idx_a[i] (int64, optional) or i, i in [0, n_rows)
// Bt: [R, Nout, K] contiguous (K-major), seg_off: [R+1] int64 row offsets, tile_off: [R+1] int64 prefix of ceil(len/128)
// out: [n_rows, Nout] same 16-bit type, or fp32 when out_f32
DGLB_API int dglb200_cuda_segment_mm_tc(int dtype, const void* A, int64_t n_a_rows, int64_t n_out_rows, int64_t lda, const void* idx_a, const void* Bt,
                                        const void* seg_off, const void* tile_off, int64_t max_tiles, int R, int K,
                                        int Nout, void* out, int out_f32, void* stream) {
  DGLB_TRY
  DGLB_CHECK(dtype == kBF16 || dtype == kF16, "segment_mm_tc: bf16 / fp16 inputs only");
  DGLB_CHECK(K % 8 == 0 && Nout % 8 == 0 && lda % 8 == 0, "segment_mm_tc: K, Nout and the row stride must be multiples of 8");
  DGLB_CHECK(((uintptr_t)A % 16 == 0) && ((uintptr_t)Bt % 16 == 0) && ((uintptr_t)out % 16 == 0), "16-byte alignment required");
  if (max_tiles <= 0 || R <= 0) return 0;
  const int NB = (Nout + 255) / 256;
  const int Np_full = Nout < 256 ? ((Nout + 15) & ~15) : 256;
  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)Nout, (cuuint64_t)R};
  const cuuint64_t gstride[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * (cuuint64_t)Nout};
  const cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)Np_full, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = get_encode()(&tmap, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                             const_cast<void*>(Bt), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
  // contiguous (un-gathered) A: a 2-D tensor map over [n_a_rows, K] with row stride lda; rows of a tile that belong to
  // the next segment are loaded too and masked at the store, rows past the end of A are zero-filled by TMA
  const int tma_a = idx_a == nullptr ? 1 : 0;
  CUtensorMap tmap_a = tmap;
  if (tma_a) {
    const cuuint64_t adim[2] = {(cuuint64_t)K, (cuuint64_t)n_a_rows};
    const cuuint64_t astride[1] = {(cuuint64_t)lda * 2};
    const cuuint32_t abox[2] = {(cuuint32_t)kChunkK, (cuuint32_t)kTileM};
    rc = get_encode()(&tmap_a, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                      const_cast<void*>(A), adim, astride, abox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)rc));
  }
  // TMA-store epilogue for 16-bit outputs whose width is a multiple of 64 columns
  const int tma_out = (!out_f32 && Nout % 64 == 0) ? 1 : 0;
  CUtensorMap tmap_o = tmap;
  if (tma_out) {
    const cuuint64_t odim[2] = {(cuuint64_t)Nout, (cuuint64_t)n_out_rows};
    const cuuint64_t ostride[1] = {(cuuint64_t)Nout * 2};
    const cuuint32_t obox[2] = {64, (cuuint32_t)kTileM};
    rc = get_encode()(&tmap_o, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out, odim,
                      ostride, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(out) failed with code " + std::to_string((int)rc));
  }
  const size_t out_stage_bytes = tma_out ? (size_t)Np_full * 256 + 1024 : 0;
  const size_t stage_bytes = kTileM * 128 + (size_t)Np_full * 128;
  // 4 stages measured best (profiles/README.md: 6-8 stages were 5-10 % slower on the memory-bound shapes)
  const int nst = (int)std::max<size_t>(2, std::min<size_t>(4, ((size_t)(224 * 1024) - out_stage_bytes - 2048) / stage_bytes));
  const size_t smem = nst * stage_bytes + 1024 + 256 + out_stage_bytes;
  const unsigned grid = (unsigned)std::min<int64_t>(max_tiles * NB, sm_count());
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == kBF16) {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(seg_mm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seg_mm_tc_kernel<true><<<grid, kThreads, smem, st>>>(tmap_a, tmap, tmap_o, tma_a, tma_out, (const uint16_t*)A, lda, (const int64_t*)idx_a, (const int64_t*)seg_off, (const int64_t*)tile_off, R, K, Nout, NB, out, out_f32, nst);
  } else {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(seg_mm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seg_mm_tc_kernel<false><<<grid, kThreads, smem, st>>>(tmap_a, tmap, tmap_o, tma_a, tma_out, (const uint16_t*)A, lda, (const int64_t*)idx_a, (const int64_t*)seg_off, (const int64_t*)tile_off, R, K, Nout, NB, out, out_f32, nst);
  }
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// dW: [R, K, Nout] fp32, accumulated into (caller zeroes it). item_off: [R+1] int64 prefix of ceil(len / 2048).
DGLB_API int dglb200_cuda_segment_mm_dw_tc(int dtype, const void* A, int64_t lda, const void* idx_a, const void* dO,
                                           int64_t n_do_rows, int64_t ldo, const void* seg_off, const void* item_off, int64_t max_items,
                                           int R, int K, int Nout, void* dW, void* stream) {
  DGLB_TRY
  DGLB_CHECK(dtype == kBF16 || dtype == kF16, "segment_mm_dw_tc: bf16 / fp16 inputs only");
  DGLB_CHECK(K % 8 == 0 && Nout % 8 == 0 && lda % 8 == 0 && ldo % 8 == 0, "segment_mm_dw_tc: K, Nout and row strides must be multiples of 8");
  DGLB_CHECK(((uintptr_t)A % 16 == 0) && ((uintptr_t)dO % 16 == 0) && ((uintptr_t)dW % 16 == 0), "16-byte alignment required");
  if (max_items <= 0 || R <= 0) return 0;
  const int NB = (Nout + 255) / 256, KB = (K + 127) / 128;
  const int Np_full = Nout < 256 ? ((Nout + 15) & ~15) : 256;
  const int n_boxes = (Np_full + 63) / 64;
  const size_t stage_bytes = (size_t)(2 + n_boxes) * kChunkK * 128;
  const size_t smem = kStages * stage_bytes + 1024 + 256;
  // dO through TMA: needs a dense row stride that is a multiple of 16 bytes (always true here) - boxes of [64 rows x 64 cols]
  const int tma_do = 1;
  CUtensorMap tmap_do;
  {
    const cuuint64_t ddim[2] = {(cuuint64_t)Nout, (cuuint64_t)n_do_rows};
    const cuuint64_t dstride[1] = {(cuuint64_t)ldo * 2};
    const cuuint32_t dbox[2] = {64, (cuuint32_t)kChunkK};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = get_encode()(&tmap_do, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                               const_cast<void*>(dO), ddim, dstride, dbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(dO) failed with code " + std::to_string((int)rc));
  }
  const unsigned grid = (unsigned)std::min<int64_t>(max_items * NB * KB, sm_count());
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == kBF16) {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(seg_mm_dw_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seg_mm_dw_tc_kernel<true><<<grid, kThreads, smem, st>>>(tmap_do, tma_do, (const uint16_t*)A, lda, (const int64_t*)idx_a, (const uint16_t*)dO, ldo, (const int64_t*)seg_off, (const int64_t*)item_off, R, K, Nout, KB, NB, (float*)dW);
  } else {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(seg_mm_dw_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seg_mm_dw_tc_kernel<false><<<grid, kThreads, smem, st>>>(tmap_do, tma_do, (const uint16_t*)A, lda, (const int64_t*)idx_a, (const uint16_t*)dO, ldo, (const int64_t*)seg_off, (const int64_t*)item_off, R, K, Nout, KB, NB, (float*)dW);
  }
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
