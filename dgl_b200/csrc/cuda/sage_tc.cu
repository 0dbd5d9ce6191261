// Fused GraphSAGE layer on 5th-gen tensor cores (sm_100a):
//
//   out[d, :] = act( [ X[self(d)] | mean_k X[nbr(d,k)] ] @ W + b )          d < *n_dev  (device-side row count)
//
// Persistent, warp-specialised kernel: one CTA per SM walks output tiles (tile = blockIdx.x, += gridDim.x).
//   * 16 producer warps GATHER the self row and the sampled neighbour rows straight from the feature table
//     (16-byte lanes, up to 12 independent row loads in flight per lane, ~100 KB in flight per SM), reduce the
//     neighbour mean in fp32 registers and write the bf16 A tile into shared memory in the canonical K-major
//     SWIZZLE_128B UMMA layout. The A operand is never read from global memory as a matrix. The index lists of
//     the NEXT tile are prefetched into shared memory with cp.async while the current tile is gathered.
//   * one TMA thread streams the W^T chunks (cp.async.bulk.tensor, 128B swizzle) through the same mbarrier ring;
//   * one elected thread issues tcgen05.mma (M=128, N<=256, K=16 per instruction) into one of TWO TMEM accumulators;
//   * 4 epilogue warps drain the other accumulator (tcgen05.ld -> +bias -> ReLU -> dropout -> bf16 -> global) while
//     the producers are already gathering the next tile, so neither the epilogue nor the per-tile set-up sits on the
//     memory-bound critical path.
// The MMA is far from the bottleneck (the layer is gather-bound), so when there are few rows a tile carries only
// 64 live rows of the 128-row MMA tile: twice as many tiles -> all SMs pull on HBM and wave quantisation halves.
// The A tile can optionally be spilled to global (A_out) for the backward dW GEMM.
//
// This fuses what the reference runs as: index_select (feature gather) + g-SpMM copy_u mean + 2 cuBLAS GEMMs +
// bias/activation/dropout elementwise kernels (SURVEY.md 2.4 "Fusion status", sageconv.py:230-287).
#include "tc_common.cuh"

namespace dglb {

constexpr int kTileM = 128;
constexpr int kMaxStages = 4;
// One K chunk is one 128-byte swizzle-atom row: 64 bf16 or 32 fp32 elements (TcT<T>::kRowElems). With T = float the
// operands are split x = hi + lo on the fly (producers for A, the optimizer epilogue for W) and every K step issues the
// three TF32 products lo*hi + hi*lo + hi*hi (tc_common.cuh), so an fp32 model keeps fp32-level accuracy on tcgen05.

struct AccRow {
  float v[8];
};

template <typename T>
__device__ __forceinline__ void acc_add(AccRow& a, const uint4& q) {
  if constexpr (sizeof(T) == 2) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      a.v[2 * i] += f.x;
      a.v[2 * i + 1] += f.y;
    }
  } else {
    a.v[0] += __uint_as_float(q.x); a.v[1] += __uint_as_float(q.y);
    a.v[2] += __uint_as_float(q.z); a.v[3] += __uint_as_float(q.w);
  }
}

constexpr int kProducerWarps = 16;
constexpr int kEpilogueWarps = 8;
constexpr int kTmaWarp = kProducerWarps + kEpilogueWarps;
constexpr int kMmaWarp = kTmaWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 832
constexpr int kGatherBatch = 10;               // independent 16-byte row loads in flight per producer lane

__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory"); }

__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

template <bool RELU, typename T>
__global__ void __launch_bounds__(kThreads, 1)
sage_layer_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_wlo,
                     const T* __restrict__ X,
                     const int32_t* __restrict__ self_ids, const int32_t* __restrict__ nbr,
                     const int32_t* __restrict__ cnt, const int64_t* __restrict__ n_dev, int64_t n_cap, int fanout,
                     int Fp, int Np, const float* __restrict__ bias, T* __restrict__ out,
                     T* __restrict__ A_out, float p_drop, uint64_t seed,
                     const int64_t* __restrict__ step_dev, int kStages, int R, int out_f32, long long* __restrict__ trace) {
  constexpr int kChunkK = TcT<T>::kRowElems;   // elements per 128-byte K chunk
  constexpr int kPlanes = TcT<T>::kPlanes;     // 1 (bf16) or 2 (fp32: hi | lo)
  constexpr int kVecE = 16 / (int)sizeof(T);   // elements per 16-byte lane vector
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t n = min(*n_dev, n_cap);
  if ((int64_t)blockIdx.x * R >= n) return;  // device-side M: CTAs past the live row count do no work at all

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KH = Fp / kChunkK;  // chunks per half (self | neigh)
  const int KC = 2 * KH;
  const int Kp = 2 * Fp;
  const uint32_t a_bytes = kTileM * 128;
  const uint32_t b_bytes = (uint32_t)Np * 128;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // stage s: [A plane 0 | (A lo) | B plane 0 | (B lo)] at smem + s*stage_bytes
  const uint32_t stage_bytes = kPlanes * (a_bytes + b_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  // bars: full[4] | empty[4] | acc_full[2] | acc_empty[2] ; tmem base holder ; 2 x index lists (cnt | self | nbr)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  float* s_bias = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);  // [256] bias * dropout rescale (16-B aligned)
  int32_t* idx_base = reinterpret_cast<int32_t*>(s_bias + 256);
  const int idx_words = kTileM * (2 + fanout);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages);
  const uint32_t accf0 = smem_u32(bars + 2 * kMaxStages), acce0 = smem_u32(bars + 2 * kMaxStages + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, kProducerWarps + 1);  // producer warps + the TMA thread
      mbar_init(empty0 + 8 * s, 1);                  // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf0 + 8 * a, 1);               // tcgen05.commit after the tile's last MMA
      mbar_init(acce0 + 8 * a, kEpilogueWarps);  // every epilogue warp has drained the accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t acc_cols = Np <= 32 ? 32 : (Np <= 64 ? 64 : (Np <= 128 ? 128 : 256));
  const uint32_t tmem_cols = 2 * acc_cols;  // double-buffered accumulator
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // optional cycle-stamp trace of CTA 0 (debug / profiling aid; null in production)
#define SAGE_TRACE(slot, cond) do { if (trace && blockIdx.x == 0 && (cond)) trace[(slot)] = clock64(); } while (0)
  SAGE_TRACE(0, threadIdx.x == 0);
  const int64_t tile_stride = (int64_t)gridDim.x * R;
  const int gmax = kStages >= 4 ? 2 : 1;  // chunks per column group: a group occupies 2 * gmax ring slots at once

  if (warp < kProducerWarps) {
    // ===================== producers: gather + mean -> swizzled smem A tile =====================
    const int tid = threadIdx.x;  // 0..511

    // index lists of a tile -> smem buffer `buf` (asynchronously; rows past the live count get cnt = 0)
    auto stage_indices = [&](int64_t t0, int buf) {
      int32_t* s_cnt = idx_base + buf * idx_words;
      int32_t* s_self = s_cnt + kTileM;
      int32_t* s_nbr = s_self + kTileM;
      const int rows = (int)max((int64_t)0, min((int64_t)R, n - t0));
      for (int i = tid; i < R; i += kProducerWarps * 32) {
        if (i < rows) {
          cp_async4(smem_u32(s_cnt + i), cnt + t0 + i);
          if (self_ids) cp_async4(smem_u32(s_self + i), self_ids + t0 + i);
          else s_self[i] = (int32_t)(t0 + i);
        } else {
          s_cnt[i] = 0;
          s_self[i] = 0;
        }
      }
      const int32_t* nb_g = nbr + t0 * fanout;
      for (int i = tid; i < rows * fanout; i += kProducerWarps * 32) cp_async4(smem_u32(s_nbr + i), nb_g + i);
    };

    stage_indices((int64_t)blockIdx.x * R, 0);
    uint32_t g = 0;  // running chunk counter across tiles -> smem ring slot + parity
    int it = 0;
    for (int64_t tile0 = (int64_t)blockIdx.x * R; tile0 < n; tile0 += tile_stride, ++it) {
      cp_async_wait_all();
      producer_bar();  // this tile's lists are visible; everyone is done reading the other buffer
      if (tile0 + tile_stride < n) stage_indices(tile0 + tile_stride, (it + 1) & 1);
      SAGE_TRACE(144 + it, tid == 0 && it < 8);
      const int32_t* s_cnt = idx_base + (it & 1) * idx_words;
      const int32_t* s_self = s_cnt + kTileM;
      const int32_t* s_nbr = s_self + kTileM;
      const int rows_here = (int)min((int64_t)R, n - tile0);
      // K is walked in column groups of up to two 64-column chunks; for each group the SELF slice and the NEIGHBOUR
      // slice of the same columns are produced together (ring order: self chunks of the group, then its neighbour
      // chunks - the TMA warp streams W^T in the same order). A lane group reads 256 contiguous bytes of a feature row
      // at once (DRAM page / sector locality) and the one self-row load of a pass is in flight together with the
      // neighbour-row loads instead of costing a separate latency round trip.
      for (int k0 = 0; k0 < KH; k0 += gmax) {
        const int nch = min(gmax, KH - k0);
        const int lpr_shift = nch == 2 ? 4 : 3;  // lanes per row piece: 16 or 8
        const int rpp = (kProducerWarps * 32) >> lpr_shift;  // rows per pass: 32 or 64
        const int piece = tid & ((1 << lpr_shift) - 1);
        const int r0 = tid >> lpr_shift;
        const int kcl = piece >> 3, c = piece & 7;
        for (int u = 0; u < 2 * nch; ++u) mbar_wait(empty0 + 8 * ((g + u) % kStages), (((g + u) / kStages) & 1) ^ 1);
        uint8_t* sa_self = smem + ((g + kcl) % kStages) * stage_bytes;
        uint8_t* sa_nbr = smem + ((g + nch + kcl) % kStages) * stage_bytes;
        const int col = (k0 + kcl) * kChunkK + c * kVecE;  // column of this lane's 16 bytes inside a half
        const T* Xc = X + col;
        T* Ao = A_out ? A_out + col : nullptr;
        const int npass = R / rpp;  // 1..4
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
          const int r = pass * rpp + r0;
          const bool live = r < rows_here;
          const int cn = s_cnt[r];  // 0 for rows past the live count
          const int32_t* nb = s_nbr + r * fanout;
          uint4 sv = make_uint4(0, 0, 0, 0);
          if (live) sv = __ldg(reinterpret_cast<const uint4*>(Xc + (int64_t)s_self[r] * Fp));
          AccRow acc;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
          for (int j0 = 0; j0 < cn; j0 += kGatherBatch) {
            uint4 q[kGatherBatch];
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j)
              q[j] = (j0 + j < cn) ? __ldg(reinterpret_cast<const uint4*>(Xc + (int64_t)max(nb[j0 + j], 0) * Fp))
                                   : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j) acc_add<T>(acc, q[j]);
          }
          const float inv = cn > 0 ? 1.f / (float)cn : 0.f;
          uint4 val;
          if constexpr (sizeof(T) == 2) {
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc.v[2 * i] * inv, acc.v[2 * i + 1] * inv);
          } else {
            val = make_uint4(__float_as_uint(acc.v[0] * inv), __float_as_uint(acc.v[1] * inv),
                             __float_as_uint(acc.v[2] * inv), __float_as_uint(acc.v[3] * inv));
          }
          if (Ao && live) {
            *reinterpret_cast<uint4*>(Ao + (tile0 + r) * Kp) = sv;
            *reinterpret_cast<uint4*>(Ao + (tile0 + r) * Kp + Fp) = val;
          }
          if constexpr (kPlanes == 1) {
            *reinterpret_cast<uint4*>(sa_self + sw128_off(r, c)) = sv;
            *reinterpret_cast<uint4*>(sa_nbr + sw128_off(r, c)) = val;
          } else {
            float4 hi, lo;
            tf32_split4(*reinterpret_cast<const float4*>(&sv), hi, lo);
            *reinterpret_cast<float4*>(sa_self + sw128_off(r, c)) = hi;
            *reinterpret_cast<float4*>(sa_self + a_bytes + sw128_off(r, c)) = lo;
            tf32_split4(*reinterpret_cast<const float4*>(&val), hi, lo);
            *reinterpret_cast<float4*>(sa_nbr + sw128_off(r, c)) = hi;
            *reinterpret_cast<float4*>(sa_nbr + a_bytes + sw128_off(r, c)) = lo;
          }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0)
          for (int u = 0; u < 2 * nch; ++u) mbar_arrive(full0 + 8 * ((g + u) % kStages));
        SAGE_TRACE(16 + it * 16 + k0, tid == 0 && it < 8 && k0 < 16);
        g += 2 * nch;
      }
    }
  } else if (warp < kTmaWarp) {
    // ===================== epilogue: TMEM -> registers -> bias/ReLU/dropout -> bf16 -> global =====
    // warp w may touch TMEM lanes 32*(w%4)..+31 (kProducerWarps % 4 == 0, so quarter = warp & 3); the two warps
    // of a quarter take alternate 32-column blocks.
    const int ew = warp - kProducerWarps;
    const int quarter = ew & 3;
    const int row = quarter * 32 + lane;
    const uint64_t sd = seed + (step_dev ? (uint64_t)(*step_dev) * 0x9E3779B97F4A7C15ull : 0ull);
    const uint32_t sd32 = fmix32((uint32_t)sd ^ fmix32((uint32_t)(sd >> 32)));
    // drop probability quantised to 1/256 (the rescale uses the quantised value, so the estimator stays unbiased)
    const uint32_t thresh = p_drop > 0.f ? min(255u, (uint32_t)(fminf(p_drop, 1.f) * 256.f + 0.5f)) : 0u;
    const uint32_t thresh4 = thresh * 0x01010101u;
    const float keep_scale = 256.f / (256.f - (float)thresh);
    // bias pre-multiplied by the dropout rescale: relu(x + b) * s == relu(x * s + b * s)
    for (int i = ew * 32 + lane; i < Np; i += kEpilogueWarps * 32) s_bias[i] = bias ? bias[i] * keep_scale : 0.f;
    asm volatile("bar.sync 2, %0;" ::"n"(kEpilogueWarps * 32) : "memory");
    int it = 0;
    for (int64_t tile0 = (int64_t)blockIdx.x * R; tile0 < n; tile0 += tile_stride, ++it) {
      const int acc = it & 1;
      SAGE_TRACE(300 + it * 4, ew == 0 && lane == 0 && it < 8);
      mbar_wait(accf0 + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      SAGE_TRACE(301 + it * 4, ew == 0 && lane == 0 && it < 8);
      const int64_t grow = tile0 + row;
      const bool live = row < R && grow < n;
      const uint32_t sd_row = sd32 ^ ((uint32_t)((uint64_t)grow >> 20) * 0x85ebca77u);
      if (quarter * 32 < R) {  // warp-uniform: quarters past the live rows of a 64-row tile have nothing to drain
        for (int n0 = (ew >> 2) * 32; n0 < Np; n0 += 32 * (kEpilogueWarps / 4)) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + n0), v);
          if (live && out_f32) {  // fp32 output (logits of an output layer): bias (+ReLU), no dropout, no rounding
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + grow * Np + n0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sb = *reinterpret_cast<const float4*>(s_bias + n0 + j);
              float4 x = make_float4(__uint_as_float(v[j]) + sb.x, __uint_as_float(v[j + 1]) + sb.y,
                                     __uint_as_float(v[j + 2]) + sb.z, __uint_as_float(v[j + 3]) + sb.w);
              if (RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
              dst[j >> 2] = x;
            }
          } else if (live && sizeof(T) == 4) {
            // fp32 activations: bias + ReLU + dropout (same per-element hash as the bf16 path), 16-byte stores
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + grow * Np + n0);
            const uint32_t e0 = (uint32_t)(((uint64_t)grow * (uint64_t)Np + (uint64_t)n0) >> 2);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sb = *reinterpret_cast<const float4*>(s_bias + n0 + j);
              float4 x = make_float4(fmaf(__uint_as_float(v[j]), keep_scale, sb.x), fmaf(__uint_as_float(v[j + 1]), keep_scale, sb.y),
                                     fmaf(__uint_as_float(v[j + 2]), keep_scale, sb.z), fmaf(__uint_as_float(v[j + 3]), keep_scale, sb.w));
              if (RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
              if (thresh) {
                const uint32_t rb = fmix32(((e0 + (uint32_t)(j >> 2)) * 0x9E3779B9u) ^ sd_row);
                const uint32_t drop = __vcmpltu4(rb, thresh4);
                if (drop & 0x000000ffu) x.x = 0.f;
                if (drop & 0x0000ff00u) x.y = 0.f;
                if (drop & 0x00ff0000u) x.z = 0.f;
                if (drop & 0xff000000u) x.w = 0.f;
              }
              dst[j >> 2] = x;
            }
          } else if (live) {
            uint32_t packed[16];
            const uint32_t e0 = (uint32_t)(((uint64_t)grow * (uint64_t)Np + (uint64_t)n0) >> 2);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sb = *reinterpret_cast<const float4*>(s_bias + n0 + j);
              float x0 = fmaf(__uint_as_float(v[j]), keep_scale, sb.x);
              float x1 = fmaf(__uint_as_float(v[j + 1]), keep_scale, sb.y);
              float x2 = fmaf(__uint_as_float(v[j + 2]), keep_scale, sb.z);
              float x3 = fmaf(__uint_as_float(v[j + 3]), keep_scale, sb.w);
              if (RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
              __nv_bfloat162 h01 = __floats2bfloat162_rn(x0, x1), h23 = __floats2bfloat162_rn(x2, x3);
              uint32_t p01 = *reinterpret_cast<uint32_t*>(&h01), p23 = *reinterpret_cast<uint32_t*>(&h23);
              if (thresh) {  // one 32-bit hash -> four 8-bit keep/drop decisions, applied as bit masks on the bf16 pairs
                const uint32_t rb = fmix32(((e0 + (uint32_t)(j >> 2)) * 0x9E3779B9u) ^ sd_row);
                const uint32_t drop = __vcmpltu4(rb, thresh4);  // 0xff in every byte that is below the threshold
                p01 &= ~__byte_perm(drop, 0, 0x1100);
                p23 &= ~__byte_perm(drop, 0, 0x3322);
              }
              packed[j >> 1] = p01;
              packed[(j >> 1) + 1] = p23;
            }
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + grow * Np + n0);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce0 + 8 * acc);  // accumulator may be overwritten by tile it+2
      SAGE_TRACE(302 + it * 4, ew == 0 && lane == 0 && it < 8);
    }
  } else if (warp == kTmaWarp) {
    // ===================== TMA producer for the W^T tiles =====================
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t tile0 = (int64_t)blockIdx.x * R; tile0 < n; tile0 += tile_stride) {
        for (int k0 = 0; k0 < KH; k0 += gmax) {  // ring order of the producers: [self chunks | neighbour chunks] per group
          const int nch = min(gmax, KH - k0);
          for (int u = 0; u < 2 * nch; ++u, ++g) {
            const int kc = (u < nch ? 0 : KH) + k0 + (u < nch ? u : u - nch);  // K chunk carried by this ring slot
            const int stage = g % kStages;
            const uint32_t phase = (g / kStages) & 1;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            mbar_arrive_expect_tx(full0 + 8 * stage, kPlanes * b_bytes);
            const uint32_t sb = smem_u32(smem + stage * stage_bytes + kPlanes * a_bytes);
            tma_load_2d(sb, &tmap_w, full0 + 8 * stage, kc * kChunkK, 0);
            if constexpr (kPlanes == 2) tma_load_2d(sb + b_bytes, &tmap_wlo, full0 + 8 * stage, kc * kChunkK, 0);
          }
        }
      }
    }
  } else {
    // ===================== MMA issuer (one elected thread) =====================
    // instruction descriptor: D=F32, A=B=BF16, both K-major, N = Np, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)TcT<T>::kFmt << 7) | ((uint32_t)TcT<T>::kFmt << 10) |
                           ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    auto kdesc = [](uint32_t a) { return make_sw128_kmajor_desc(a); };
    uint32_t g = 0;
    int it = 0;
    for (int64_t tile0 = (int64_t)blockIdx.x * R; tile0 < n; tile0 += tile_stride, ++it) {
      const int acc = it & 1;
      mbar_wait(acce0 + 8 * acc, ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator (free on first use)
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * acc_cols);
      for (int kc = 0; kc < KC; ++kc, ++g) {
        const int stage = g % kStages;
        const uint32_t phase = (g / kStages) & 1;
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(smem + stage * stage_bytes), b0 = a0 + kPlanes * a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 K steps of 32 bytes per 128-byte chunk
            umma_step<T>(tmem_d, a0 + k * 32, a_bytes, kdesc, b0 + k * 32, b_bytes, kdesc, idesc, (kc | k) != 0);
          umma_commit(empty0 + 8 * stage);                    // smem slot reusable once these MMAs retire
          if (kc == KC - 1) umma_commit(accf0 + 8 * acc);     // accumulator complete -> epilogue
          SAGE_TRACE(160 + it * 16 + kc, it < 8 && kc < 16);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  SAGE_TRACE(1, threadIdx.x == 0);
#undef SAGE_TRACE
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

}  // namespace dglb

using namespace dglb;

namespace {

template <typename T>
void launch_sage_layer(const void* X, const void* self_ids, const void* nbr, const void* cnt, const void* n_dev,
                       int64_t n_cap, int fanout, int Fp, int Np, const void* Wt, const void* Wt_lo, const void* bias,
                       void* out, void* A_out, int relu, float p_drop, int64_t seed, const void* step_dev, int out_f32,
                       void* trace, cudaStream_t st) {
  constexpr int kChunkK = TcT<T>::kRowElems;
  constexpr int kPlanes = TcT<T>::kPlanes;
  DGLB_CHECK(Fp % kChunkK == 0 && Fp >= kChunkK, "Fp must be a multiple of one 128-byte K chunk");
  DGLB_CHECK(kPlanes == 1 || Wt_lo != nullptr, "fp32 layers need the lo plane of the split weights");
  CUtensorMap tmap, tmap_lo;
  const cuuint64_t gdim[2] = {(cuuint64_t)(2 * Fp), (cuuint64_t)Np};
  const cuuint64_t gstride[1] = {(cuuint64_t)(2 * Fp) * sizeof(T)};
  const cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)Np};
  const cuuint32_t estr[2] = {1, 1};
  for (int pl = 0; pl < kPlanes; ++pl) {
    CUresult r = get_encode()(pl ? &tmap_lo : &tmap, TcT<T>::kTmap, 2, const_cast<void*>(pl ? Wt_lo : Wt), gdim, gstride, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  }
  if (kPlanes == 1) tmap_lo = tmap;
  const size_t stage_bytes = kPlanes * (kTileM * 128 + (size_t)Np * 128);
  const size_t fixed = 1024 + 256 + 1024 + 2 * (size_t)kTileM * 4 * (2 + fanout);
  DGLB_CHECK(fixed + 2 * stage_bytes <= 227 * 1024, "fanout too large for the fused layer kernel");
  const int stages = (int)std::min<size_t>(kMaxStages, (227 * 1024 - fixed) / stage_bytes);
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    DGLB_CUDA_CHECK(cudaGetDevice(&dev));
    DGLB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // gather-bound layer: with fewer than 8 waves of 128-row tiles use 64 live rows per (128-row) MMA tile
  const int64_t tiles128 = (n_cap + kTileM - 1) / kTileM;
  const int R = tiles128 < 8 * (int64_t)num_sms ? 64 : 128;
  const unsigned grid = (unsigned)std::min<int64_t>((n_cap + R - 1) / R, num_sms);
  auto go = [&](auto kern) {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(tmap, tmap_lo, (const T*)X, (const int32_t*)self_ids, (const int32_t*)nbr,
                                       (const int32_t*)cnt, (const int64_t*)n_dev, n_cap, fanout, Fp, Np,
                                       (const float*)bias, (T*)out, (T*)A_out, p_drop, (uint64_t)seed,
                                       (const int64_t*)step_dev, stages, R, out_f32, (long long*)trace);
  };
  if (relu) go(sage_layer_tc_kernel<true, T>);
  else go(sage_layer_tc_kernel<false, T>);
  DGLB_LAUNCH_CHECK();
}

}  // namespace

// dtype kBF16: X [rows, Fp] bf16, Wt [Np, 2*Fp] bf16 (K contiguous), out [n_cap, Np] bf16 (fp32 when out_f32),
//              A_out [n_cap, 2*Fp] bf16 or null.
// dtype kF32 : the same tensors in fp32; Wt / Wt_lo are the hi / lo planes of the TF32 split of the weights.
DGLB_API int dglb200_engine_sage_layer_tc(int dtype, const void* X, const void* self_ids, const void* nbr, const void* cnt,
                                          const void* n_dev, int64_t n_cap, int fanout, int Fp, int Np,
                                          const void* Wt, const void* Wt_lo, const void* bias, void* out, void* A_out,
                                          int relu, float p_drop, int64_t seed, const void* step_dev, int out_f32,
                                          void* trace, void* stream) {
  DGLB_TRY
  DGLB_CHECK(dtype == kBF16 || dtype == kF32, "sage_layer_tc: bf16 or fp32");
  DGLB_CHECK(Np % 32 == 0 && Np >= 32 && Np <= 256, "Np must be a multiple of 32 in [32, 256]");
  DGLB_CHECK(((uintptr_t)X % 16 == 0) && ((uintptr_t)Wt % 16 == 0) && ((uintptr_t)out % 16 == 0), "16-byte alignment required");
  DGLB_CHECK(!out_f32 || p_drop == 0.f, "fp32 logits output does not support dropout");
  if (n_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == kBF16)
    launch_sage_layer<__nv_bfloat16>(X, self_ids, nbr, cnt, n_dev, n_cap, fanout, Fp, Np, Wt, Wt_lo, bias, out, A_out, relu,
                                     p_drop, seed, step_dev, out_f32, trace, st);
  else
    launch_sage_layer<float>(X, self_ids, nbr, cnt, n_dev, n_cap, fanout, Fp, Np, Wt, Wt_lo, bias, out, A_out, relu, p_drop,
                             seed, step_dev, out_f32, trace, st);
  DGLB_CATCH
}
