// Fused GraphSAGE layer on 5th-gen tensor cores (sm_100a):
//
//   out[d, :] = act( [ X[self(d)] | mean_k X[nbr(d,k)] ] @ W + b )          d < *n_dev  (device-side row count)
//
// One CTA = one 128-row output tile. The A operand is never read from global memory as a matrix: four
// producer warps GATHER the self row and the sampled neighbour rows straight from the feature table
// (16-byte lanes, up to 2*fanout independent loads in flight per lane), reduce the neighbour mean in fp32
// registers and write the bf16 tile into shared memory in the canonical K-major SWIZZLE_128B UMMA layout.
// W^T tiles arrive by TMA (cp.async.bulk.tensor, 128B swizzle) on the same mbarrier pipeline. One elected
// thread issues tcgen05.mma (M=128, N<=256, K=16 per instruction) with the accumulator in TMEM; the producer
// warps then become the epilogue: tcgen05.ld -> +bias -> ReLU -> dropout -> bf16 -> global.
// The tile of A can optionally be spilled to global (A_out) for the backward dW GEMM.
//
// This fuses what the reference runs as: index_select (feature gather) + g-SpMM copy_u mean + 2 cuBLAS GEMMs +
// bias/activation/dropout elementwise kernels (SURVEY.md 2.4 "Fusion status", sageconv.py:230-287).
#include <cuda.h>

#include "common.cuh"

namespace dglb {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;  // bf16 elements per K chunk = 128 bytes = one swizzle atom row
constexpr int kStages = 2;
constexpr int kProducerThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct AccRow {
  float v[8];
};

__device__ __forceinline__ void acc_add(AccRow& a, const uint4& q) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    a.v[2 * i] += f.x;
    a.v[2 * i + 1] += f.y;
  }
}

// swizzled byte offset of (row r, 16-byte chunk c) inside a [rows x 128 B] K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

template <bool RELU>
__global__ void __launch_bounds__(192, 2)
sage_layer_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __nv_bfloat16* __restrict__ X,
                     const int32_t* __restrict__ self_ids, const int32_t* __restrict__ nbr,
                     const int32_t* __restrict__ cnt, const int64_t* __restrict__ n_dev, int64_t n_cap, int fanout,
                     int Fp, int Np, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                     __nv_bfloat16* __restrict__ A_out, float p_drop, uint64_t seed,
                     const int64_t* __restrict__ step_dev) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t n = min(*n_dev, n_cap);
  const int64_t tile0 = (int64_t)blockIdx.x * kTileM;
  if (tile0 >= n) return;  // device-side M: tiles past the live row count do no work at all

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KH = Fp / kChunkK;  // chunks per half (self | neigh)
  const int KC = 2 * KH;
  const int Kp = 2 * Fp;
  const uint32_t a_bytes = kTileM * 128;
  const uint32_t b_bytes = (uint32_t)Np * 128;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA[kStages];
  uint8_t* sB[kStages];
#pragma unroll
  for (int s = 0; s < kStages; ++s) {
    sA[s] = smem + s * (a_bytes + b_bytes);
    sB[s] = sA[s] + a_bytes;
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (a_bytes + b_bytes));
  // bars[0..S) full, bars[S..2S) empty, bars[2S] tmem_full ; then tmem base holder
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages), tfull = smem_u32(bars + 2 * kStages);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 4 + 1);  // 4 producer warps + the TMA thread
      mbar_init(empty0 + 8 * s, 1);     // one tcgen05.commit
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem_cols = Np <= 32 ? 32 : (Np <= 64 ? 64 : (Np <= 128 ? 128 : 256));
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < 4) {
    // ===================== producers: gather + mean -> swizzled smem A tile =====================
    const int tid = threadIdx.x;       // 0..127
    const int c = tid & 7;             // 16-byte chunk inside the 128-byte row slice
    const int r0 = tid >> 3;           // 0..15
    for (int kc = 0; kc < KC; ++kc) {
      const int stage = kc % kStages;
      const uint32_t phase = (kc / kStages) & 1;
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      const bool self_half = kc < KH;
      const int col0 = (self_half ? kc : kc - KH) * kChunkK + c * 8;  // element offset inside the X row
#pragma unroll 2
      for (int pass = 0; pass < kTileM / 16; ++pass) {
        const int r = pass * 16 + r0;
        const int64_t grow = tile0 + r;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (grow < n) {
          if (self_half) {
            const int64_t sid = self_ids ? (int64_t)self_ids[grow] : grow;
            val = __ldg(reinterpret_cast<const uint4*>(X + sid * Fp + col0));
          } else {
            const int cn = cnt[grow];
            const int32_t* nb = nbr + grow * fanout;
            AccRow acc;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
            int k = 0;
            for (; k + 4 <= cn; k += 4) {
              const int64_t u0 = max(nb[k], 0), u1 = max(nb[k + 1], 0), u2 = max(nb[k + 2], 0), u3 = max(nb[k + 3], 0);
              const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(X + u0 * Fp + col0));
              const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(X + u1 * Fp + col0));
              const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(X + u2 * Fp + col0));
              const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(X + u3 * Fp + col0));
              acc_add(acc, q0); acc_add(acc, q1); acc_add(acc, q2); acc_add(acc, q3);
            }
            for (; k < cn; ++k) {
              const int64_t u = max(nb[k], 0);
              acc_add(acc, __ldg(reinterpret_cast<const uint4*>(X + u * Fp + col0)));
            }
            const float inv = cn > 0 ? 1.f / (float)cn : 0.f;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc.v[2 * i] * inv, acc.v[2 * i + 1] * inv);
          }
          if (A_out) *reinterpret_cast<uint4*>(A_out + grow * Kp + kc * kChunkK + c * 8) = val;
        }
        *reinterpret_cast<uint4*>(sA[stage] + sw128_off(r, c)) = val;
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + 8 * stage);
    }
    // ===================== epilogue: TMEM -> registers -> bias/ReLU/dropout -> bf16 -> global =====
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int64_t grow = tile0 + row;
    const uint64_t sd = seed + (step_dev ? (uint64_t)(*step_dev) * 0x9E3779B97F4A7C15ull : 0ull);
    const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
    for (int n0 = 0; n0 < Np; n0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
      if (grow < n) {
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float a = __uint_as_float(v[j]) + bias[n0 + j];
          float b = __uint_as_float(v[j + 1]) + bias[n0 + j + 1];
          if (RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
          if (p_drop > 0.f) {
            const uint64_t rb = rng_u64(sd, 0x51ull, (uint64_t)(grow * Np + n0 + j));
            const float ua = (float)(rb & 0xffffff) * (1.f / 16777216.f);
            const float ub = (float)((rb >> 24) & 0xffffff) * (1.f / 16777216.f);
            a = ua < p_drop ? 0.f : a * keep_scale;
            b = ub < p_drop ? 0.f : b * keep_scale;
          }
          __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
          packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + grow * Np + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ===================== TMA producer for the W^T tiles =====================
    if (lane == 0) {
      for (int kc = 0; kc < KC; ++kc) {
        const int stage = kc % kStages;
        const uint32_t phase = (kc / kStages) & 1;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        mbar_arrive_expect_tx(full0 + 8 * stage, b_bytes);
        tma_load_2d(smem_u32(sB[stage]), &tmap_w, full0 + 8 * stage, kc * kChunkK, 0);
      }
    }
  } else {
    // ===================== MMA issuer (one elected thread) =====================
    // instruction descriptor: D=F32, A=B=BF16, both K-major, N = Np, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    for (int kc = 0; kc < KC; ++kc) {
      const int stage = kc % kStages;
      const uint32_t phase = (kc / kStages) & 1;
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a0 = smem_u32(sA[stage]), b0 = smem_u32(sB[stage]);
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k) {
          umma_bf16(tmem_base, make_sw128_kmajor_desc(a0 + k * 32), make_sw128_kmajor_desc(b0 + k * 32), idesc,
                    (kc | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * stage);          // smem slot reusable once these MMAs retire
        if (kc == KC - 1) umma_commit(tfull);     // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available from the driver");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace dglb

using namespace dglb;

// X: [rows, Fp] bf16, Wt: [Np, 2*Fp] bf16 (K contiguous), out: [n_cap, Np] bf16, A_out: [n_cap, 2*Fp] bf16 or null
DGLB_API int dglb200_engine_sage_layer_tc(const void* X, const void* self_ids, const void* nbr, const void* cnt,
                                          const void* n_dev, int64_t n_cap, int fanout, int Fp, int Np,
                                          const void* Wt, const void* bias, void* out, void* A_out, int relu,
                                          float p_drop, int64_t seed, const void* step_dev, void* stream) {
  DGLB_TRY
  DGLB_CHECK(Fp % 64 == 0 && Fp >= 64, "Fp must be a multiple of 64");
  DGLB_CHECK(Np % 32 == 0 && Np >= 32 && Np <= 256, "Np must be a multiple of 32 in [32, 256]");
  DGLB_CHECK(((uintptr_t)X % 16 == 0) && ((uintptr_t)Wt % 16 == 0) && ((uintptr_t)out % 16 == 0), "16-byte alignment required");
  if (n_cap == 0) return 0;
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)(2 * Fp), (cuuint64_t)Np};
  const cuuint64_t gstride[1] = {(cuuint64_t)(2 * Fp) * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)Np};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Wt), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  const size_t smem = (size_t)kStages * (kTileM * 128 + (size_t)Np * 128) + 1024 + 128;
  const unsigned grid = (unsigned)((n_cap + kTileM - 1) / kTileM);
  cudaStream_t st = (cudaStream_t)stream;
  if (relu) {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(sage_layer_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sage_layer_tc_kernel<true><<<grid, 192, smem, st>>>(tmap, (const __nv_bfloat16*)X, (const int32_t*)self_ids, (const int32_t*)nbr, (const int32_t*)cnt, (const int64_t*)n_dev, n_cap, fanout, Fp, Np, (const float*)bias, (__nv_bfloat16*)out, (__nv_bfloat16*)A_out, p_drop, (uint64_t)seed, (const int64_t*)step_dev);
  } else {
    DGLB_CUDA_CHECK(cudaFuncSetAttribute(sage_layer_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sage_layer_tc_kernel<false><<<grid, 192, smem, st>>>(tmap, (const __nv_bfloat16*)X, (const int32_t*)self_ids, (const int32_t*)nbr, (const int32_t*)cnt, (const int64_t*)n_dev, n_cap, fanout, Fp, Np, (const float*)bias, (__nv_bfloat16*)out, (__nv_bfloat16*)A_out, p_drop, (uint64_t)seed, (const int64_t*)step_dev);
  }
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
