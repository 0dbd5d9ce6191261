// Backward of the fused GraphSAGE layer on 5th-gen tensor cores (sm_100a). Two kernels replace the
// cuBLAS dW / dA GEMMs + relu_bwd + scatter_grad + zero_rows chain of the unfused backward:
//
//   sage_bwd_dw_kernel   D[CP x CQ] += P^T Q,   P = dZ (built on the fly) or A,  Q = A or dZ    (weight gradient)
//       * dZ = dropout/ReLU backward of the incoming fp32 gradient:  dZ = (H > 0) ? dH * scale : 0, built by the
//         producer warps straight into the MN-major SWIZZLE_128B operand image (never round-trips through memory for
//         the GEMM), its column sums (bias gradient) reduced in the same pass, and - when the layer below needs it -
//         written once to global as the A operand of the dA kernel;
//       * both operands are MN-major (the reduction runs over the ROW index), Q arrives by TMA (bf16) or through the
//         producers (fp32, split hi/lo on the fly); split-K over persistent CTAs, each CTA owns a (128 x NBS) block of D
//         in TMEM for its whole row range and adds it to the fp32 gradient with red.global.add.v4.f32 at the end.
//
//   sage_bwd_da_kernel   dA[rows x 2F] = dZ [rows x Np] @ Wt [Np x 2F] and, in the EPILOGUE, the scatter of the layer:
//         dH_below[self(d)]   += dA[d, 0:F]                    (column block 0)
//         dH_below[nbr(d, k)] += dA[d, F:2F] / cnt[d]          (column block 1, k < cnt[d])
//       dZ tiles (K-major) and Wt boxes (MN-major: Wt is stored [Np, 2F], i.e. N contiguous) arrive by TMA, one thread
//       issues tcgen05.mma into one of two TMEM accumulators, 16 epilogue warps drain the other one with tcgen05.ld and
//       issue vector reductions - dA is never materialised.
//
// T = __nv_bfloat16 (kind::f16) or float (kind::tf32, 3-product hi/lo split, see tc_common.cuh).
// Replaces python/dgl/backend/pytorch/sparse.py:195-248 (GSpMM backward) + the cuBLAS GEMMs of nn.Linear backward.
#include "tc_common.cuh"

namespace dglb {
namespace sbw {

constexpr int kProducerWarps = 16;
constexpr int kEpilogueWarps = 8;
constexpr int kTmaWarp = kProducerWarps + kEpilogueWarps;
constexpr int kMmaWarp = kTmaWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 832
constexpr int kMaxStages = 4;

template <typename T>
struct Geo {
  static constexpr int kBC = TcT<T>::kRowElems;                // elements per 128-byte box row (64 / 32)
  static constexpr int kKR = sizeof(T) == 2 ? 64 : 32;         // reduce rows per smem chunk
  static constexpr int kBoxBytes = kKR * 128;                  // one [kKR x 128 B] box
  static constexpr int kEPP = 16 / (int)sizeof(T);             // elements per 16-byte piece
  static constexpr int kMmaRowBytes = TcT<T>::kMmaK * 128;     // MN-major descriptor advance per MMA K step
  static constexpr int kPBytes = 128 * 128;                    // a 128-column MN-major block, one plane
};

// ---------------------------------------------------------------------------------------------------------------
//  dW
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool QTMA>
__global__ void __launch_bounds__(kThreads, 1)
sage_bwd_dw_kernel(const __grid_constant__ CUtensorMap tmap_q, const float* __restrict__ dH, const T* __restrict__ Hmask,
                   const T* __restrict__ P, int ldp, float scale, const T* __restrict__ Q, int ldq,
                   const int64_t* __restrict__ n_dev, int64_t n_cap, int MB, int NB, int NBS, float* __restrict__ out,
                   int64_t ldi, int64_t ldj, float* __restrict__ gb, T* __restrict__ dz_out, float* __restrict__ dz_out_lo,
                   int nst) {
  using G = Geo<T>;
  constexpr int kPlanes = TcT<T>::kPlanes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t n = min(*n_dev, n_cap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Qn = MB * NB;
  const int groups = gridDim.x / Qn;
  const int grp = blockIdx.x / Qn, q = blockIdx.x % Qn;
  const int mb = q / NB, nb = q % NB;
  const int64_t nch = (n + G::kKR - 1) / G::kKR;
  const int64_t per = (nch + groups - 1) / groups;
  const int64_t c_begin = (int64_t)grp * per;
  const int64_t c_end = min(nch, c_begin + per);
  if (grp >= groups || c_begin >= c_end) return;  // whole CTA: nothing to reduce

  const uint32_t q_bytes = (uint32_t)NBS * 128;   // NBS/kBC boxes of [kKR x 128 B]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = kPlanes * (G::kPBytes + q_bytes);  // [P hi | (P lo) | Q hi | (Q lo)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2);
  float* s_gb = reinterpret_cast<float*>(bars + 2 * kMaxStages + 4);  // [128] bias-gradient partials
  constexpr int kScrStride = 20;                                       // epilogue transpose scratch: 8 warps x [32 x 20] fp32
  float* s_scr = s_gb + 128;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages), accf = smem_u32(bars + 2 * kMaxStages);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(full0 + 8 * s, kProducerWarps + (QTMA ? 1 : 0));
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accf, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 128) s_gb[threadIdx.x] = 0.f;
  const uint32_t tmem_cols = NBS <= 32 ? 32 : (NBS <= 64 ? 64 : (NBS <= 128 ? 128 : 256));
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < kProducerWarps) {
    // ===================== producers: P block (dZ built on the fly, or plain rows) and - unless TMA - the Q block ====
    const int tid = threadIdx.x;
    constexpr int PPR = 128 / G::kEPP;          // 16-byte pieces per P row (16 / 32)
    constexpr int RSTEP = 512 / PPR;            // rows covered by one pass of the 512 producer threads (32 / 16)
    constexpr int NPASS = G::kKR / RSTEP;       // 2
    const int piece = tid % PPR, prow = tid / PPR;
    const int pbox = piece >> 3, pc = piece & 7;
    const int pcol = mb * 128 + piece * G::kEPP;            // first element column of this thread's piece
    const int PPRQ = NBS / G::kEPP;                         // pieces per Q row
    const int qpasses = QTMA ? 0 : (G::kKR * PPRQ) / 512;   // 512 % PPRQ == 0 for NBS in {32, 64, 128, 256}
    const int qpiece = tid % PPRQ, qrow0 = tid / PPRQ, qrstep = 512 / PPRQ;
    const int qbox = qpiece >> 3, qc = qpiece & 7;
    const int qcol = nb * NBS + qpiece * G::kEPP;
    float colsum[G::kEPP];
#pragma unroll
    for (int i = 0; i < G::kEPP; ++i) colsum[i] = 0.f;
    const bool side = nb == 0;  // one CTA per (row range, mb) owns the bias gradient and the dZ write-out
    uint32_t g = 0;
    for (int64_t ch = c_begin; ch < c_end; ++ch, ++g) {
      const int64_t row0 = ch * G::kKR;
      // ---- issue every load of this chunk before touching the ring slot
      uint4 pv[NPASS];     // P piece (T)            [transform: built below]
      float4 dh[NPASS][2]; // transform mode: incoming gradient (8 floats for bf16 pieces, 4 for fp32)
#pragma unroll
      for (int h = 0; h < NPASS; ++h) {
        const int64_t r = row0 + prow + h * RSTEP;
        const bool live = r < n;
        if (dH) {
          dh[h][0] = live ? __ldg(reinterpret_cast<const float4*>(dH + r * ldp + pcol)) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (G::kEPP == 8)
            dh[h][1] = live ? __ldg(reinterpret_cast<const float4*>(dH + r * ldp + pcol + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          pv[h] = live ? __ldg(reinterpret_cast<const uint4*>(Hmask + r * ldp + pcol)) : make_uint4(0, 0, 0, 0);
        } else {
          pv[h] = live ? __ldg(reinterpret_cast<const uint4*>(P + r * ldp + pcol)) : make_uint4(0, 0, 0, 0);
        }
      }
      uint4 qv[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (h < qpasses) {
          const int64_t r = row0 + qrow0 + h * qrstep;
          qv[h] = r < n ? __ldg(reinterpret_cast<const uint4*>(Q + r * ldq + qcol)) : make_uint4(0, 0, 0, 0);
        }
      }
      // ---- dropout / ReLU backward in registers
      if (dH) {
#pragma unroll
        for (int h = 0; h < NPASS; ++h) {
          float z[G::kEPP];
          if constexpr (sizeof(T) == 2) {
            const __nv_bfloat162* hm = reinterpret_cast<const __nv_bfloat162*>(&pv[h]);
            const float* d = reinterpret_cast<const float*>(&dh[h][0]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 m = __bfloat1622float2(hm[i]);
              z[2 * i] = m.x > 0.f ? d[2 * i] * scale : 0.f;
              z[2 * i + 1] = m.y > 0.f ? d[2 * i + 1] * scale : 0.f;
            }
            __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&pv[h]);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(z[2 * i], z[2 * i + 1]);
          } else {
            const float* m = reinterpret_cast<const float*>(&pv[h]);
            const float* d = reinterpret_cast<const float*>(&dh[h][0]);
#pragma unroll
            for (int i = 0; i < 4; ++i) z[i] = m[i] > 0.f ? d[i] * scale : 0.f;
            pv[h] = make_uint4(__float_as_uint(z[0]), __float_as_uint(z[1]), __float_as_uint(z[2]), __float_as_uint(z[3]));
          }
#pragma unroll
          for (int i = 0; i < G::kEPP; ++i) colsum[i] += z[i];
        }
      }
      // ---- into the ring slot
      const int stage = g % nst;
      mbar_wait(empty0 + 8 * stage, ((g / nst) & 1) ^ 1);
      uint8_t* sp = smem + stage * stage_bytes;
      uint8_t* sq = sp + kPlanes * G::kPBytes;
#pragma unroll
      for (int h = 0; h < NPASS; ++h) {
        const int rr = prow + h * RSTEP;
        const uint32_t off = pbox * G::kBoxBytes + mn_box_off<T>(rr, pc);
        if constexpr (kPlanes == 1) {
          *reinterpret_cast<uint4*>(sp + off) = pv[h];
        } else {
          float4 hi, lo;
          tf32_split4(*reinterpret_cast<const float4*>(&pv[h]), hi, lo);
          *reinterpret_cast<float4*>(sp + off) = hi;
          *reinterpret_cast<float4*>(sp + G::kPBytes + off) = lo;
          if (dH && dz_out && side) {
            const int64_t r = row0 + rr;
            if (r < n) {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(dz_out) + r * ldp + pcol) = hi;
              *reinterpret_cast<float4*>(dz_out_lo + r * ldp + pcol) = lo;
            }
          }
        }
        if constexpr (kPlanes == 1) {
          if (dH && dz_out && side) {
            const int64_t r = row0 + rr;
            if (r < n) *reinterpret_cast<uint4*>(dz_out + r * ldp + pcol) = pv[h];
          }
        }
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (h < qpasses) {
          const int rr = qrow0 + h * qrstep;
          const uint32_t off = qbox * G::kBoxBytes + mn_box_off<T>(rr, qc);
          if constexpr (kPlanes == 1) {
            *reinterpret_cast<uint4*>(sq + off) = qv[h];
          } else {
            float4 hi, lo;
            tf32_split4(*reinterpret_cast<const float4*>(&qv[h]), hi, lo);
            *reinterpret_cast<float4*>(sq + off) = hi;
            *reinterpret_cast<float4*>(sq + q_bytes + off) = lo;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + 8 * stage);
    }
    // ---- bias gradient: column sums of dZ over this CTA's rows
    if (dH && gb && side) {
#pragma unroll
      for (int i = 0; i < G::kEPP; ++i) atomicAdd(&s_gb[piece * G::kEPP + i], colsum[i]);
      asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory");
      if (tid < 128) {
        const float v = s_gb[tid];
        if (v != 0.f) atomicAdd(gb + mb * 128 + tid, v);
      }
    }
  } else if (warp < kTmaWarp) {
    // ===================== epilogue: D block (TMEM) -> red.add into the fp32 gradient =====================
    // A TMEM lane is one ROW i of D. With a row-major destination (ldj == 1) a warp would touch 32 rows x 16 bytes per
    // reduction, so each [32 x 16] piece is transposed through a per-warp smem scratch and leaves as 8 rows x 64
    // contiguous bytes; with a column-major destination (ldi == 1) lanes are already contiguous in memory.
    const int ew = warp - kProducerWarps;
    const int quarter = ew & 3;
    const int i0 = mb * 128 + quarter * 32;
    float* scr = s_scr + ew * (32 * kScrStride);
    const int sub = lane >> 2, c4 = lane & 3;
    mbar_wait(accf, 0);
    tc_fence_after();
    for (int n0 = (ew >> 2) * 16; n0 < NBS; n0 += 16 * (kEpilogueWarps / 4)) {
      uint32_t v[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)n0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int64_t j0 = (int64_t)nb * NBS + n0;
      if (ldj == 1) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<uint4*>(scr + lane * kScrStride + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 8) {
          const float4 x = *reinterpret_cast<const float4*>(scr + (r0 + sub) * kScrStride + c4 * 4);
          red_add_v4(out + (int64_t)(i0 + r0 + sub) * ldi + j0 + c4 * 4, x.x, x.y, x.z, x.w);
        }
      } else {
        float* dst = out + (int64_t)(i0 + lane) * ldi + j0 * ldj;
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(dst + (int64_t)j * ldj, __uint_as_float(v[j]));
      }
    }
    tc_fence_before();
  } else if (warp == kTmaWarp) {
    if (QTMA && lane == 0) {
      uint32_t g = 0;
      const int nboxes = NBS / G::kBC;
      for (int64_t ch = c_begin; ch < c_end; ++ch, ++g) {
        const int stage = g % nst;
        mbar_wait(empty0 + 8 * stage, ((g / nst) & 1) ^ 1);
        mbar_arrive_expect_tx(full0 + 8 * stage, q_bytes);
        const uint32_t sq = smem_u32(smem + stage * stage_bytes + kPlanes * G::kPBytes);
        for (int b = 0; b < nboxes; ++b)
          tma_load_2d(sq + b * G::kBoxBytes, &tmap_q, full0 + 8 * stage, nb * NBS + b * G::kBC, (int)(ch * G::kKR));
      }
    }
  } else {
    // ===================== MMA issuer: both operands MN-major =====================
    const uint32_t fmt = (uint32_t)TcT<T>::kFmt;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(NBS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto mdesc = [](uint32_t a) { return make_mnmajor_desc<T>(a, (uint32_t)G::kBoxBytes); };
    uint32_t g = 0;
    for (int64_t ch = c_begin; ch < c_end; ++ch, ++g) {
      const int stage = g % nst;
      mbar_wait(full0 + 8 * stage, (g / nst) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t p0 = smem_u32(smem + stage * stage_bytes), q0 = p0 + kPlanes * G::kPBytes;
#pragma unroll
        for (int k = 0; k < G::kKR / TcT<T>::kMmaK; ++k)
          umma_step<T>(tmem_base, p0 + k * G::kMmaRowBytes, G::kPBytes, mdesc, q0 + k * G::kMmaRowBytes, q_bytes, mdesc,
                       idesc, (g | (uint32_t)k) != 0);
        umma_commit(empty0 + 8 * stage);
        if (ch == c_end - 1) umma_commit(accf);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
//  dA + scatter
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDaEpiWarps = 16;
constexpr int kDaTmaWarp = kDaEpiWarps;
constexpr int kDaMmaWarp = kDaTmaWarp + 1;
constexpr int kDaThreads = (kDaMmaWarp + 1) * 32;  // 576

template <typename T>
__global__ void __launch_bounds__(kDaThreads, 1)
sage_bwd_da_kernel(const __grid_constant__ CUtensorMap tmap_dz, const __grid_constant__ CUtensorMap tmap_dz_lo,
                   const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_w_lo,
                   const int32_t* __restrict__ nbr, const int32_t* __restrict__ cnt, const int64_t* __restrict__ n_dev,
                   int64_t n_cap, int fanout, int Np, int F, int CB, float* __restrict__ dH_below, int nst) {
  using G = Geo<T>;
  constexpr int kPlanes = TcT<T>::kPlanes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int64_t n = min(*n_dev, n_cap);
  const int ncb = 2 * F / CB;                      // column blocks of a row tile: the first F/CB are the self half
  const int64_t n_tiles = ((n + 127) / 128) * ncb;
  if ((int64_t)blockIdx.x >= n_tiles) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = Np / G::kBC;                      // K chunks (Np is a multiple of one 128-byte row)
  const uint32_t a_bytes = 128 * 128;              // dZ tile [128 rows x 128 B], K-major
  const uint32_t b_bytes = (uint32_t)CB * 128;     // CB/kBC boxes of [kBC k-rows x 128 B], MN-major
  const uint32_t kbox_bytes = G::kBC * 128;        // one Wt box: kBC reduce rows x 128 B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = kPlanes * (a_bytes + b_bytes);  // [A hi | (A lo) | B hi | (B lo)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  // per-warp transpose scratch of the epilogue: 32 rows x 16 columns fp32, row stride 20 floats (conflict-free 16-byte
  // stores by row, 64-byte segment loads by 8 rows)
  constexpr int kScrStride = 20;
  float* scratch = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6) + warp * (32 * kScrStride);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages);
  const uint32_t accf0 = smem_u32(bars + 2 * kMaxStages), acce0 = smem_u32(bars + 2 * kMaxStages + 2);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accf0 + 8 * a, 1);
      mbar_init(acce0 + 8 * a, kDaEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t acc_cols = CB <= 32 ? 32 : (CB <= 64 ? 64 : (CB <= 128 ? 128 : 256));
  const uint32_t tmem_cols = 2 * acc_cols;
  if (warp == kDaMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < kDaEpiWarps) {
    // ===================== epilogue: accumulator -> scatter into the gradient of the layer below =====================
    // A TMEM lane is a ROW of the tile, so a warp reading the accumulator holds 32 different rows: reductions issued
    // straight from those registers would touch 32 destination rows with 16 bytes each. Every [32 rows x 16 cols] piece
    // therefore takes one trip through a per-warp shared-memory scratch and leaves as 8 rows x 64 contiguous bytes per
    // warp instruction (lane -> row r0 + lane/4, 16-byte column group lane%4).
    const int quarter = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter of the block
    const int cols_per = CB / 4;                   // CB is a multiple of 64: 16 / 32 / 64 columns per warp
    const int sub = lane >> 2, c4 = lane & 3;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int64_t row0 = (t / ncb) * 128 + quarter * 32;
      const int cb = (int)(t % ncb);
      const bool nbr_half = cb * CB >= F;
      const int colbase = cb * CB - (nbr_half ? F : 0) + cq * cols_per;  // column inside dH_below
      const int acc = it & 1;
      mbar_wait(accf0 + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < cols_per; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + cq * cols_per + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __syncwarp();  // previous piece fully consumed
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<uint4*>(scratch + lane * kScrStride + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int col = colbase + c0 + c4 * 4;
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 8) {
          const int64_t row = row0 + r0 + sub;
          if (row >= n) continue;
          const float4 x = *reinterpret_cast<const float4*>(scratch + (r0 + sub) * kScrStride + c4 * 4);
          if (!nbr_half) {
            red_add_v4(dH_below + row * F + col, x.x, x.y, x.z, x.w);
          } else {
            const int cn = cnt[row];
            const float w = cn > 0 ? 1.f / (float)cn : 0.f;
            const int32_t* nb = nbr + row * fanout;
            for (int k = 0; k < cn; ++k)
              red_add_v4(dH_below + (int64_t)max(nb[k], 0) * F + col, x.x * w, x.y * w, x.z * w, x.w * w);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce0 + 8 * acc);
    }
  } else if (warp == kDaTmaWarp) {
    if (lane == 0) {
      uint32_t g = 0;
      const int nboxes = CB / G::kBC;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int row0 = (int)((t / ncb) * 128);
        const int cb = (int)(t % ncb);
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int stage = g % nst;
          mbar_wait(empty0 + 8 * stage, ((g / nst) & 1) ^ 1);
          mbar_arrive_expect_tx(full0 + 8 * stage, stage_bytes);
          const uint32_t sa = smem_u32(smem + stage * stage_bytes), sb = sa + kPlanes * a_bytes;
          tma_load_2d(sa, &tmap_dz, full0 + 8 * stage, kc * G::kBC, row0);
          if constexpr (kPlanes == 2) tma_load_2d(sa + a_bytes, &tmap_dz_lo, full0 + 8 * stage, kc * G::kBC, row0);
          for (int b = 0; b < nboxes; ++b) {
            tma_load_2d(sb + b * kbox_bytes, &tmap_w, full0 + 8 * stage, cb * CB + b * G::kBC, kc * G::kBC);
            if constexpr (kPlanes == 2)
              tma_load_2d(sb + b_bytes + b * kbox_bytes, &tmap_w_lo, full0 + 8 * stage, cb * CB + b * G::kBC, kc * G::kBC);
          }
        }
      }
    }
  } else {
    // ===================== MMA issuer: A K-major (dZ), B MN-major (Wt rows are the reduce index) =====================
    const uint32_t fmt = (uint32_t)TcT<T>::kFmt;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(CB >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    auto kdesc = [](uint32_t a) { return make_sw128_kmajor_desc(a); };
    auto mdesc = [kbox_bytes](uint32_t a) { return make_mnmajor_desc<T>(a, kbox_bytes); };
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
          const uint32_t a0 = smem_u32(smem + stage * stage_bytes), b0 = a0 + kPlanes * a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_step<T>(tmem_d, a0 + k * 32, a_bytes, kdesc, b0 + k * G::kMmaRowBytes, b_bytes, mdesc, idesc, (kc | k) != 0);
          umma_commit(empty0 + 8 * stage);
          if (kc == KC - 1) umma_commit(accf0 + 8 * acc);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (warp == kDaMmaWarp) {
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

template <typename T>
static void encode_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                      uint32_t box_inner, uint32_t box_outer, bool mn_major = false) {
  const cuuint64_t dim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t stride[1] = {(cuuint64_t)row_stride_elems * sizeof(T)};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  // MN-major 32-bit operands need the 32-byte-atom flavour of the 128-byte swizzle (tc_common.cuh)
  const CUtensorMapSwizzle sw = (mn_major && sizeof(T) == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult rc = get_encode()(m, TcT<T>::kTmap, 2, const_cast<void*>(base), dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
}

template <typename T>
static void launch_dw(const void* dH, const void* Hmask, const void* P, int ldp, int CP, float scale, const void* Q, int ldq,
                      int CQ, const void* n_dev, int64_t n_cap, void* out, int64_t ldi, int64_t ldj, void* gb, void* dz_out,
                      void* dz_out_lo, int nbs_hint, cudaStream_t st) {
  using G = Geo<T>;
  constexpr int kPlanes = TcT<T>::kPlanes;
  constexpr bool kQTMA = sizeof(T) == 2;
  DGLB_CHECK(CP % 128 == 0 && CP >= 128, "dW: the M-side width must be a multiple of 128");
  int NBS = nbs_hint > 0 ? nbs_hint : (CQ % 256 == 0 ? 256 : (CQ % 128 == 0 ? 128 : CQ));
  DGLB_CHECK(NBS <= 256 && NBS % G::kBC == 0 && CQ % NBS == 0 && (512 % (NBS / G::kEPP)) == 0 &&
                 (G::kKR * (NBS / G::kEPP)) % 512 == 0,
             "dW: unsupported N-side block");
  DGLB_CHECK(ldp % G::kEPP == 0 && ldq % G::kEPP == 0, "dW: row strides must keep 16-byte alignment");
  DGLB_CHECK(kPlanes == 1 || dz_out == nullptr || dz_out_lo != nullptr, "fp32 dZ write-out needs both planes");
  const int MB = CP / 128, NB = CQ / NBS;
  const size_t stage_bytes = kPlanes * ((size_t)G::kPBytes + (size_t)NBS * 128);
  const size_t fixed_dw = 1024 + 1024 + (size_t)kEpilogueWarps * 32 * 20 * 4;  // slack + barriers/bias + transpose scratch
  const int nst = (int)std::max<size_t>(1, std::min<size_t>(kMaxStages, ((size_t)(226 * 1024) - fixed_dw) / stage_bytes));
  const size_t smem = nst * stage_bytes + fixed_dw;
  const int64_t chunks = (n_cap + G::kKR - 1) / G::kKR;
  const int Qn = MB * NB;
  const int groups = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count() / Qn, chunks));
  CUtensorMap tq;
  if (kQTMA) encode_2d<T>(&tq, Q, (uint64_t)CQ, (uint64_t)n_cap, (uint64_t)ldq, G::kBC, G::kKR, true);
  else memset(&tq, 0, sizeof(tq));
  auto kern = sage_bwd_dw_kernel<T, kQTMA>;
  DGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)(groups * Qn), kThreads, smem, st>>>(tq, (const float*)dH, (const T*)Hmask, (const T*)P, ldp, scale,
                                                       (const T*)Q, ldq, (const int64_t*)n_dev, n_cap, MB, NB, NBS,
                                                       (float*)out, ldi, ldj, (float*)gb, (T*)dz_out, (float*)dz_out_lo, nst);
  DGLB_LAUNCH_CHECK();
}

template <typename T>
static void launch_da(const void* dz, const void* dz_lo, const void* Wt, const void* Wt_lo, const void* nbr, const void* cnt,
                      const void* n_dev, int64_t n_cap, int fanout, int Np, int F, void* dH_below, int cb_hint, cudaStream_t st) {
  using G = Geo<T>;
  constexpr int kPlanes = TcT<T>::kPlanes;
  DGLB_CHECK(F % 64 == 0 && F >= 64 && F <= 256, "dA: F must be a multiple of 64 in [64, 256]");
  DGLB_CHECK(Np % G::kBC == 0 && Np >= G::kBC, "dA: Np must be a multiple of one 128-byte row");
  DGLB_CHECK(kPlanes == 1 || (dz_lo && Wt_lo), "fp32 dA needs the lo planes");
  // column block of a tile: the scatter epilogue is the expensive part, so prefer enough (row tile x column block)
  // work items to occupy every SM a few times over; the dZ tile re-read per block comes from L2
  const int64_t row_tiles = (n_cap + 127) / 128;
  int CB = cb_hint;
  if (CB <= 0) {
    CB = F % 256 == 0 ? 256 : (F % 128 == 0 ? 128 : 64);
    while (CB > 64 && row_tiles * (2 * F / CB) < 3 * (int64_t)sm_count() && F % (CB / 2) == 0) CB /= 2;
  }
  DGLB_CHECK(CB % 64 == 0 && CB <= 256 && F % CB == 0, "dA: unsupported column block");
  CUtensorMap tdz, tdzl, tw, twl;
  encode_2d<T>(&tdz, dz, (uint64_t)Np, (uint64_t)n_cap, (uint64_t)Np, G::kBC, 128);
  encode_2d<T>(&tw, Wt, (uint64_t)(2 * F), (uint64_t)Np, (uint64_t)(2 * F), G::kBC, G::kBC, true);
  if (kPlanes == 2) {
    encode_2d<T>(&tdzl, dz_lo, (uint64_t)Np, (uint64_t)n_cap, (uint64_t)Np, G::kBC, 128);
    encode_2d<T>(&twl, Wt_lo, (uint64_t)(2 * F), (uint64_t)Np, (uint64_t)(2 * F), G::kBC, G::kBC, true);
  } else {
    tdzl = tdz;
    twl = tw;
  }
  const size_t stage_bytes = kPlanes * ((size_t)128 * 128 + (size_t)CB * 128);
  const size_t fixed = 1024 + 256 + (size_t)kDaEpiWarps * 32 * 20 * 4;  // alignment slack + barriers + transpose scratch
  const int nst = (int)std::max<size_t>(1, std::min<size_t>(kMaxStages, ((size_t)(226 * 1024) - fixed) / stage_bytes));
  const size_t smem = nst * stage_bytes + fixed;
  const int64_t tiles = row_tiles * (2 * F / CB);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, sm_count());
  auto kern = sage_bwd_da_kernel<T>;
  DGLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kDaThreads, smem, st>>>(tdz, tdzl, tw, twl, (const int32_t*)nbr, (const int32_t*)cnt, (const int64_t*)n_dev,
                                      n_cap, fanout, Np, F, CB, (float*)dH_below, nst);
  DGLB_LAUNCH_CHECK();
}

}  // namespace sbw
}  // namespace dglb

using namespace dglb;

// Weight gradient of one fused layer:  out[i * ldi + j * ldj] += sum_r P[r, i] * Q[r, j],  r < *n_dev.
//   dH != null : P = dZ built on the fly, dZ[r, i] = Hmask[r, i] > 0 ? dH[r, i] * scale : 0 (dH fp32, Hmask in T, both
//                [n_cap, ldp]); gb (optional, fp32 [CP]) receives the column sums; dz_out (optional, [n_cap, ldp] in T;
//                fp32: hi plane + dz_out_lo) receives dZ for the dA kernel.
//   dH == null : P is given ([n_cap, ldp] in T).
//   Q [n_cap, ldq] in T, CQ live columns. CP % 128 == 0. nbs: N-side block (0 = auto).
DGLB_API int dglb200_engine_sage_bwd_dw(int dtype, const void* dH, const void* Hmask, const void* P, int ldp, int CP,
                                        float scale, const void* Q, int ldq, int CQ, const void* n_dev, int64_t n_cap,
                                        void* out, int64_t ldi, int64_t ldj, void* gb, void* dz_out, void* dz_out_lo,
                                        int nbs, void* stream) {
  DGLB_TRY
  DGLB_CHECK(dtype == kBF16 || dtype == kF32, "sage_bwd_dw: bf16 or fp32");
  DGLB_CHECK((dH != nullptr) != (P != nullptr), "sage_bwd_dw: pass either (dH, Hmask) or P");
  if (n_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == kBF16)
    sbw::launch_dw<__nv_bfloat16>(dH, Hmask, P, ldp, CP, scale, Q, ldq, CQ, n_dev, n_cap, out, ldi, ldj, gb, dz_out, dz_out_lo, nbs, st);
  else
    sbw::launch_dw<float>(dH, Hmask, P, ldp, CP, scale, Q, ldq, CQ, n_dev, n_cap, out, ldi, ldj, gb, dz_out, dz_out_lo, nbs, st);
  DGLB_CATCH
}

// Input gradient of one fused layer, scattered:  dA = dz [n_cap, Np] @ Wt [Np, 2F];
//   dH_below[d] += dA[d, :F];  dH_below[nbr[d, k]] += dA[d, F:] / cnt[d]  (k < cnt[d]);  dH_below fp32 [*, F], pre-zeroed.
//   cb: column block of a work item (64 | 128 | 256, 0 = chosen from the tile count).
DGLB_API int dglb200_engine_sage_bwd_da(int dtype, const void* dz, const void* dz_lo, const void* Wt, const void* Wt_lo,
                                        const void* nbr, const void* cnt, const void* n_dev, int64_t n_cap, int fanout,
                                        int Np, int F, void* dH_below, int cb, void* stream) {
  DGLB_TRY
  DGLB_CHECK(dtype == kBF16 || dtype == kF32, "sage_bwd_da: bf16 or fp32");
  if (n_cap == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == kBF16) sbw::launch_da<__nv_bfloat16>(dz, dz_lo, Wt, Wt_lo, nbr, cnt, n_dev, n_cap, fanout, Np, F, dH_below, cb, st);
  else sbw::launch_da<float>(dz, dz_lo, Wt, Wt_lo, nbr, cnt, n_dev, n_cap, fanout, Np, F, dH_below, cb, st);
  DGLB_CATCH
}
