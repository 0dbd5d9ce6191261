// Shared sm_100a tensor-core plumbing: mbarrier / TMA / tcgen05 (UMMA, TMEM) PTX wrappers, SWIZZLE_128B smem
// descriptors (K-major and MN-major) and the driver entry point for tensor-map encoding.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dglb {

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
  // try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase flips instead of
  // burning issue slots that the gather / epilogue warps of the same SM sub-partition need
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
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

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
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

// MN-major SWIZZLE_128B operand: the tile is a row of [k-rows x 128 B] boxes (64 MN elements each, one k index per
// 128-byte row, 16-byte chunks XOR-swizzled by (row & 7) - the image TMA / sw128_off() produce). LBO = byte stride
// between 64-element boxes along MN, SBO = 1024 B between groups of 8 k-rows.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t box_stride_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((box_stride_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset
  d |= (uint64_t)(1024 >> 4) << 32;                         // stride byte offset
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN-major operands of 32-bit element types (kind::tf32) have ONE legal swizzled layout: SWIZZLE_128B with a 32-byte
// base ("128B_BASE32B", layout type 1): the 32-byte chunks of a 128-byte row are XOR-swizzled by (k-row & 3), an atom is
// 4 k-rows x 128 B (512 B) and atoms repeat along K with SBO = 512 B. TMA writes the same image with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B. (16-byte-base SWIZZLE_128B is only legal for K-major tf32.)
__device__ __forceinline__ uint64_t make_sw128b32_mnmajor_desc(uint32_t smem_addr, uint32_t box_stride_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((box_stride_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset: next 128-byte-wide box along MN
  d |= (uint64_t)(512 >> 4) << 32;                          // stride byte offset: next 4-row atom along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                   // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of (k-row r, 16-byte chunk c) inside a [rows x 128 B] box of that layout
__device__ __forceinline__ uint32_t sw128b32_off(int r, int c) {
  return (uint32_t)(r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
}

// swizzled byte offset of (row r, 16-byte chunk c) inside a [rows x 128 B] K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// Operand-type traits shared by the engine kernels (sage_tc.cu, sage_bwd_tc.cu).
//   bf16 : kind::f16, one operand plane, 64 elements per 128-byte swizzle row, K = 16 per instruction.
//   fp32 : kind::tf32 on an error-compensated split x = hi + lo (hi = x rounded to TF32, lo = x - hi, exact in fp32).
//          Every product is issued as lo*hi + hi*lo + hi*hi into the same fp32 TMEM accumulator, which restores
//          fp32-level accuracy (relative error ~2^-21 per product, below the fp32 accumulation noise of K >= 64) on the
//          tensor cores - the "3xTF32" scheme. Two operand planes, 32 elements per 128-byte row, K = 8 per instruction.
template <typename T>
struct TcT;
template <>
struct TcT<__nv_bfloat16> {
  static constexpr int kPlanes = 1, kRowElems = 64, kMmaK = 16, kFmt = 1, kProducts = 1;
  static constexpr CUtensorMapDataType kTmap = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};
template <>
struct TcT<float> {
  static constexpr int kPlanes = 2, kRowElems = 32, kMmaK = 8, kFmt = 2, kProducts = 3;
  static constexpr CUtensorMapDataType kTmap = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// One logical MMA step on operand planes laid out `a_plane` / `b_plane` bytes apart (fp32: plane 0 = hi, plane 1 = lo).
// DescFn turns a shared-memory byte address into the (K-major or MN-major) matrix descriptor.
template <typename T, typename DA, typename DB>
__device__ __forceinline__ void umma_step(uint32_t tmem_d, uint32_t a_addr, uint32_t a_plane, DA da, uint32_t b_addr,
                                          uint32_t b_plane, DB db, uint32_t idesc, bool accumulate) {
  if constexpr (TcT<T>::kPlanes == 1) {
    umma_bf16(tmem_d, da(a_addr), db(b_addr), idesc, accumulate ? 1u : 0u);
  } else {
    // small terms first: lo*hi, hi*lo, then hi*hi
    umma_tf32(tmem_d, da(a_addr + a_plane), db(b_addr), idesc, accumulate ? 1u : 0u);
    umma_tf32(tmem_d, da(a_addr), db(b_addr + b_plane), idesc, 1u);
    umma_tf32(tmem_d, da(a_addr), db(b_addr), idesc, 1u);
  }
}

// MN-major operand helpers selected by the element type (see make_sw128b32_mnmajor_desc for the 32-bit case)
template <typename T>
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t box_stride_bytes) {
  if constexpr (sizeof(T) == 4) return make_sw128b32_mnmajor_desc(smem_addr, box_stride_bytes);
  else return make_sw128_mnmajor_desc(smem_addr, box_stride_bytes);
}
template <typename T>
__device__ __forceinline__ uint32_t mn_box_off(int r, int c) {
  if constexpr (sizeof(T) == 4) return sw128b32_off(r, c);
  else return sw128_off(r, c);
}

// x = hi + lo with hi exactly representable in TF32 (round-to-nearest on the 10-bit mantissa)
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h & 0xFFFFE000u);
  lo = x - hi;
}
__device__ __forceinline__ void tf32_split4(const float4& x, float4& hi, float4& lo) {
  tf32_split(x.x, hi.x, lo.x); tf32_split(x.y, hi.y, lo.y); tf32_split(x.z, hi.z, lo.z); tf32_split(x.w, hi.w, lo.w);
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// row-granular bulk copy global -> shared (the TMA engine without a tensor map; SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor-map encoding is a DRIVER call: it needs the primary context current on the calling thread. PyTorch's autograd
// worker threads have not necessarily touched this library's runtime yet, so bind the context first (cudaFree(0)).
static inline EncodeTiledFn get_encode() {
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(0);
    ctx_bound = true;
  }
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
