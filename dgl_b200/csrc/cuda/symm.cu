// Symmetric heap over the CUDA virtual-memory-management API + NVLink SHARP (NVLS) multicast.
//
// Every rank of one NVSwitch domain creates a physical allocation (cuMemCreate), exports it as a POSIX file descriptor
// and maps ALL ranks' allocations into one contiguous virtual range: peer r's copy lives at base + r * stride on every
// rank (a symmetric heap: no per-buffer pointer tables, addresses are computed). On top of that a multicast object
// binds the ranks' allocations together: a store through the multicast mapping lands in every rank's copy
// (multimem.st), a multimem.ld_reduce returns the SUM over all copies, reduced inside the NVSwitch. The file
// descriptors travel between the processes in Python (SCM_RIGHTS over a unix socket, parallel/symm.py).
//
// Kernels in this file:
//   nvls_barrier_kernel        one multimem.red per rank + local spinning: a barrier that costs one NVLink round trip
//   nvls_allreduce_adam_kernel two-shot all-reduce fused with Adam: rank r pulls the reduced gradient of ITS slice with
//                              multimem.ld_reduce, applies Adam to the slice, and multicasts the new fp32 parameters and
//                              their tensor-core operand copies (bf16 shadow or TF32 hi/lo planes) to every rank with
//                              multimem.st - replicas stay bit-identical because each element has exactly one writer.
// Replaces DDP's NCCL all-reduce + optimizer step; the CUDA-IPC/ld.global path in comm.cu /
// engine.cu remains the fallback where multicast is not available.
#include <cuda.h>

#include "common.cuh"

namespace dglb {
namespace symm {

struct Driver {
  decltype(&cuMemCreate) MemCreate;
  decltype(&cuMemRelease) MemRelease;
  decltype(&cuMemExportToShareableHandle) MemExport;
  decltype(&cuMemImportFromShareableHandle) MemImport;
  decltype(&cuMemAddressReserve) AddressReserve;
  decltype(&cuMemAddressFree) AddressFree;
  decltype(&cuMemMap) MemMap;
  decltype(&cuMemUnmap) MemUnmap;
  decltype(&cuMemSetAccess) SetAccess;
  decltype(&cuMemGetAllocationGranularity) GetGranularity;
  decltype(&cuMulticastCreate) McCreate;
  decltype(&cuMulticastAddDevice) McAddDevice;
  decltype(&cuMulticastBindMem) McBindMem;
  decltype(&cuMulticastGetGranularity) McGetGranularity;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute;
  decltype(&cuDeviceGet) DeviceGet;
  decltype(&cuGetErrorString) GetErrorString;
};

template <typename F>
static void load(F& fn, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
    throw std::runtime_error(std::string("driver entry point not available: ") + name);
  fn = reinterpret_cast<F>(p);
}

static const Driver& drv() {
  static Driver d;
  static bool ready = false;
  if (!ready) {
    cudaFree(0);  // make sure the primary context exists on this thread
    load(d.MemCreate, "cuMemCreate");
    load(d.MemRelease, "cuMemRelease");
    load(d.MemExport, "cuMemExportToShareableHandle");
    load(d.MemImport, "cuMemImportFromShareableHandle");
    load(d.AddressReserve, "cuMemAddressReserve");
    load(d.AddressFree, "cuMemAddressFree");
    load(d.MemMap, "cuMemMap");
    load(d.MemUnmap, "cuMemUnmap");
    load(d.SetAccess, "cuMemSetAccess");
    load(d.GetGranularity, "cuMemGetAllocationGranularity");
    load(d.McCreate, "cuMulticastCreate");
    load(d.McAddDevice, "cuMulticastAddDevice");
    load(d.McBindMem, "cuMulticastBindMem");
    load(d.McGetGranularity, "cuMulticastGetGranularity");
    load(d.DeviceGetAttribute, "cuDeviceGetAttribute");
    load(d.DeviceGet, "cuDeviceGet");
    load(d.GetErrorString, "cuGetErrorString");
    ready = true;
  }
  return d;
}

static void cu_check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* msg = nullptr;
  drv().GetErrorString(r, &msg);
  throw std::runtime_error(std::string(what) + " failed: " + (msg ? msg : "unknown") + " (" + std::to_string((int)r) + ")");
}
#define CU(expr) cu_check(drv().expr, #expr)

static CUmemAllocationProp alloc_prop(int dev) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

static void set_rw(int dev, CUdeviceptr p, size_t size) {
  CUmemAccessDesc a = {};
  a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a.location.id = dev;
  a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CU(SetAccess(p, size, &a, 1));
}

// ------------------------------------------------------------------------------------------------------------ kernels
__device__ __forceinline__ float4 mm_ld_reduce_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_v4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mm_st_b64(void* mc, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y))
               : "memory");
}

// One arrival per rank through the multicast mapping increments the counter in EVERY rank's copy; each rank then spins
// on its own copy until `world` arrivals of this epoch are in. Counters only grow: target = epoch * world.
__global__ void nvls_barrier_kernel(unsigned int* mc_counter, unsigned int* local_counter, unsigned long long* epoch_dev,
                                    int world, int* error_flag, long long timeout_cycles) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long epoch = *epoch_dev + 1;
  *epoch_dev = epoch;
  __threadfence_system();  // everything this rank wrote before the barrier is visible before its arrival
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_counter), "r"(1u) : "memory");
  const unsigned int target = (unsigned int)(epoch * (unsigned long long)world);
  const long long t0 = clock64();
  unsigned int seen;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local_counter) : "memory");
    if (clock64() - t0 > timeout_cycles) {
      if (error_flag) *error_flag = 1;
      break;
    }
  } while ((int)(seen - target) < 0);
}

// grad_mc / param_mc / shadow*_mc: multicast mappings; m, v, param_local: this rank's private / local copies.
// Rank r owns elements [lo, hi).
__global__ void __launch_bounds__(256) nvls_allreduce_adam_kernel(const float* __restrict__ grad_mc, float* __restrict__ param_mc,
                                                                  const float* __restrict__ param_local, float* __restrict__ m,
                                                                  float* __restrict__ v, __nv_bfloat16* __restrict__ shadow_mc,
                                                                  float* __restrict__ hi_mc, float* __restrict__ lo_mc,
                                                                  int64_t lo, int64_t hi, int world, float lr, float beta1,
                                                                  float beta2, float eps, float weight_decay,
                                                                  const int64_t* __restrict__ step_dev) {
  const int64_t t = *step_dev + 1;
  const float bc1 = 1.f - __powf(beta1, (float)t);
  const float bc2 = 1.f - __powf(beta2, (float)t);
  const float inv_world = 1.f / (float)world;
  for (int64_t i = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 g = mm_ld_reduce_v4(grad_mc + i);  // sum over all ranks, reduced in the switch
    float4 p4 = *reinterpret_cast<const float4*>(param_local + i);
    float4 m4 = *reinterpret_cast<float4*>(m + i);
    float4 v4 = *reinterpret_cast<float4*>(v + i);
    float* gp = reinterpret_cast<float*>(&g);
    float* pp = reinterpret_cast<float*>(&p4);
    float* mp = reinterpret_cast<float*>(&m4);
    float* vp = reinterpret_cast<float*>(&v4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = gp[k] * inv_world + weight_decay * pp[k];
      mp[k] = beta1 * mp[k] + (1.f - beta1) * gg;
      vp[k] = beta2 * vp[k] + (1.f - beta2) * gg * gg;
      pp[k] -= lr * (mp[k] / bc1) / (sqrtf(vp[k] / bc2) + eps);
    }
    *reinterpret_cast<float4*>(m + i) = m4;
    *reinterpret_cast<float4*>(v + i) = v4;
    mm_st_v4(param_mc + i, p4);  // every replica receives the owner's value: bit-identical parameters everywhere
    if (shadow_mc) {
      __nv_bfloat162 a = __floats2bfloat162_rn(pp[0], pp[1]), b = __floats2bfloat162_rn(pp[2], pp[3]);
      mm_st_b64(shadow_mc + i, make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b)));
    }
    if (hi_mc) {
      float4 h4, l4;
      float* hp = reinterpret_cast<float*>(&h4);
      float* lp = reinterpret_cast<float*>(&l4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(pp[k]));
        hp[k] = __uint_as_float(h & 0xFFFFE000u);
        lp[k] = pp[k] - hp[k];
      }
      mm_st_v4(hi_mc + i, h4);
      mm_st_v4(lo_mc + i, l4);
    }
  }
}

__global__ void bump_kernel(int64_t* step_dev) { *step_dev += 1; }

}  // namespace symm
}  // namespace dglb

using namespace dglb;
using namespace dglb::symm;

// vmm_ok: VMM with POSIX-fd export; mc_ok: multicast objects; gran: allocation granularity to round sizes to.
DGLB_API int dglb200_symm_query(int dev, int* vmm_ok, int* mc_ok, int64_t* gran) {
  DGLB_TRY
  *vmm_ok = *mc_ok = 0;
  *gran = 2 << 20;
  CUdevice d;
  CU(DeviceGet(&d, dev));
  int a = 0;
  CU(DeviceGetAttribute(&a, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d));
  *vmm_ok = a;
  a = 0;
  CU(DeviceGetAttribute(&a, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d));
  *mc_ok = a;
  CUmemAllocationProp p = alloc_prop(dev);
  size_t g = 0;
  CU(GetGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (*mc_ok) {
    CUmulticastObjectProp mp = {};
    mp.numDevices = 2;
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (drv().McGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
  }
  *gran = (int64_t)g;
  DGLB_CATCH
}

// physical allocation of `bytes` (a multiple of the granularity) on `dev`; returns the handle and an exportable fd
DGLB_API int dglb200_symm_create(int dev, int64_t bytes, uint64_t* handle, int* fd) {
  DGLB_TRY
  CUmemAllocationProp p = alloc_prop(dev);
  CUmemGenericAllocationHandle h;
  CU(MemCreate(&h, (size_t)bytes, &p, 0));
  int f = -1;
  CU(MemExport(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *handle = (uint64_t)h;
  *fd = f;
  DGLB_CATCH
}

DGLB_API int dglb200_symm_import(int fd, uint64_t* handle) {
  DGLB_TRY
  CUmemGenericAllocationHandle h;
  CU(MemImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  *handle = (uint64_t)h;
  DGLB_CATCH
}

// map `world` handles (index = rank) back to back: rank r at base + r * bytes
DGLB_API int dglb200_symm_map(int dev, const uint64_t* handles, int world, int64_t bytes, int64_t gran, void** base) {
  DGLB_TRY
  CUdeviceptr va = 0;
  CU(AddressReserve(&va, (size_t)bytes * world, (size_t)gran, 0, 0));
  for (int r = 0; r < world; ++r) CU(MemMap(va + (size_t)r * bytes, (size_t)bytes, 0, (CUmemGenericAllocationHandle)handles[r], 0));
  set_rw(dev, va, (size_t)bytes * world);
  *base = (void*)va;
  DGLB_CATCH
}

DGLB_API int dglb200_symm_mc_create(int world, int64_t bytes, uint64_t* mc, int* fd) {
  DGLB_TRY
  CUmulticastObjectProp mp = {};
  mp.numDevices = (unsigned)world;
  mp.size = (size_t)bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  CU(McCreate(&h, &mp));
  int f = -1;
  CU(MemExport(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *mc = (uint64_t)h;
  *fd = f;
  DGLB_CATCH
}

DGLB_API int dglb200_symm_mc_add_device(uint64_t mc, int dev) {
  DGLB_TRY
  CUdevice d;
  CU(DeviceGet(&d, dev));
  CU(McAddDevice((CUmemGenericAllocationHandle)mc, d));
  DGLB_CATCH
}

// every device must have been added (by every rank) before any rank binds
DGLB_API int dglb200_symm_mc_bind_map(int dev, uint64_t mc, uint64_t local_handle, int64_t bytes, int64_t gran, void** mc_ptr) {
  DGLB_TRY
  CU(McBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)local_handle, 0, (size_t)bytes, 0));
  CUdeviceptr va = 0;
  CU(AddressReserve(&va, (size_t)bytes, (size_t)gran, 0, 0));
  CU(MemMap(va, (size_t)bytes, 0, (CUmemGenericAllocationHandle)mc, 0));
  set_rw(dev, va, (size_t)bytes);
  *mc_ptr = (void*)va;
  DGLB_CATCH
}

DGLB_API int dglb200_symm_unmap(void* base, int64_t bytes) {
  DGLB_TRY
  if (base) {
    CU(MemUnmap((CUdeviceptr)base, (size_t)bytes));
    CU(AddressFree((CUdeviceptr)base, (size_t)bytes));
  }
  DGLB_CATCH
}

DGLB_API int dglb200_symm_release(uint64_t handle) {
  DGLB_TRY
  CU(MemRelease((CUmemGenericAllocationHandle)handle));
  DGLB_CATCH
}

DGLB_API int dglb200_symm_barrier(void* mc_counter, void* local_counter, void* epoch_dev, int world, void* error_flag,
                                  double timeout_s, void* stream) {
  DGLB_TRY
  const long long cycles = (long long)(timeout_s * 1.9e9);
  nvls_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned int*)mc_counter, (unsigned int*)local_counter,
                                                          (unsigned long long*)epoch_dev, world, (int*)error_flag, cycles);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// Two-shot all-reduce + Adam over the multicast mappings. numel % (4 * world) need not hold: slices are 4-aligned.
DGLB_API int dglb200_symm_allreduce_adam(const void* grad_mc, void* param_mc, const void* param_local, void* m, void* v,
                                         void* shadow_mc, void* hi_mc, void* lo_mc, int64_t numel, int world, int rank,
                                         float lr, float beta1, float beta2, float eps, float weight_decay, void* step_dev,
                                         void* stream) {
  DGLB_TRY
  DGLB_CHECK(numel % 4 == 0, "flat parameter buffer must be padded to a multiple of 4 elements");
  int64_t slice = (numel + world - 1) / world;
  slice = (slice + 3) / 4 * 4;
  const int64_t lo = std::min<int64_t>(numel, (int64_t)rank * slice), hi = std::min<int64_t>(numel, lo + slice);
  cudaStream_t st = (cudaStream_t)stream;
  if (hi > lo) {
    int64_t blocks = ((hi - lo) / 4 + 255) / 256;
    blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, 148 * 2));
    nvls_allreduce_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float*)grad_mc, (float*)param_mc, (const float*)param_local,
                                                                (float*)m, (float*)v, (__nv_bfloat16*)shadow_mc, (float*)hi_mc,
                                                                (float*)lo_mc, lo, hi, world, lr, beta1, beta2, eps, weight_decay,
                                                                (const int64_t*)step_dev);
  }
  bump_kernel<<<1, 1, 0, st>>>((int64_t*)step_dev);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
