// Peer-memory (NVLink 5 / NVSwitch) communication kernels for one-process-per-GPU jobs.
//
// Symmetric buffers are plain cudaMalloc allocations exported with CUDA IPC and mapped into every
// rank, so any SM can load/store any peer's shard directly (no NCCL call, no proxy thread, no
// staging copy). On top of that:
//   * device-side barrier (release/acquire flags over NVLink) — graph capturable, no host round trip
//   * one-shot all-reduce (each rank pulls all peers' buffers with 16B P2P loads, sums, scales,
//     optional bf16 cast) — the latency-optimal choice for the < 1 MB gradient sets of GNNs
//   * halo-fused g-SpMM: aggregation whose gather resolves a global source id to (owner rank,
//     local row) and loads the row straight from the owner's HBM over NVLink into registers, so the
//     halo all-gather that a partitioned full-graph layer needs is never materialised
//   * p2p row gather / scatter-add for sharded embedding tables (sparse pull / push)
// Reference behaviour being replaced: python/dgl/cuda/nccl.py:7-189 (3x all_to_all_single + a
// blocking stream sync), graphbolt cooperative_conv.py:27-93, DDP all-reduce.
#include "common.cuh"

namespace dglb {

constexpr int kMaxPeers = 8;

struct Peers {
  void* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint64_t* addr, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* addr) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// flags[r] lives on every rank: flags_of(peer)[my_rank] = epoch announces arrival to `peer`.
__global__ void barrier_kernel(Peers flags, int rank, int world, uint64_t* epoch_local, int* error_flag,
                               long long timeout_cycles) {
  __shared__ uint64_t e_sh;
  if (threadIdx.x == 0) e_sh = *epoch_local + 1;
  __syncthreads();
  const uint64_t e = e_sh;
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint64_t*>(flags.p[t]) + rank, e);
    const uint64_t* mine = reinterpret_cast<const uint64_t*>(flags.p[rank]) + t;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < e) {
      if (clock64() - t0 > timeout_cycles) {
        if (error_flag) atomicExch(error_flag, 1);
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_local = e;
}

// out = scale * sum_r peers[r][i]; optional bf16 copy.
__global__ void __launch_bounds__(256) allreduce_oneshot_kernel(Peers bufs, int world, float* __restrict__ out,
                                                                __nv_bfloat16* __restrict__ out_bf16, int64_t numel,
                                                                float scale) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < numel;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int r = 0; r < world; ++r) {
      const float4 x = ld_volatile_f4(reinterpret_cast<const float*>(bufs.p[r]) + i);
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    *reinterpret_cast<float4*>(out + i) = acc;
    if (out_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(out_bf16 + i) = pk;
    }
  }
}

// Halo-fused aggregation on a range-partitioned feature table:
//   out[v] = (mean|sum)_{u in N(v)} X_global[u],  X_global[u] = shards[u / shard_rows][u % shard_rows]
// rows = local destination rows (CSC of the local partition with GLOBAL source ids).
template <typename T, typename I, int VB>
__global__ void __launch_bounds__(256) halo_spmm_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                        int64_t num_rows, Peers shards, int64_t shard_rows, int F,
                                                        T* __restrict__ out, int mean, int sub,
                                                        const float* __restrict__ edge_w) {
  constexpr int VE = VB / sizeof(T);
  using V = typename std::conditional<VB == 16, uint4, typename std::conditional<VB == 8, uint2, uint32_t>::type>::type;
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int nvec = F / VE;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; row < num_rows;
       row += (int64_t)gridDim.x * rows_per_block) {
    const int64_t s = indptr[row], t = indptr[row + 1];
    for (int c = lane; c < nvec; c += sub) {
      float acc[VE];
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[i] = 0.f;
      int64_t p = s;
      for (; p + 4 <= t; p += 4) {  // 4 (possibly remote) rows in flight: NVLink latency ~2 us needs MLP
        V v[4];
        float w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t u = indices[p + k];
          const int64_t owner = u / shard_rows;
          const int64_t lr = u - owner * shard_rows;
          v[k] = *(reinterpret_cast<const V*>(reinterpret_cast<const T*>(shards.p[owner]) + lr * F) + c);
          w[k] = edge_w ? edge_w[p + k] : 1.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const T* q = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
          for (int i = 0; i < VE; ++i) acc[i] = fmaf(to_f<T>(q[i]), w[k], acc[i]);
        }
      }
      for (; p < t; ++p) {
        const int64_t u = indices[p];
        const int64_t owner = u / shard_rows;
        const int64_t lr = u - owner * shard_rows;
        V v = *(reinterpret_cast<const V*>(reinterpret_cast<const T*>(shards.p[owner]) + lr * F) + c);
        const float w = edge_w ? edge_w[p] : 1.f;
        const T* q = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] = fmaf(to_f<T>(q[i]), w, acc[i]);
      }
      const float sc = mean ? (t > s ? 1.f / (float)(t - s) : 0.f) : 1.f;
      V o;
      T* po = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int i = 0; i < VE; ++i) po[i] = from_f<T>(acc[i] * sc);
      reinterpret_cast<V*>(out + row * F)[c] = o;
    }
  }
}

// Backward of the halo-fused aggregation: dX_global[u] += w * dOut[v] pushed to the owner with
// vector red.add over NVLink (dX shards fp32).
template <typename T, typename I>
__global__ void __launch_bounds__(256) halo_spmm_bwd_kernel(const I* __restrict__ indptr,
                                                            const I* __restrict__ indices, int64_t num_rows,
                                                            Peers dshards, int64_t shard_rows, int F,
                                                            const T* __restrict__ dout, int mean, int sub) {
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; row < num_rows;
       row += (int64_t)gridDim.x * rows_per_block) {
    const int64_t s = indptr[row], t = indptr[row + 1];
    const float sc = mean ? (t > s ? 1.f / (float)(t - s) : 0.f) : 1.f;
    for (int f = lane * 4; f < F; f += sub * 4) {
      const float g0 = to_f<T>(dout[row * F + f]) * sc, g1 = to_f<T>(dout[row * F + f + 1]) * sc,
                  g2 = to_f<T>(dout[row * F + f + 2]) * sc, g3 = to_f<T>(dout[row * F + f + 3]) * sc;
      for (int64_t p = s; p < t; ++p) {
        const int64_t u = indices[p];
        const int64_t owner = u / shard_rows;
        float* dst = reinterpret_cast<float*>(dshards.p[owner]) + (u - owner * shard_rows) * F + f;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g0), "f"(g1), "f"(g2), "f"(g3) : "memory");
      }
    }
  }
}

// sharded-table row gather: out[i] = shards[idx[i] / shard_rows][idx[i] % shard_rows]
template <typename I>
__global__ void __launch_bounds__(256) p2p_gather_kernel(Peers shards, int64_t shard_rows, const I* __restrict__ idx,
                                                         int64_t n, char* __restrict__ out, int64_t row_bytes, int sub) {
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int64_t nvec = row_bytes / 16;
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; r < n;
       r += (int64_t)gridDim.x * rows_per_block) {
    const int64_t u = idx[r];
    const int64_t owner = u / shard_rows;
    const uint4* sp = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(shards.p[owner]) + (u - owner * shard_rows) * row_bytes);
    uint4* dp = reinterpret_cast<uint4*>(out + r * row_bytes);
    for (int64_t c = lane; c < nvec; c += sub) dp[c] = sp[c];
  }
}

// sharded-table scatter-add (sparse gradient push): shards[owner][row] += vals[i]   (fp32)
template <typename I>
__global__ void __launch_bounds__(256) p2p_scatter_add_kernel(Peers shards, int64_t shard_rows,
                                                              const I* __restrict__ idx, int64_t n,
                                                              const float* __restrict__ vals, int F, int sub) {
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; r < n;
       r += (int64_t)gridDim.x * rows_per_block) {
    const int64_t u = idx[r];
    const int64_t owner = u / shard_rows;
    float* dst = reinterpret_cast<float*>(shards.p[owner]) + (u - owner * shard_rows) * F;
    for (int f = lane * 4; f < F; f += sub * 4) {
      const float4 v = *reinterpret_cast<const float4*>(vals + r * F + f);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + f), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Pipelined halo exchange + aggregation in ONE cooperative kernel (static plan, see parallel/p2p.py HaloPlan).
//   * The plan deduplicated the remote source ids: every halo row crosses NVLink exactly once. The first
//     `puller_warps` warps of EVERY CTA pull chunks of `chunk_rows` halo rows (taken from an atomic chunk counter) with
//     16 B peer loads, 8 in flight per lane, and write them behind the local rows of `table`. Remote-load throughput is
//     bounded per SM by its outstanding-request budget, so the pull is spread over all SMs instead of a few CTAs.
//   * All other warps (and the pullers once the chunks are gone) take row groups from an atomic work counter. The row
//     order is INTERIOR rows (all sources local) first, then BOUNDARY rows: interior aggregation overlaps the transfer,
//     and a warp reaching a boundary group waits on a shared-memory flag that ONE warp per CTA raises after it has seen
//     the device-wide completed-chunk counter reach n_chunks (one global poller per CTA, nobody scans flags).
// The halo all-gather of the reference (NCCL all_gather / KVStore pull before the layer) never exists as a separate
// phase. Cooperative launch guarantees that all pullers are resident, so the wait cannot deadlock; a cycle-count
// timeout raises the communicator's error flag instead of hanging the GPU.
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* addr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

template <typename T, typename I, int VB>
__global__ void __launch_bounds__(256) halo_pipe_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                        const int64_t* __restrict__ order, int64_t n_order,
                                                        int64_t n_interior, T* table, Peers shards, int64_t shard_rows,
                                                        const I* __restrict__ halo_ids, int64_t n_halo, int chunk_rows,
                                                        int puller_warps, unsigned* counters, int F,
                                                        T* __restrict__ out, int mean, int sub, int group_rows,
                                                        int* error_flag) {
  constexpr int VE = VB / sizeof(T);
  using V = typename std::conditional<VB == 16, uint4, typename std::conditional<VB == 8, uint2, uint32_t>::type>::type;
  const int nvec = F / VE;
  const int n_chunks = (int)((n_halo + chunk_rows - 1) / chunk_rows);
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
  __shared__ volatile int halo_ready;
  if (threadIdx.x == 0) halo_ready = n_chunks == 0 ? 1 : 0;
  __syncthreads();
  // ------------------------------------------------ phase A (puller warps): pull chunks over NVLink
  if (wid < puller_warps && n_chunks > 0) {
    V* dst_base = reinterpret_cast<V*>(table + shard_rows * (int64_t)F);
    for (;;) {
      unsigned c = 0;
      if (ln == 0) c = atomicAdd(counters + 1, 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      if ((int)c >= n_chunks) break;
      const int64_t h0 = (int64_t)c * chunk_rows;
      const int hn = (int)((n_halo - h0) < chunk_rows ? (n_halo - h0) : chunk_rows);
      const int items = hn * nvec;
      constexpr int U = 8;
      for (int base = ln; base < items; base += 32 * U) {
        V v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int it = base + k * 32;
          if (it < items) {
            const int h = it / nvec;
            const int vc = it - h * nvec;
            const int64_t u = halo_ids[h0 + h];
            const int64_t owner = u / shard_rows;
            const int64_t lr = u - owner * shard_rows;
            v[k] = __ldcg(reinterpret_cast<const V*>(reinterpret_cast<const T*>(shards.p[owner]) + lr * F) + vc);
          }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int it = base + k * 32;
          if (it < items) dst_base[h0 * nvec + it] = v[k];
        }
      }
      __syncwarp();
      if (ln == 0) {
        __threadfence();
        atomicAdd(counters + 2, 1u);  // completed chunks
      }
    }
    if (wid == 0) {  // this CTA's poller: wait for the device-wide completion, then release the CTA's boundary work
      if (ln == 0) {
        const long long t0 = clock64();
        while ((int)ld_acquire_gpu_u32(counters + 2) < n_chunks) {
          if (clock64() - t0 > 4000000000ll) {
            if (error_flag) atomicExch(error_flag, 2);
            break;
          }
          __nanosleep(256);
        }
        __threadfence();
        halo_ready = 1;
      }
      __syncwarp();
    }
  }
  // ------------------------------------------------ phase B (all warps): aggregate row groups, warp-granular
  const int rows_per_warp = 32 / sub;
  const int lane = ln % sub;
  const int sw = ln / sub;
  const int64_t n_groups = (n_order + group_rows - 1) / group_rows;
  bool ready = false;
  for (;;) {
    unsigned g32 = 0;
    if (ln == 0) g32 = atomicAdd(counters, 1u);
    const int64_t g = __shfl_sync(0xffffffffu, g32, 0);
    if (g >= n_groups) break;
    const int64_t r0 = g * group_rows;
    const int64_t r1 = (r0 + group_rows) < n_order ? (r0 + group_rows) : n_order;
    if (!ready && r1 > n_interior) {
      while (!halo_ready) __nanosleep(128);
      __threadfence();
      ready = true;
    }
    for (int64_t ri = r0 + sw; ri < r1; ri += rows_per_warp) {
      const int64_t row = order[ri];
      const int64_t s = indptr[row], t = indptr[row + 1];
      for (int c = lane; c < nvec; c += sub) {
        float acc[VE];
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] = 0.f;
        int64_t p = s;
        for (; p + 4 <= t; p += 4) {
          V v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __ldcg(reinterpret_cast<const V*>(table + (int64_t)indices[p + k] * F) + c);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const T* q = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
            for (int i = 0; i < VE; ++i) acc[i] += to_f<T>(q[i]);
          }
        }
        for (; p < t; ++p) {
          const V v = __ldcg(reinterpret_cast<const V*>(table + (int64_t)indices[p] * F) + c);
          const T* q = reinterpret_cast<const T*>(&v);
#pragma unroll
          for (int i = 0; i < VE; ++i) acc[i] += to_f<T>(q[i]);
        }
        const float sc = mean ? (t > s ? 1.f / (float)(t - s) : 0.f) : 1.f;
        V o;
        T* po = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int i = 0; i < VE; ++i) po[i] = from_f<T>(acc[i] * sc);
        reinterpret_cast<V*>(out + row * F)[c] = o;
      }
    }
  }
}

static Peers make_peers(const void* const* ptrs, int world) {
  Peers p;
  for (int r = 0; r < kMaxPeers; ++r) p.p[r] = r < world ? const_cast<void*>(ptrs[r]) : nullptr;
  return p;
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_comm_malloc(int64_t bytes, void** out) {
  DGLB_TRY
  DGLB_CUDA_CHECK(cudaMalloc(out, (size_t)bytes));
  DGLB_CUDA_CHECK(cudaMemset(*out, 0, (size_t)bytes));
  DGLB_CATCH
}

DGLB_API int dglb200_comm_free(void* p) {
  DGLB_TRY
  DGLB_CUDA_CHECK(cudaFree(p));
  DGLB_CATCH
}

DGLB_API int dglb200_comm_ipc_get(void* p, void* handle64) {
  DGLB_TRY
  cudaIpcMemHandle_t h;
  DGLB_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(handle64, &h, 64);
  DGLB_CATCH
}

DGLB_API int dglb200_comm_ipc_open(const void* handle64, void** out) {
  DGLB_TRY
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DGLB_CUDA_CHECK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  DGLB_CATCH
}

DGLB_API int dglb200_comm_ipc_close(void* p) {
  DGLB_TRY
  DGLB_CUDA_CHECK(cudaIpcCloseMemHandle(p));
  DGLB_CATCH
}

// single-process multi-device mode (tests / notebooks): enable peer access between two devices
DGLB_API int dglb200_comm_enable_peer(int dev, int peer) {
  DGLB_TRY
  int can = 0;
  DGLB_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dev, peer));
  DGLB_CHECK(can, "devices cannot access each other (no P2P path)");
  int cur;
  DGLB_CUDA_CHECK(cudaGetDevice(&cur));
  DGLB_CUDA_CHECK(cudaSetDevice(dev));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) DGLB_CUDA_CHECK(e);
  cudaGetLastError();
  DGLB_CUDA_CHECK(cudaSetDevice(cur));
  DGLB_CATCH
}

DGLB_API int dglb200_comm_barrier(const void* const* flag_ptrs, int rank, int world, void* epoch_local,
                                  void* error_flag, double timeout_s, void* stream) {
  DGLB_TRY
  DGLB_CHECK(world >= 1 && world <= kMaxPeers, "world must be in [1, 8]");
  Peers f = make_peers(flag_ptrs, world);
  const long long cycles = (long long)(timeout_s * 1.9e9);
  barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, (uint64_t*)epoch_local, (int*)error_flag, cycles);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_comm_allreduce(const void* const* buf_ptrs, int world, void* out, void* out_bf16, int64_t numel,
                                    float scale, void* stream) {
  DGLB_TRY
  DGLB_CHECK(numel % 4 == 0, "numel must be a multiple of 4");
  Peers b = make_peers(buf_ptrs, world);
  int64_t blocks = (numel / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  allreduce_oneshot_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b, world, (float*)out, (__nv_bfloat16*)out_bf16, numel, scale);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_comm_halo_spmm(int dtype, int idtype, const void* indptr, const void* indices, int64_t num_rows,
                                    const void* const* shard_ptrs, int world, int64_t shard_rows, int F, void* out,
                                    int mean, const void* edge_w, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  Peers sh = make_peers(shard_ptrs, world);
  auto launch = [&](auto tag, auto itag) {
    using T = decltype(tag);
    using I = decltype(itag);
    const int64_t row_bytes = (int64_t)F * sizeof(T);
    const int vb = row_bytes % 16 == 0 ? 16 : (row_bytes % 8 == 0 ? 8 : 4);
    DGLB_CHECK(row_bytes % 4 == 0, "row bytes must be a multiple of 4");
    const int nvec = (int)(row_bytes / vb);
    int sub = 1;
    while (sub < nvec && sub < 32) sub <<= 1;
    const int rpb = 256 / sub;
    int64_t blocks = (num_rows + rpb - 1) / rpb;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
#define HS(VBV) halo_spmm_kernel<T, I, VBV><<<(unsigned)blocks, 256, 0, st>>>((const I*)indptr, (const I*)indices, num_rows, sh, shard_rows, F, (T*)out, mean, sub, (const float*)edge_w)
    if (vb == 16) HS(16); else if (vb == 8) HS(8); else HS(4);
#undef HS
  };
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kBF16) launch(__nv_bfloat16(), I());
    else if (dtype == kF16) launch(__half(), I());
    else if (dtype == kF32) launch(float(), I());
    else throw std::runtime_error("halo_spmm: unsupported dtype");
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// One-kernel halo pull + aggregation (HaloPlan). `counters`: four u32 (row-group counter, chunk counter, completed
// chunks, pad), zeroed here on the stream.
DGLB_API int dglb200_comm_halo_pipe(int dtype, int idtype, const void* indptr, const void* indices, const void* order,
                                    int64_t n_order, int64_t n_interior, void* table, const void* const* shard_ptrs,
                                    int world, int64_t shard_rows, const void* halo_ids, int64_t n_halo, int chunk_rows,
                                    int puller_warps, void* counters, int F, void* out, int mean, void* error_flag,
                                    void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  if (n_order == 0) return 0;
  DGLB_CHECK(puller_warps >= 1 && puller_warps <= 8, "halo_pipe: puller warps per CTA must be in [1, 8]");
  Peers sh = make_peers(shard_ptrs, world);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    DGLB_CUDA_CHECK(cudaGetDevice(&dev));
    DGLB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  DGLB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned), st));
  auto launch = [&](auto tag, auto itag) {
    using T = decltype(tag);
    using I = decltype(itag);
    const int64_t row_bytes = (int64_t)F * sizeof(T);
    DGLB_CHECK(row_bytes % 4 == 0, "row bytes must be a multiple of 4");
    const int vb = row_bytes % 16 == 0 ? 16 : (row_bytes % 8 == 0 ? 8 : 4);
    const int nvec = (int)(row_bytes / vb);
    int sub = 1;
    while (sub < nvec && sub < 32) sub <<= 1;
    int group_rows = (32 / sub) * 8;
    int chunk = chunk_rows, np = puller_warps, mn = mean, Fv = F;
    int64_t n_ord = n_order, n_int = n_interior, srows = shard_rows, nh = n_halo;
    const void *ip = indptr, *ix = indices, *od = order, *hid = halo_ids;
    void *tb = table, *ct = counters, *o = out, *ef = error_flag;
    void* args[] = {&ip, &ix, &od, &n_ord, &n_int, &tb, &sh, &srows, &hid, &nh, &chunk, &np, &ct, &Fv, &o, &mn, &sub, &group_rows, &ef};
    auto go = [&](auto kern) {
      int per_sm = 0;
      DGLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
      if (per_sm > 6) per_sm = 6;
      const int blocks = sm_count * (per_sm < 1 ? 1 : per_sm);
      DGLB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kern, dim3(blocks), dim3(256), args, 0, st));
    };
    if (vb == 16) go(halo_pipe_kernel<T, I, 16>);
    else if (vb == 8) go(halo_pipe_kernel<T, I, 8>);
    else go(halo_pipe_kernel<T, I, 4>);
  };
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kBF16) launch(__nv_bfloat16(), I());
    else if (dtype == kF16) launch(__half(), I());
    else if (dtype == kF32) launch(float(), I());
    else throw std::runtime_error("halo_pipe: unsupported dtype");
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_comm_halo_spmm_bwd(int dtype, int idtype, const void* indptr, const void* indices,
                                        int64_t num_rows, const void* const* dshard_ptrs, int world,
                                        int64_t shard_rows, int F, const void* dout, int mean, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_CHECK(F % 4 == 0, "F must be a multiple of 4");
  Peers sh = make_peers(dshard_ptrs, world);
  int sub = 1;
  while (sub < F / 4 && sub < 32) sub <<= 1;
  const int rpb = 256 / sub;
  int64_t blocks = (num_rows + rpb - 1) / rpb;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kBF16) halo_spmm_bwd_kernel<__nv_bfloat16, I><<<(unsigned)blocks, 256, 0, st>>>((const I*)indptr, (const I*)indices, num_rows, sh, shard_rows, F, (const __nv_bfloat16*)dout, mean, sub);
    else if (dtype == kF32) halo_spmm_bwd_kernel<float, I><<<(unsigned)blocks, 256, 0, st>>>((const I*)indptr, (const I*)indices, num_rows, sh, shard_rows, F, (const float*)dout, mean, sub);
    else if (dtype == kF16) halo_spmm_bwd_kernel<__half, I><<<(unsigned)blocks, 256, 0, st>>>((const I*)indptr, (const I*)indices, num_rows, sh, shard_rows, F, (const __half*)dout, mean, sub);
    else throw std::runtime_error("halo_spmm_bwd: unsupported dtype");
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_comm_p2p_gather(int idtype, const void* const* shard_ptrs, int world, int64_t shard_rows,
                                     const void* idx, int64_t n, void* out, int64_t row_bytes, void* stream) {
  DGLB_TRY
  DGLB_CHECK(row_bytes % 16 == 0, "row bytes must be a multiple of 16");
  Peers sh = make_peers(shard_ptrs, world);
  int sub = 1;
  while (sub < row_bytes / 16 && sub < 32) sub <<= 1;
  const int rpb = 256 / sub;
  int64_t blocks = (n + rpb - 1) / rpb;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  DGLB_ID_SWITCH(idtype, I, {
    p2p_gather_kernel<I><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sh, shard_rows, (const I*)idx, n, (char*)out, row_bytes, sub);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_comm_p2p_scatter_add(int idtype, const void* const* shard_ptrs, int world, int64_t shard_rows,
                                          const void* idx, int64_t n, const void* vals, int F, void* stream) {
  DGLB_TRY
  DGLB_CHECK(F % 4 == 0, "F must be a multiple of 4");
  Peers sh = make_peers(shard_ptrs, world);
  int sub = 1;
  while (sub < F / 4 && sub < 32) sub <<= 1;
  const int rpb = 256 / sub;
  int64_t blocks = (n + rpb - 1) / rpb;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  DGLB_ID_SWITCH(idtype, I, {
    p2p_scatter_add_kernel<I><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sh, shard_rows, (const I*)idx, n, (const float*)vals, F, sub);
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
