// Fused GraphSAGE mini-batch training engine kernels (sm_100a).
//
// The whole step (sample 3 layers -> compact -> aggregate+concat -> GEMM -> bias/ReLU/dropout ->
// loss -> backward -> gradient all-reduce -> Adam) runs with ZERO host synchronisation: every
// buffer is sized by a static upper bound, actual row counts live in device memory (`n_dev`) and
// every kernel masks on them, so the entire step is captured once in a CUDA graph and replayed.
// The reference performs >= 2 host syncs per sampled layer plus ~60 launches per step
// (rowwise_sampling.cu:304-335, cuda_to_block.cu:193-212).
//
// Design notes
//  * Sampled neighbourhoods are ELL blocks [rows, fanout] (perfectly regular: no load imbalance).
//  * The input layer aggregates straight from the global feature table using global ids, so the
//    largest frontier is never deduplicated and its features are never materialised (fusion of
//    "feature gather" with "first-layer aggregation").
//  * agg_concat writes the GEMM A-operand [self | mean(neigh)] in bf16; the tcgen05 kernel in
//    sage_tc.cu consumes the same layout straight from shared memory.
#include "common.cuh"

namespace dglb {

// ----------------------------------------------------------------------------------------------
// A[d, 0:F]  = X[self_ids ? self_ids[d] : d]
// A[d, F:2F] = mean_{k < cnt[d]} X[nbr[d, k]]        (rows >= *n_dev are zero-filled)
// sub lanes per row, VB-byte vectors, fp32 accumulation, output bf16/T.
template <typename T, typename I, int VB>
__global__ void __launch_bounds__(256) agg_concat_kernel(const T* __restrict__ X, const I* __restrict__ self_ids,
                                                         const I* __restrict__ nbr, const int32_t* __restrict__ cnt,
                                                         const int64_t* __restrict__ n_dev, int64_t n_cap, int fanout,
                                                         int F, T* __restrict__ A, int sub, int zero_pad) {
  constexpr int VE = VB / sizeof(T);
  using V = typename std::conditional<VB == 16, uint4, typename std::conditional<VB == 8, uint2, uint32_t>::type>::type;
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int nvec = F / VE;
  const int64_t n = min(*n_dev, n_cap);
  const int64_t n_loop = zero_pad ? n_cap : n;
  for (int64_t d = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; d < n_loop;
       d += (int64_t)gridDim.x * rows_per_block) {
    V* arow = reinterpret_cast<V*>(A + d * 2 * F);
    if (d >= n) {
      V z;
      memset(&z, 0, sizeof(V));
      for (int c = lane; c < 2 * nvec; c += sub) arow[c] = z;
      continue;
    }
    const int64_t sid = self_ids ? (int64_t)self_ids[d] : d;
    const V* srow = reinterpret_cast<const V*>(X + sid * F);
    const int c_ = cnt[d];
    const float inv = c_ > 0 ? 1.f / (float)c_ : 0.f;
    const I* nb = nbr + d * fanout;
    for (int c = lane; c < nvec; c += sub) {
      arow[c] = __ldg(srow + c);
      float acc[VE];
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[i] = 0.f;
      int k = 0;
      for (; k + 4 <= c_; k += 4) {  // 4 neighbour rows in flight
        // ids < 0 only appear after a (flagged) capacity overflow: clamp to row 0, weight 0 is not
        // needed because the overflowing step is reported and capacities are raised
        const int64_t u0 = max((int64_t)nb[k], (int64_t)0), u1 = max((int64_t)nb[k + 1], (int64_t)0),
                      u2 = max((int64_t)nb[k + 2], (int64_t)0), u3 = max((int64_t)nb[k + 3], (int64_t)0);
        V v0 = __ldg(reinterpret_cast<const V*>(X + u0 * F) + c);
        V v1 = __ldg(reinterpret_cast<const V*>(X + u1 * F) + c);
        V v2 = __ldg(reinterpret_cast<const V*>(X + u2 * F) + c);
        V v3 = __ldg(reinterpret_cast<const V*>(X + u3 * F) + c);
        const T* p0 = reinterpret_cast<const T*>(&v0);
        const T* p1 = reinterpret_cast<const T*>(&v1);
        const T* p2 = reinterpret_cast<const T*>(&v2);
        const T* p3 = reinterpret_cast<const T*>(&v3);
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] += (to_f<T>(p0[i]) + to_f<T>(p1[i])) + (to_f<T>(p2[i]) + to_f<T>(p3[i]));
      }
      for (; k < c_; ++k) {
        const int64_t u = max((int64_t)nb[k], (int64_t)0);
        V v = __ldg(reinterpret_cast<const V*>(X + u * F) + c);
        const T* p = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] += to_f<T>(p[i]);
      }
      V o;
      T* po = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int i = 0; i < VE; ++i) po[i] = from_f<T>(acc[i] * inv);
      arow[nvec + c] = o;
    }
  }
}

// H[r, c] = dropout(relu(Z[r, c] + b[c]))  (in place, rows >= n -> 0). RNG: (seed + *step, r*N+c).
template <typename T>
__global__ void __launch_bounds__(256) bias_relu_dropout_kernel(T* __restrict__ Z, const float* __restrict__ bias,
                                                                const int64_t* __restrict__ n_dev, int64_t n_cap,
                                                                int N, float p_drop, uint64_t seed,
                                                                const int64_t* __restrict__ step_dev, int relu) {
  const int64_t n = *n_dev;
  const uint64_t sd = seed + (step_dev ? (uint64_t)(*step_dev) * 0x9E3779B97F4A7C15ull : 0ull);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int64_t total = min(n, n_cap) * N;  // rows >= n are never consumed
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < total;
       i += (int64_t)gridDim.x * blockDim.x * 2) {
    const int64_t r = i / N;
    const int c = (int)(i - r * N);
    float v0 = 0.f, v1 = 0.f;
    {
      v0 = to_f<T>(Z[i]) + bias[c];
      v1 = to_f<T>(Z[i + 1]) + bias[c + 1];
      if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      if (p_drop > 0.f) {
        const uint64_t rb = rng_u64(sd, 0x51ull, (uint64_t)i);
        const float u0 = (float)(rb & 0xffffff) * (1.f / 16777216.f);
        const float u1 = (float)((rb >> 24) & 0xffffff) * (1.f / 16777216.f);
        v0 = u0 < p_drop ? 0.f : v0 * keep_scale;
        v1 = u1 < p_drop ? 0.f : v1 * keep_scale;
      }
    }
    Z[i] = from_f<T>(v0);
    Z[i + 1] = from_f<T>(v1);
  }
}

// Cross entropy: loss += sum_i nll_i / n ; dlogits = (softmax - onehot)/n (rows >= n -> 0).
// One warp per row. logits fp32 [n_cap, C] (bias already added by caller when bias == nullptr).
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, const float* __restrict__ bias,
                                                      const int32_t* __restrict__ labels_all,
                                                      const int32_t* __restrict__ seeds,
                                                      const int64_t* __restrict__ n_dev, int64_t n_cap, int C,
                                                      float* __restrict__ dlogits, float* __restrict__ loss_out,
                                                      int32_t* __restrict__ correct_out) {
  const int64_t n = *n_dev;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_cap) return;
  if (r >= n) {
    for (int c = lane; c < C; c += 32) dlogits[r * C + c] = 0.f;
    return;
  }
  const int y = labels_all[seeds[r]];
  float mx = -INFINITY;
  int amax = 0;
  for (int c = lane; c < C; c += 32) {
    const float v = logits[r * C + c] + (bias ? bias[c] : 0.f);
    if (v > mx) { mx = v; amax = c; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
    if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
  }
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += __expf(logits[r * C + c] + (bias ? bias[c] : 0.f) - mx);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv_n = 1.f / (float)n;
  const float lse = mx + __logf(sum);
  for (int c = lane; c < C; c += 32) {
    const float v = logits[r * C + c] + (bias ? bias[c] : 0.f);
    const float p = __expf(v - lse);
    dlogits[r * C + c] = (p - (c == y ? 1.f : 0.f)) * inv_n;
  }
  if (lane == 0) {
    const float vy = logits[r * C + y] + (bias ? bias[y] : 0.f);
    atomicAdd(loss_out, (lse - vy) * inv_n);
    if (correct_out && amax == y) atomicAdd(correct_out, 1);
  }
}

// Cross entropy, second generation: strided logits (e.g. the 64-column padded output of the fused layer kernel),
// optional bias, gradient emitted in fp32 [n_cap, C] and/or bf16 [n_cap, ldd] (padding columns zeroed, ready to be the
// operand of the backward GEMMs) and the bias gradient reduced in the same pass (block partial sums -> atomics).
__global__ void __launch_bounds__(256) ce_loss2_kernel(const float* __restrict__ logits, int ldl,
                                                       const float* __restrict__ bias,
                                                       const int32_t* __restrict__ labels_all,
                                                       const int32_t* __restrict__ seeds,
                                                       const int64_t* __restrict__ n_dev, int64_t n_cap, int C,
                                                       float* __restrict__ dlog32, __nv_bfloat16* __restrict__ dlog16,
                                                       int ldd, float* __restrict__ db, float* __restrict__ loss_out,
                                                       int32_t* __restrict__ correct_out, float* __restrict__ dlog_hi,
                                                       float* __restrict__ dlog_lo) {
  __shared__ float s_db[8][64];
  const int64_t n = min(*n_dev, n_cap);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 8 + w;
  float g[2] = {0.f, 0.f};  // this lane's columns: lane, lane + 32   (C <= 64)
  if (r < n) {
    const int y = labels_all[seeds[r]];
    float v[2];
    float mx = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = lane + 32 * q;
      v[q] = c < C ? logits[r * ldl + c] + (bias ? bias[c] : 0.f) : -INFINITY;
      if (v[q] > mx) { mx = v[q]; amax = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
      if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
    }
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) sum += lane + 32 * q < C ? __expf(v[q] - mx) : 0.f;
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv_n = 1.f / (float)n;
    const float lse = mx + __logf(sum);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = lane + 32 * q;
      if (c < C) g[q] = (__expf(v[q] - lse) - (c == y ? 1.f : 0.f)) * inv_n;
    }
    const float vy = __shfl_sync(0xffffffffu, v[y >> 5], y & 31);
    if (lane == 0) {
      atomicAdd(loss_out, (lse - vy) * inv_n);
      if (correct_out && amax == y) atomicAdd(correct_out, 1);
    }
  }
  if (r < n_cap) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = lane + 32 * q;
      if (dlog_lo) {
        // fp32 tensor-core path: the gradient in full ([n_cap, ldd], padding zeroed: operand of the dW kernel, which
        // splits on the fly) and as the hi / lo planes of its TF32 split (TMA operands of the dA kernel)
        if (c < ldd) {
          uint32_t h;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(g[q]));
          const float hi = __uint_as_float(h & 0xFFFFE000u);
          dlog32[r * ldd + c] = g[q];
          dlog_hi[r * ldd + c] = hi;
          dlog_lo[r * ldd + c] = g[q] - hi;
        }
      } else if (dlog32 && c < C) {
        dlog32[r * C + c] = g[q];
      }
      if (dlog16 && c < ldd) dlog16[r * ldd + c] = __float2bfloat16(g[q]);
    }
  }
  if (db) {
    s_db[w][lane] = g[0];
    s_db[w][lane + 32] = g[1];
    __syncthreads();
    if (threadIdx.x < C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += s_db[k][threadIdx.x];
      if (t != 0.f) atomicAdd(db + threadIdx.x, t);
    }
  }
}

// dH[self(d)] += dA[d, 0:F];  dH[nbr[d,k]] += dA[d, F:2F] / cnt[d]     (dH fp32, pre-zeroed)
template <typename T, typename I>
__global__ void __launch_bounds__(256) scatter_grad_kernel(const T* __restrict__ dA, const I* __restrict__ nbr,
                                                           const int32_t* __restrict__ cnt,
                                                           const int64_t* __restrict__ n_dev, int fanout, int F,
                                                           float* __restrict__ dH, int sub) {
  const int rows_per_block = blockDim.x / sub;
  const int lane = threadIdx.x % sub;
  const int64_t n = *n_dev;
  for (int64_t d = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / sub; d < n;
       d += (int64_t)gridDim.x * rows_per_block) {
    const T* row = dA + d * 2 * F;
    const int c_ = cnt[d];
    const float inv = c_ > 0 ? 1.f / (float)c_ : 0.f;
    const I* nb = nbr + d * fanout;
    for (int f = lane * 4; f < F; f += sub * 4) {
      float s0 = to_f<T>(row[f]), s1 = to_f<T>(row[f + 1]), s2 = to_f<T>(row[f + 2]), s3 = to_f<T>(row[f + 3]);
      float* dst = dH + d * F + f;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(s0), "f"(s1), "f"(s2), "f"(s3) : "memory");
      const float g0 = to_f<T>(row[F + f]) * inv, g1 = to_f<T>(row[F + f + 1]) * inv,
                  g2 = to_f<T>(row[F + f + 2]) * inv, g3 = to_f<T>(row[F + f + 3]) * inv;
      for (int k = 0; k < c_; ++k) {
        float* dn = dH + (int64_t)nb[k] * F + f;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dn), "f"(g0), "f"(g1), "f"(g2), "f"(g3) : "memory");
      }
    }
  }
}

// dZ[r,c] = dH[r,c] * (H[r,c] > 0) * scale for r < n, 0 for n <= r < n_cap;  db[c] += sum_r dZ[r,c].
// 8 columns per thread (16B bf16 / 32B fp32 vectors), N/8 threads per row, block-level column sums
// in shared memory, one atomicAdd per column per block.
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dH, const T* __restrict__ H,
                                                       const int64_t* __restrict__ n_dev, int64_t n_cap, int N,
                                                       float scale, T* __restrict__ dZ, float* __restrict__ db,
                                                       int rows_per_block) {
  extern __shared__ float s_db[];
  const int64_t n = min(*n_dev, n_cap);
  const int tpr = N / 8;                      // threads per row
  const int rpi = blockDim.x / tpr;           // rows per iteration
  const int c0 = (threadIdx.x % tpr) * 8;
  const int rsub = threadIdx.x / tpr;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_db[i] = 0.f;
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int64_t r = r0 + rsub; r < r0 + rows_per_block && r < n_cap; r += rpi) {
    T out[8];
    if (r < n) {
      const float4 g0 = *reinterpret_cast<const float4*>(dH + r * N + c0);
      const float4 g1 = *reinterpret_cast<const float4*>(dH + r * N + c0 + 4);
      T h[8];
      if (sizeof(T) == 2) *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(H + r * N + c0);
      else {
        *reinterpret_cast<float4*>(h) = *reinterpret_cast<const float4*>(H + r * N + c0);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(h) + 4) = *reinterpret_cast<const float4*>(H + r * N + c0 + 4);
      }
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = to_f<T>(h[k]) > 0.f ? g[k] * scale : 0.f;
        out[k] = from_f<T>(v);
        acc[k] += to_f<T>(out[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) out[k] = from_f<T>(0.f);
    }
    if (sizeof(T) == 2) *reinterpret_cast<uint4*>(dZ + r * N + c0) = *reinterpret_cast<uint4*>(out);
    else {
      *reinterpret_cast<float4*>(dZ + r * N + c0) = *reinterpret_cast<float4*>(out);
      *reinterpret_cast<float4*>(dZ + r * N + c0 + 4) = *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + 4);
    }
  }
  if (db) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (acc[k] != 0.f) atomicAdd(&s_db[c0 + k], acc[k]);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x)
      if (s_db[i] != 0.f) atomicAdd(db + i, s_db[i]);
  }
}

// zero the first *n_dev rows of a row-major matrix (row_bytes multiple of 16)
__global__ void __launch_bounds__(256) zero_rows_kernel(uint4* __restrict__ p, const int64_t* __restrict__ n_dev,
                                                        int64_t n_cap, int64_t vec_per_row) {
  const int64_t total = min(*n_dev, n_cap) * vec_per_row;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) p[i] = z;
}

// column sums of a fp32 matrix (bias grad of the output layer)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t rows, int N,
                                                     float* __restrict__ out) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int64_t r = 0; r < rows; ++r) acc += X[r * N + c];
    out[c] = acc;
  }
}

// Fused (all-reduce ->) Adam. grads of all ranks are read through peer pointers (NVLink P2P
// loads over NVSwitch), averaged, and consumed by the optimizer update in the same kernel; the
// bf16 shadow copy used by the tensor-core GEMMs is refreshed in the epilogue.
// peers[r] = pointer to rank r's flat fp32 gradient buffer (peers[my_rank] = local).
struct PeerPtrs {
  const float* p[8];
};

__global__ void __launch_bounds__(256) allreduce_adam_kernel(float* __restrict__ param, PeerPtrs peers, int world,
                                                             float* __restrict__ m, float* __restrict__ v,
                                                             __nv_bfloat16* __restrict__ shadow, int64_t numel,
                                                             float lr, float beta1, float beta2, float eps,
                                                             float weight_decay, int64_t* __restrict__ step_dev,
                                                             int bump_step, int64_t lo = 0,
                                                             float* __restrict__ param_pub = nullptr,
                                                             float* __restrict__ shadow_hi = nullptr,
                                                             float* __restrict__ shadow_lo = nullptr) {
  // [lo, numel) is the element range this launch owns: the whole buffer for the one-shot all-reduce, this rank's slice
  // for the two-shot variant (reduce-scatter + Adam on the slice, updated fp32 values published in param_pub)
  const int64_t t = *step_dev + 1;
  const float bc1 = 1.f - __powf(beta1, (float)t);
  const float bc2 = 1.f - __powf(beta2, (float)t);
  const float inv_world = 1.f / (float)world;
  for (int64_t i = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < numel;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int r = 0; r < world; ++r) {
      float4 x;  // volatile: peer lines may sit in L1 from a previous step
      asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(peers.p[r] + i));
      g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
    }
    float4 p4 = *reinterpret_cast<float4*>(param + i);
    float4 m4 = *reinterpret_cast<float4*>(m + i);
    float4 v4 = *reinterpret_cast<float4*>(v + i);
    float* gp = reinterpret_cast<float*>(&g);
    float* pp = reinterpret_cast<float*>(&p4);
    float* mp = reinterpret_cast<float*>(&m4);
    float* vp = reinterpret_cast<float*>(&v4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = gp[k] * inv_world + weight_decay * pp[k];
      mp[k] = beta1 * mp[k] + (1.f - beta1) * gg;
      vp[k] = beta2 * vp[k] + (1.f - beta2) * gg * gg;
      const float mh = mp[k] / bc1;
      const float vh = vp[k] / bc2;
      pp[k] -= lr * mh / (sqrtf(vh) + eps);
    }
    *reinterpret_cast<float4*>(param + i) = p4;
    if (param_pub) *reinterpret_cast<float4*>(param_pub + i) = p4;
    *reinterpret_cast<float4*>(m + i) = m4;
    *reinterpret_cast<float4*>(v + i) = v4;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]);
      __nv_bfloat162 hi = __floats2bfloat162_rn(pp[2], pp[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(shadow + i) = pk;
    }
    if (shadow_hi) {
      // fp32 tensor-core operands: x = hi + lo, hi exactly representable in TF32 (see tc_common.cuh: tf32_split)
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
      *reinterpret_cast<float4*>(shadow_hi + i) = h4;
      *reinterpret_cast<float4*>(shadow_lo + i) = l4;
    }
  }
  if (bump_step && blockIdx.x == 0 && threadIdx.x == 0) {
    // grid-wide ordering is not needed: t was read by every thread before any block could finish
    // a full pass only if the counter is bumped by a trailing 1-thread kernel; see bump_step_kernel.
  }
}

__global__ void bump_step_kernel(int64_t* step_dev) { *step_dev += 1; }

// seeds_dev[i] = train_ids[perm[(cursor + i) % n_train]] helper for device-side epoch iteration
__global__ void __launch_bounds__(256) take_batch_kernel(const int32_t* __restrict__ ids,
                                                         const int64_t* __restrict__ cursor_dev, int64_t n_ids,
                                                         int64_t batch, int32_t* __restrict__ out,
                                                         int64_t* __restrict__ n_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cur = *cursor_dev;
  if (i < batch) out[i] = (cur + i < n_ids) ? ids[cur + i] : -1;
  if (i == 0) *n_out = min(batch, max((int64_t)0, n_ids - cur));
}

}  // namespace dglb

using namespace dglb;

static int pick_sub(int nvec) {
  int sub = 1;
  while (sub < nvec && sub < 32) sub <<= 1;
  return sub;
}

// dtype: kBF16 / kF16 / kF32 for X and A
DGLB_API int dglb200_engine_agg_concat(int dtype, int idtype, const void* X, const void* self_ids, const void* nbr,
                                       const void* cnt, const void* n_dev, int64_t n_cap, int fanout, int F, void* A,
                                       int zero_pad, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  if (n_cap == 0) return 0;
  auto launch = [&](auto tag, auto itag) {
    using T = decltype(tag);
    using I = decltype(itag);
    const int64_t row_bytes = (int64_t)F * sizeof(T);
    int vb = (row_bytes % 16 == 0 && ((uintptr_t)X % 16 == 0) && ((uintptr_t)A % 16 == 0)) ? 16 : (row_bytes % 8 == 0 ? 8 : 4);
    DGLB_CHECK(row_bytes % 4 == 0, "feature row bytes must be a multiple of 4");
    const int nvec = (int)(row_bytes / vb);
    const int sub = pick_sub(nvec);
    const int rpb = 256 / sub;
    int64_t blocks = (n_cap + rpb - 1) / rpb;
    if (blocks > 148 * 16) blocks = 148 * 16;
#define AC(VBV) agg_concat_kernel<T, I, VBV><<<(unsigned)blocks, 256, 0, st>>>((const T*)X, (const I*)self_ids, (const I*)nbr, (const int32_t*)cnt, (const int64_t*)n_dev, n_cap, fanout, F, (T*)A, sub, zero_pad)
    if (vb == 16) AC(16); else if (vb == 8) AC(8); else AC(4);
#undef AC
  };
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kBF16) launch(__nv_bfloat16(), I());
    else if (dtype == kF16) launch(__half(), I());
    else if (dtype == kF32) launch(float(), I());
    else throw std::runtime_error("agg_concat: unsupported dtype");
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_bias_relu_dropout(int dtype, void* Z, const void* bias, const void* n_dev, int64_t n_cap,
                                              int N, float p_drop, int64_t seed, const void* step_dev, int relu,
                                              void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_CHECK(N % 2 == 0, "N must be even");
  int64_t blocks = (n_cap * N / 2 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (dtype == kBF16)
    bias_relu_dropout_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)Z, (const float*)bias, (const int64_t*)n_dev, n_cap, N, p_drop, (uint64_t)seed, (const int64_t*)step_dev, relu);
  else if (dtype == kF16)
    bias_relu_dropout_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((__half*)Z, (const float*)bias, (const int64_t*)n_dev, n_cap, N, p_drop, (uint64_t)seed, (const int64_t*)step_dev, relu);
  else if (dtype == kF32)
    bias_relu_dropout_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)Z, (const float*)bias, (const int64_t*)n_dev, n_cap, N, p_drop, (uint64_t)seed, (const int64_t*)step_dev, relu);
  else throw std::runtime_error("unsupported dtype");
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_ce_loss(const void* logits, const void* bias, const void* labels_all, const void* seeds,
                                    const void* n_dev, int64_t n_cap, int C, void* dlogits, void* loss_out,
                                    void* correct_out, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  ce_loss_kernel<<<(unsigned)((n_cap + 7) / 8), 256, 0, st>>>((const float*)logits, (const float*)bias, (const int32_t*)labels_all, (const int32_t*)seeds, (const int64_t*)n_dev, n_cap, C, (float*)dlogits, (float*)loss_out, (int32_t*)correct_out);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_ce_loss2(const void* logits, int ldl, const void* bias, const void* labels_all,
                                     const void* seeds, const void* n_dev, int64_t n_cap, int C, void* dlog32,
                                     void* dlog16, int ldd, void* db, void* loss_out, void* correct_out, void* dlog_hi,
                                     void* dlog_lo, void* stream) {
  DGLB_TRY
  DGLB_CHECK(C >= 1 && C <= 64 && ldd <= 64, "ce_loss2 supports up to 64 classes");
  if (n_cap == 0) return 0;
  ce_loss2_kernel<<<(unsigned)((n_cap + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const float*)logits, ldl, (const float*)bias, (const int32_t*)labels_all, (const int32_t*)seeds,
      (const int64_t*)n_dev, n_cap, C, (float*)dlog32, (__nv_bfloat16*)dlog16, ldd, (float*)db, (float*)loss_out,
      (int32_t*)correct_out, (float*)dlog_hi, (float*)dlog_lo);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_scatter_grad(int dtype, int idtype, const void* dA, const void* nbr, const void* cnt,
                                         const void* n_dev, int64_t n_cap, int fanout, int F, void* dH, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_CHECK(F % 4 == 0, "F must be a multiple of 4");
  const int sub = pick_sub(F / 4);
  const int rpb = 256 / sub;
  int64_t blocks = (n_cap + rpb - 1) / rpb;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  DGLB_ID_SWITCH(idtype, I, {
    if (dtype == kBF16) scatter_grad_kernel<__nv_bfloat16, I><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)dA, (const I*)nbr, (const int32_t*)cnt, (const int64_t*)n_dev, fanout, F, (float*)dH, sub);
    else if (dtype == kF32) scatter_grad_kernel<float, I><<<(unsigned)blocks, 256, 0, st>>>((const float*)dA, (const I*)nbr, (const int32_t*)cnt, (const int64_t*)n_dev, fanout, F, (float*)dH, sub);
    else if (dtype == kF16) scatter_grad_kernel<__half, I><<<(unsigned)blocks, 256, 0, st>>>((const __half*)dA, (const I*)nbr, (const int32_t*)cnt, (const int64_t*)n_dev, fanout, F, (float*)dH, sub);
    else throw std::runtime_error("unsupported dtype");
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_relu_bwd(int dtype, const void* dH, const void* H, const void* n_dev, int64_t n_cap, int N,
                                     float scale, void* dZ, void* db, void* stream) {
  DGLB_TRY
  cudaStream_t st = (cudaStream_t)stream;
  DGLB_CHECK(N % 8 == 0 && 256 % (N / 8) == 0 && N <= 2048, "relu_bwd: N must be a multiple of 8 dividing 2048");
  const int rpb = 64;
  const unsigned blocks = (unsigned)((n_cap + rpb - 1) / rpb);
  if (blocks == 0) return 0;
  const size_t smem = (size_t)N * sizeof(float);
  if (dtype == kBF16) relu_bwd_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>((const float*)dH, (const __nv_bfloat16*)H, (const int64_t*)n_dev, n_cap, N, scale, (__nv_bfloat16*)dZ, (float*)db, rpb);
  else if (dtype == kF16) relu_bwd_kernel<__half><<<blocks, 256, smem, st>>>((const float*)dH, (const __half*)H, (const int64_t*)n_dev, n_cap, N, scale, (__half*)dZ, (float*)db, rpb);
  else if (dtype == kF32) relu_bwd_kernel<float><<<blocks, 256, smem, st>>>((const float*)dH, (const float*)H, (const int64_t*)n_dev, n_cap, N, scale, (float*)dZ, (float*)db, rpb);
  else throw std::runtime_error("unsupported dtype");
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_zero_rows(void* p, const void* n_dev, int64_t n_cap, int64_t row_bytes, void* stream) {
  DGLB_TRY
  DGLB_CHECK(row_bytes % 16 == 0, "row bytes must be a multiple of 16");
  int64_t blocks = (n_cap * (row_bytes / 16) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  zero_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint4*)p, (const int64_t*)n_dev, n_cap, row_bytes / 16);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_colsum(const void* X, int64_t rows, int N, void* out, void* stream) {
  DGLB_TRY
  colsum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float*)X, rows, N, (float*)out);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// second half of the two-shot all-reduce + Adam: every rank pulls the slices the OTHER ranks updated (fp32 parameters in
// their symmetric param buffers) into its local parameters and bf16 shadow. slice r = [r * slice, min(numel, (r+1) * slice)).
__global__ void __launch_bounds__(256) param_allgather_kernel(float* __restrict__ param, PeerPtrs pubs, int world, int rank,
                                                              __nv_bfloat16* __restrict__ shadow, int64_t numel,
                                                              int64_t slice) {
  const int64_t mine_lo = (int64_t)rank * slice, mine_hi = min(numel, mine_lo + slice);
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < numel;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    if (i >= mine_lo && i < mine_hi) continue;
    const int r = (int)(i / slice);
    float4 x;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(pubs.p[r] + i));
    *reinterpret_cast<float4*>(param + i) = x;
    if (shadow) {
      __nv_bfloat162 l2 = __floats2bfloat162_rn(x.x, x.y), h2 = __floats2bfloat162_rn(x.z, x.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&l2);
      pk.y = *reinterpret_cast<uint32_t*>(&h2);
      *reinterpret_cast<uint2*>(shadow + i) = pk;
    }
  }
}

// peer_ptrs: host array of `world` device pointers (each rank's flat fp32 grad buffer, P2P mapped)
DGLB_API int dglb200_engine_allreduce_adam(void* param, const void* const* peer_ptrs, int world, void* m, void* v,
                                           void* shadow, int64_t numel, float lr, float beta1, float beta2, float eps,
                                           float weight_decay, void* step_dev, void* shadow_hi, void* shadow_lo,
                                           void* stream) {
  DGLB_TRY
  DGLB_CHECK(world >= 1 && world <= 8, "world must be in [1, 8]");
  DGLB_CHECK(numel % 4 == 0, "flat parameter buffer must be padded to a multiple of 4 elements");
  PeerPtrs pp;
  for (int r = 0; r < 8; ++r) pp.p[r] = r < world ? (const float*)peer_ptrs[r] : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (numel / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  allreduce_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>((float*)param, pp, world, (float*)m, (float*)v, (__nv_bfloat16*)shadow, numel, lr, beta1, beta2, eps, weight_decay, (int64_t*)step_dev, 0, 0, nullptr, (float*)shadow_hi, (float*)shadow_lo);
  bump_step_kernel<<<1, 1, 0, st>>>((int64_t*)step_dev);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// Two-shot variant for large parameter sets (bandwidth-optimal: 2 * (world-1)/world * bytes per GPU instead of
// (world-1) * bytes): phase 1 reduces + applies Adam on this rank's slice and publishes the new fp32 values in its
// symmetric param buffer; after a device barrier phase 2 (dglb200_engine_param_allgather) pulls the other slices.
DGLB_API int dglb200_engine_allreduce_adam_rs(void* param, const void* const* peer_grad_ptrs, int world, int rank, void* m,
                                              void* v, void* shadow, int64_t numel, int64_t slice, float lr, float beta1,
                                              float beta2, float eps, float weight_decay, void* step_dev, void* param_pub,
                                              void* stream) {
  DGLB_TRY
  DGLB_CHECK(world >= 1 && world <= 8 && rank >= 0 && rank < world, "bad world / rank");
  DGLB_CHECK(numel % 4 == 0 && slice % 4 == 0 && slice * world >= numel, "numel and slice must be multiples of 4, slice * world >= numel");
  PeerPtrs pp;
  for (int r = 0; r < 8; ++r) pp.p[r] = r < world ? (const float*)peer_grad_ptrs[r] : nullptr;
  const int64_t lo = (int64_t)rank * slice, hi = std::min<int64_t>(numel, lo + slice);
  if (hi > lo) {
    int64_t blocks = ((hi - lo) / 4 + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    allreduce_adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)param, pp, world, (float*)m, (float*)v, (__nv_bfloat16*)shadow, hi, lr, beta1, beta2, eps, weight_decay, (int64_t*)step_dev, 0, lo, (float*)param_pub);
  }
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_param_allgather(void* param, const void* const* peer_pub_ptrs, int world, int rank, void* shadow,
                                            int64_t numel, int64_t slice, void* step_dev, void* stream) {
  DGLB_TRY
  PeerPtrs pp;
  for (int r = 0; r < 8; ++r) pp.p[r] = r < world ? (const float*)peer_pub_ptrs[r] : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (numel / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  if (world > 1) param_allgather_kernel<<<(unsigned)blocks, 256, 0, st>>>((float*)param, pp, world, rank, (__nv_bfloat16*)shadow, numel, slice);
  if (step_dev) bump_step_kernel<<<1, 1, 0, st>>>((int64_t*)step_dev);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_engine_take_batch(const void* ids, const void* cursor_dev, int64_t n_ids, int64_t batch, void* out,
                                       void* n_out, void* stream) {
  DGLB_TRY
  take_batch_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const int32_t*)ids, (const int64_t*)cursor_dev, n_ids, batch, (int32_t*)out, (int64_t*)n_out);
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
