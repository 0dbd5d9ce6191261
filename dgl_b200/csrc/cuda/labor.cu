// LABOR layer sampling with importance iterations and continuous-seed variates, all on the device.
//
// Notation (Balin & Catalyurek, "Layer-Neighbor Sampling"): seed s keeps in-edge (t -> s) iff r_t <= min(1, c_s * pi_t * a_ts)
// where r_t is ONE uniform variate per source vertex shared by every seed of the layer, a_ts an optional edge weight,
// pi_t the per-vertex importance (LABOR-0: 1) and c_s the per-seed normaliser solving  sum_t min(1, c_s pi_t a_ts) = k_s,
// k_s = min(fanout, #eligible in-edges). LABOR-i repeats  pi_t <- max_s c_s pi_t a_ts ; re-solve c  i times ("*" = to
// convergence), which shrinks the expected number of distinct sources.
//
//   labor_solve_c    one 8-lane group per seed. c_s by the monotone fixed point  c <- (k - |saturated|) / sum_unsaturated w
//                    (f(c) = sum min(1, c w) is concave piecewise linear, so the iteration converges from below in a handful
//                    of passes over the row) - no bisection, no host round trips.
//   labor_pi_max     pi_next[t] = max over seeds of c_s * pi_t * a_ts via atomicMax on the float bit pattern (values >= 0).
//   labor_pi_commit  pi[t] <- pi_next[t]  /  pi_next[t] <- 0  /  pi[t] <- 1   over the touched vertices only (the buffers are
//                    indexed by global vertex id and cached between calls, so no O(N) pass per mini-batch).
//   labor_pick       count / fill passes: Bernoulli(r_t <= p) per edge with the shared variate, importance weights 1 / p
//                    Hajek-normalised per seed. The variate interpolates between two seeds,
//                    r = Phi(cos(pi/2 * x) Phi^-1(r1) + sin(pi/2 * x) Phi^-1(r2)), so consecutive mini-batches can move
//                    smoothly from one random stream to the next (batch dependency).
// Reference counterparts: src/array/cuda/labor_sampling.cu:206-274 (importance iterations with CUB segmented reductions
// and host-side loops), graphbolt/include/graphbolt/continuous_seed.h:20-75, graphbolt/src/cuda/neighbor_sampler.cu:118-145.
#include "common.cuh"

namespace dglb {
namespace labor {

constexpr int kGroup = 8;

__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = kGroup >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

__device__ __forceinline__ float variate(uint64_t seed1, uint64_t seed2, float mix, uint64_t vertex) {
  const float r1 = rng_uniform(seed1, 0x4c41424full, vertex);
  if (mix <= 0.f) return r1;
  const float r2 = rng_uniform(seed2, 0x4c41424full, vertex);
  float s, c;
  sincospif(0.5f * mix, &s, &c);
  return normcdff(c * normcdfinvf(r1) + s * normcdfinvf(r2));
}

template <typename I, typename W>
__device__ __forceinline__ float edge_weight(const W* __restrict__ prob, const I* __restrict__ eids, int64_t q) {
  if (prob == nullptr) return 1.f;
  return (float)prob[eids ? (int64_t)eids[q] : q];
}

template <typename I, typename W>
__global__ void __launch_bounds__(256) solve_c_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                      const I* __restrict__ eids, const I* __restrict__ seeds, int64_t n,
                                                      int fanout, const float* __restrict__ pi, const W* __restrict__ prob,
                                                      float* __restrict__ c_out) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  const int lane = threadIdx.x % kGroup;
  if (i >= n) return;
  const unsigned mask = ((1u << kGroup) - 1) << ((threadIdx.x & 31) & ~(kGroup - 1));
  const I s = seeds[i];
  const int64_t b = indptr[s], e = indptr[s + 1];
  float sum_w = 0.f, live = 0.f;
  for (int64_t q = b + lane; q < e; q += kGroup) {
    const float w = pi[indices[q]] * edge_weight<I, W>(prob, eids, q);
    sum_w += w;
    live += w > 0.f ? 1.f : 0.f;
  }
  sum_w = group_sum(sum_w, mask);
  live = group_sum(live, mask);
  float c = 3.0e38f;                              // "keep every eligible edge"
  if (fanout >= 0 && live > (float)fanout && sum_w > 0.f) {
    const float k = (float)fanout;
    c = k / sum_w;
    for (int it = 0; it < 64; ++it) {
      float sat = 0.f, rest = 0.f;
      for (int64_t q = b + lane; q < e; q += kGroup) {
        const float w = pi[indices[q]] * edge_weight<I, W>(prob, eids, q);
        if (c * w >= 1.f) sat += 1.f; else rest += w;
      }
      sat = group_sum(sat, mask);
      rest = group_sum(rest, mask);
      if (rest <= 0.f || sat >= k) break;
      const float next = (k - sat) / rest;
      if (next <= c * 1.000001f) break;
      c = next;
    }
  }
  if (lane == 0) c_out[i] = c;
}

// mode 0: pi_next[t] = max(pi_next[t], min(c_s * pi_t * a, big));  1: pi[t] = max(pi_next[t], tiny);  2: pi_next[t] = 0;  3: pi[t] = 1
template <typename I, typename W>
__global__ void __launch_bounds__(256) pi_pass_kernel(int mode, const I* __restrict__ indptr, const I* __restrict__ indices,
                                                      const I* __restrict__ eids, const I* __restrict__ seeds, int64_t n,
                                                      const float* __restrict__ c, float* __restrict__ pi,
                                                      float* __restrict__ pi_next, const W* __restrict__ prob) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  const int lane = threadIdx.x % kGroup;
  if (i >= n) return;
  const I s = seeds[i];
  const int64_t b = indptr[s], e = indptr[s + 1];
  const float cs = mode == 0 ? c[i] : 0.f;
  for (int64_t q = b + lane; q < e; q += kGroup) {
    const int64_t t = indices[q];
    if (mode == 0) {
      const float cand = fminf(cs, 1.0e30f) * pi[t] * edge_weight<I, W>(prob, eids, q);
      atomicMax(reinterpret_cast<int*>(pi_next + t), __float_as_int(fminf(cand, 1.0e30f)));
    } else if (mode == 1) {
      pi[t] = fmaxf(pi_next[t], 1.0e-12f);
    } else if (mode == 2) {
      pi_next[t] = 0.f;
    } else {
      pi[t] = 1.f;
    }
  }
}

template <typename I, typename W, bool FILL>
__global__ void __launch_bounds__(256) pick_kernel(const I* __restrict__ indptr, const I* __restrict__ indices,
                                                   const I* __restrict__ eids, const I* __restrict__ seeds, int64_t n,
                                                   const float* __restrict__ c, const float* __restrict__ pi,
                                                   const W* __restrict__ prob, uint64_t seed1, uint64_t seed2, float mix,
                                                   int64_t* __restrict__ counts, float* __restrict__ wsum,
                                                   const int64_t* __restrict__ offsets, I* __restrict__ out_src,
                                                   I* __restrict__ out_seedpos, I* __restrict__ out_eid,
                                                   float* __restrict__ out_w) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  const int lane = threadIdx.x % kGroup;
  if (i >= n) return;
  const int sh = (threadIdx.x & 31) & ~(kGroup - 1);
  const unsigned mask = ((1u << kGroup) - 1) << sh;
  const I s = seeds[i];
  const int64_t b = indptr[s], e = indptr[s + 1];
  const float cs = c[i];
  const float norm = FILL ? 1.f / fmaxf(wsum[i], 1.0e-20f) : 0.f;
  int64_t base = FILL ? offsets[i] : 0, cnt = 0;
  float acc = 0.f;
  for (int64_t q0 = b; q0 < e; q0 += kGroup) {
    const int64_t q = q0 + lane;
    bool keep = false;
    I t = 0;
    float p = 0.f;
    if (q < e) {
      t = indices[q];
      const float w = pi[t] * edge_weight<I, W>(prob, eids, q);
      p = w > 0.f ? fminf(1.f, fminf(cs, 1.0e30f) * w) : 0.f;
      keep = p > 0.f && variate(seed1, seed2, mix, (uint64_t)t) <= p;
    }
    const unsigned m = __ballot_sync(mask, keep) & mask;
    if (keep) {
      if (FILL) {
        const int64_t pos = base + cnt + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
        out_src[pos] = t;
        out_seedpos[pos] = (I)i;
        out_eid[pos] = eids ? eids[q] : (I)q;
        out_w[pos] = norm / p;
      } else {
        acc += 1.f / p;
      }
    }
    cnt += __popc(m);
  }
  if (!FILL) {
    acc = group_sum(acc, mask);
    if (lane == 0) {
      counts[i] = cnt;
      wsum[i] = acc;
    }
  }
}

}  // namespace labor
}  // namespace dglb

using namespace dglb;

#define LABOR_PROB_SWITCH(code, W, ...)                                             \
  switch (code) {                                                                   \
    case -1: case kF32: { using W = float; __VA_ARGS__; break; }                    \
    case kF64: { using W = double; __VA_ARGS__; break; }                            \
    case kF16: { using W = __half; __VA_ARGS__; break; }                            \
    case kBF16: { using W = __nv_bfloat16; __VA_ARGS__; break; }                    \
    default: throw std::runtime_error("labor: unsupported edge weight dtype");      \
  }

static unsigned labor_blocks(int64_t n) { return (unsigned)((n * labor::kGroup + 255) / 256); }

// c[i] for every seed given the current pi (prob_dtype = -1: no edge weights)
DGLB_API int dglb200_cuda_labor_solve_c(int idtype, const void* indptr, const void* indices, const void* eids,
                                        const void* seeds, int64_t n, int fanout, const void* pi, int prob_dtype,
                                        const void* prob, void* c, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    LABOR_PROB_SWITCH(prob_dtype, W, {
      labor::solve_c_kernel<I, W><<<labor_blocks(n), 256, 0, (cudaStream_t)stream>>>(
          (const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, fanout, (const float*)pi, (const W*)prob, (float*)c);
    });
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

DGLB_API int dglb200_cuda_labor_pi_pass(int idtype, int mode, const void* indptr, const void* indices, const void* eids,
                                        const void* seeds, int64_t n, const void* c, void* pi, void* pi_next, int prob_dtype,
                                        const void* prob, void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    LABOR_PROB_SWITCH(prob_dtype, W, {
      labor::pi_pass_kernel<I, W><<<labor_blocks(n), 256, 0, (cudaStream_t)stream>>>(
          mode, (const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, (const float*)c, (float*)pi,
          (float*)pi_next, (const W*)prob);
    });
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}

// pass 0: counts[n] (int64) + wsum[n]; pass 1: offsets = exclusive scan of counts, fills src / seed position / eid / weight
DGLB_API int dglb200_cuda_labor_pick(int idtype, int pass, const void* indptr, const void* indices, const void* eids,
                                     const void* seeds, int64_t n, const void* c, const void* pi, int prob_dtype,
                                     const void* prob, int64_t seed1, int64_t seed2, float mix, void* counts, void* wsum,
                                     const void* offsets, void* out_src, void* out_seedpos, void* out_eid, void* out_w,
                                     void* stream) {
  DGLB_TRY
  if (n == 0) return 0;
  DGLB_ID_SWITCH(idtype, I, {
    LABOR_PROB_SWITCH(prob_dtype, W, {
      if (pass == 0)
        labor::pick_kernel<I, W, false><<<labor_blocks(n), 256, 0, (cudaStream_t)stream>>>(
            (const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, (const float*)c, (const float*)pi,
            (const W*)prob, (uint64_t)seed1, (uint64_t)seed2, mix, (int64_t*)counts, (float*)wsum, nullptr, nullptr, nullptr,
            nullptr, nullptr);
      else
        labor::pick_kernel<I, W, true><<<labor_blocks(n), 256, 0, (cudaStream_t)stream>>>(
            (const I*)indptr, (const I*)indices, (const I*)eids, (const I*)seeds, n, (const float*)c, (const float*)pi,
            (const W*)prob, (uint64_t)seed1, (uint64_t)seed2, mix, nullptr, (float*)wsum, (const int64_t*)offsets, (I*)out_src,
            (I*)out_seedpos, (I*)out_eid, (float*)out_w);
    });
  });
  DGLB_LAUNCH_CHECK();
  DGLB_CATCH
}
