// Shared device/host helpers for the sm_100a kernel library.
// Conventions: every exported entry point is `extern "C" int f(..., void* stream)`,
// enqueues on the given stream (PyTorch's current stream), never synchronises and never
// allocates, so all of them are CUDA-graph capturable.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#define DGLB_API extern "C" __attribute__((visibility("default")))

namespace dglb {

void set_last_error(const std::string& msg);

enum DType : int { kF32 = 0, kF64 = 1, kF16 = 2, kBF16 = 3 };
enum IdType : int { kI32 = 0, kI64 = 1 };
enum BinOp : int { kAdd = 0, kSub = 1, kMul = 2, kDiv = 3, kCopyLhs = 4, kCopyRhs = 5, kDot = 6 };
enum RedOp : int { kSum = 0, kMax = 1, kMin = 2 };

constexpr int kNumSMs = 148;  // B200

struct Bcast {
  int64_t ldiv, lmod, rdiv, rmod;
};

template <typename T>
struct AccT {
  using type = float;
};
template <>
struct AccT<double> {
  using type = double;
};

template <typename T>
__device__ __forceinline__ float to_f(T v) {
  return static_cast<float>(v);
}
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <>
__device__ __forceinline__ float to_f<__half>(__half v) {
  return __half2float(v);
}
template <typename T>
__device__ __forceinline__ T from_f(float v) {
  return static_cast<T>(v);
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ __half from_f<__half>(float v) {
  return __float2half_rn(v);
}

template <typename T, typename A>
__device__ __forceinline__ A ld_acc(const T* p) {
  return static_cast<A>(to_f<T>(*p));
}
template <>
__device__ __forceinline__ double ld_acc<double, double>(const double* p) {
  return *p;
}
template <typename T, typename A>
__device__ __forceinline__ T st_acc(A v) {
  return from_f<T>(static_cast<float>(v));
}
template <>
__device__ __forceinline__ double st_acc<double, double>(double v) {
  return v;
}

template <typename A, int OP>
__device__ __forceinline__ A apply_op(A l, A r) {
  if (OP == kAdd) return l + r;
  if (OP == kSub) return l - r;
  if (OP == kMul) return l * r;
  if (OP == kDiv) return l / r;
  if (OP == kCopyLhs) return l;
  if (OP == kCopyRhs) return r;
  return l * r;
}

// counter-based RNG: splitmix64 over (seed, stream, counter)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t stream, uint64_t ctr) {
  return mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (stream + 1)) + 0xD1B54A32D192ED03ull * (ctr + 1));
}
__device__ __forceinline__ uint32_t rng_below(uint64_t seed, uint64_t stream, uint64_t ctr, uint32_t n) {
  return (uint32_t)(((rng_u64(seed, stream, ctr) >> 32) * (uint64_t)n) >> 32);
}
__device__ __forceinline__ float rng_uniform(uint64_t seed, uint64_t stream, uint64_t ctr) {
  return ((rng_u64(seed, stream, ctr) >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace dglb

#define DGLB_TRY try {
#define DGLB_CATCH                                   \
  }                                                  \
  catch (const std::exception& e) {                  \
    dglb::set_last_error(e.what());                  \
    return 1;                                        \
  }                                                  \
  catch (...) {                                      \
    dglb::set_last_error("unknown C++ exception");   \
    return 1;                                        \
  }                                                  \
  return 0;

#define DGLB_CHECK(cond, msg)                                                                \
  do {                                                                                       \
    if (!(cond)) throw std::runtime_error(std::string("check failed: " #cond " ") + (msg)); \
  } while (0)

#define DGLB_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));      \
  } while (0)

#define DGLB_LAUNCH_CHECK() DGLB_CUDA_CHECK(cudaGetLastError())

#define DGLB_FLOAT_SWITCH(code, T, ...)                                   \
  switch (code) {                                                         \
    case dglb::kF32: { using T = float; __VA_ARGS__; break; }             \
    case dglb::kF16: { using T = __half; __VA_ARGS__; break; }            \
    case dglb::kBF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }    \
    case dglb::kF64: { using T = double; __VA_ARGS__; break; }            \
    default: throw std::runtime_error("unsupported dtype code");          \
  }

#define DGLB_ID_SWITCH(code, I, ...)                                      \
  switch (code) {                                                         \
    case dglb::kI32: { using I = int32_t; __VA_ARGS__; break; }           \
    case dglb::kI64: { using I = int64_t; __VA_ARGS__; break; }           \
    default: throw std::runtime_error("unsupported idtype code");         \
  }
