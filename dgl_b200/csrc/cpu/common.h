// Host-runtime common helpers: error channel, dtype codes, 16-bit float conversion,
// OpenMP parallel_for. Counterpart of the reference's runtime layer
// (/root/reference/include/dgl/runtime/parallel_for.h, src/runtime/*), reduced to
// what a torch-allocated, ctypes-bound design needs.
#pragma once
#include <omp.h>

#include <atomic>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#define DGLB_API extern "C" __attribute__((visibility("default")))

namespace dglb {

void set_last_error(const std::string& msg);

enum DType : int { kF32 = 0, kF64 = 1, kF16 = 2, kBF16 = 3 };
enum IdType : int { kI32 = 0, kI64 = 1 };
enum BinOp : int { kAdd = 0, kSub = 1, kMul = 2, kDiv = 3, kCopyLhs = 4, kCopyRhs = 5, kDot = 6 };
enum RedOp : int { kSum = 0, kMax = 1, kMin = 2 };

struct bf16_t {
  uint16_t v;
  bf16_t() = default;
  bf16_t(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) {
      v = uint16_t((u >> 16) | 0x40);
    } else {
      u += 0x7fffu + ((u >> 16) & 1u);
      v = uint16_t(u >> 16);
    }
  }
  operator float() const {
    uint32_t u = uint32_t(v) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }
};

struct f16_t {
  _Float16 v;
  f16_t() = default;
  f16_t(float f) : v((_Float16)f) {}
  operator float() const { return (float)v; }
};

template <typename T>
struct AccT {
  using type = float;
};
template <>
struct AccT<double> {
  using type = double;
};

template <typename F>
inline void parallel_for(int64_t begin, int64_t end, int64_t grain, F&& f) {
  const int64_t n = end - begin;
  if (n <= 0) return;
  if (n <= grain || omp_get_max_threads() == 1 || omp_in_parallel()) {
    f(begin, end);
    return;
  }
#pragma omp parallel
  {
    const int64_t nt = omp_get_num_threads();
    const int64_t tid = omp_get_thread_num();
    const int64_t chunk = (n + nt - 1) / nt;
    const int64_t b = begin + tid * chunk;
    const int64_t e = std::min(end, b + chunk);
    if (b < e) f(b, e);
  }
}

// dynamic-chunk variant for skewed rows
template <typename F>
inline void parallel_for_dynamic(int64_t begin, int64_t end, int64_t chunk, F&& f) {
  const int64_t n = end - begin;
  if (n <= 0) return;
  if (n <= chunk || omp_get_max_threads() == 1 || omp_in_parallel()) {
    f(begin, end);
    return;
  }
  const int64_t nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t b = begin + c * chunk;
    f(b, std::min(end, b + chunk));
  }
}

}  // namespace dglb

// Process-wide OpenMP team size. omp_set_num_threads() only changes the CALLING thread's setting, and threads that
// were not created by OpenMP (Python's I/O pool, DataLoader threads) start from the environment default - the visible
// core count, not the cgroup CPU budget the loader computed. Every exported entry point therefore re-applies the global
// value to its calling thread first (a relaxed load and a compare when nothing changed).
namespace dglb {
inline std::atomic<int>& team_size_global() {
  static std::atomic<int> v{0};
  return v;
}
// per-thread override (> 0): threads of a pool that runs several native calls side by side take a share of the budget
inline int& team_size_thread() {
  static thread_local int v = 0;
  return v;
}
inline void sync_team() {
  int t = team_size_thread();
  if (t <= 0) t = team_size_global().load(std::memory_order_relaxed);
  if (t > 0 && omp_get_max_threads() != t) omp_set_num_threads(t);
}
}  // namespace dglb

#define DGLB_TRY try { dglb::sync_team();
#define DGLB_CATCH                               \
  }                                              \
  catch (const std::exception& e) {              \
    dglb::set_last_error(e.what());              \
    return 1;                                    \
  }                                              \
  catch (...) {                                  \
    dglb::set_last_error("unknown C++ exception"); \
    return 1;                                    \
  }                                              \
  return 0;

#define DGLB_CHECK(cond, msg)                                                       \
  do {                                                                              \
    if (!(cond)) throw std::runtime_error(std::string("check failed: " #cond " ") + (msg)); \
  } while (0)

#define DGLB_DTYPE_SWITCH(code, T, ...)                         \
  switch (code) {                                               \
    case dglb::kF32: { using T = float; __VA_ARGS__; break; }   \
    case dglb::kF64: { using T = double; __VA_ARGS__; break; }  \
    case dglb::kF16: { using T = dglb::f16_t; __VA_ARGS__; break; } \
    case dglb::kBF16: { using T = dglb::bf16_t; __VA_ARGS__; break; } \
    default: throw std::runtime_error("unsupported dtype code"); \
  }

#define DGLB_ID_SWITCH(code, I, ...)                              \
  switch (code) {                                                 \
    case dglb::kI32: { using I = int32_t; __VA_ARGS__; break; }   \
    case dglb::kI64: { using I = int64_t; __VA_ARGS__; break; }   \
    default: throw std::runtime_error("unsupported idtype code"); \
  }
