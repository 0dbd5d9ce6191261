// dgl.random.choice: draw `num` indices from [0, population) uniformly or by weight, with or
// without replacement.
//
// Behavioural reference: src/random/cpu/choice.cc + python/dgl/random.py:21.  Strategy per case:
//   uniform, replace      : independent draws, parallel counter-based streams
//   uniform, no replace   : rejection into an open-addressing set when the draw is sparse (< 1/16 of
//                           the population), partial Fisher-Yates otherwise
//   weighted, replace     : prefix sums + binary search per draw
//   weighted, no replace  : Efraimidis-Spirakis keys (log u / w), keep the `num` largest
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "common.h"

namespace dglb {
namespace {

struct Stream {
  uint64_t s;
  Stream(uint64_t seed, uint64_t lane) : s(seed * 0x9E3779B97F4A7C15ull + (lane + 1) * 0xD1B54A32D192ED03ull) { next(); }
  inline uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
  inline double unit() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }  // (0, 1)
};

// Draws are generated in fixed blocks of kBlock outputs, block j always using stream j: the result is a pure function of
// (seed, num) and does not depend on how many threads happen to run the blocks.
constexpr int64_t kBlock = 4096;

template <typename F>
void for_each_block(int64_t num, F&& body) {
  const int64_t nblocks = (num + kBlock - 1) / kBlock;
  parallel_for_dynamic(0, nblocks, 4, [&](int64_t b, int64_t e) {
    for (int64_t j = b; j < e; ++j) body(j, j * kBlock, std::min(num, (j + 1) * kBlock));
  });
}

void uniform_replace(int64_t num, int64_t population, uint64_t seed, int64_t* out) {
  for_each_block(num, [&](int64_t block, int64_t lo, int64_t hi) {
    Stream rng(seed, (uint64_t)block);
    for (int64_t i = lo; i < hi; ++i) out[i] = (int64_t)rng.below((uint64_t)population);
  });
}

void uniform_no_replace(int64_t num, int64_t population, uint64_t seed, int64_t* out) {
  DGLB_CHECK(num <= population, "Cannot take more samples than the population when 'replace=False'");
  Stream rng(seed, 0);
  if (num * 16 < population) {
    int64_t cap = 16;
    while (cap < num * 2) cap <<= 1;
    std::vector<int64_t> table(cap, -1);
    for (int64_t i = 0; i < num;) {
      const int64_t v = (int64_t)rng.below((uint64_t)population);
      uint64_t h = ((uint64_t)v * 0x9E3779B97F4A7C15ull) >> 20 & (uint64_t)(cap - 1);
      while (table[h] >= 0 && table[h] != v) h = (h + 1) & (uint64_t)(cap - 1);
      if (table[h] == v) continue;
      table[h] = v;
      out[i++] = v;
    }
    return;
  }
  std::vector<int64_t> pool(population);
  std::iota(pool.begin(), pool.end(), int64_t(0));
  for (int64_t i = 0; i < num; ++i) {
    const int64_t j = i + (int64_t)rng.below((uint64_t)(population - i));
    std::swap(pool[i], pool[j]);
    out[i] = pool[i];
  }
}

template <typename T>
void weighted_replace(int64_t num, int64_t population, const T* prob, uint64_t seed, int64_t* out) {
  std::vector<double> cdf(population);
  double acc = 0;
  for (int64_t i = 0; i < population; ++i) {
    const double w = (double)(float)prob[i];
    DGLB_CHECK(w >= 0, "probabilities must be non-negative");
    acc += w;
    cdf[i] = acc;
  }
  DGLB_CHECK(acc > 0, "probabilities must not all be zero");
  for_each_block(num, [&](int64_t block, int64_t b, int64_t e) {
    Stream rng(seed, (uint64_t)block);
    for (int64_t i = b; i < e; ++i) {
      const double u = rng.unit() * acc;
      int64_t p = std::upper_bound(cdf.begin(), cdf.end(), u) - cdf.begin();
      if (p >= population) p = population - 1;
      // never return a zero-weight entry because of rounding at a plateau edge
      while (p > 0 && (double)(float)prob[p] == 0) --p;
      out[i] = p;
    }
  });
}

template <typename T>
void weighted_no_replace(int64_t num, int64_t population, const T* prob, uint64_t seed, int64_t* out) {
  std::vector<std::pair<double, int64_t>> keys;
  keys.reserve(population);
  Stream rng(seed, 0);
  for (int64_t i = 0; i < population; ++i) {
    const double w = (double)(float)prob[i];
    DGLB_CHECK(w >= 0, "probabilities must be non-negative");
    const double u = rng.unit();  // drawn for every entry so results do not depend on which weights are zero
    if (w > 0) keys.emplace_back(std::log(u) / w, i);
  }
  DGLB_CHECK((int64_t)keys.size() >= num,
             "Cannot take more samples than entries with non-zero probability when 'replace=False'");
  std::nth_element(keys.begin(), keys.begin() + num, keys.end(),
                   [](const std::pair<double, int64_t>& a, const std::pair<double, int64_t>& b) { return a.first > b.first; });
  std::sort(keys.begin(), keys.begin() + num,
            [](const std::pair<double, int64_t>& a, const std::pair<double, int64_t>& b) { return a.first > b.first; });
  for (int64_t i = 0; i < num; ++i) out[i] = keys[i].second;
}

}  // namespace
}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cpu_choice(int64_t num, int64_t population, int dtype, const void* prob, int replace,
                                int64_t seed, int64_t* out) {
  DGLB_TRY
  DGLB_CHECK(population > 0 || num == 0, "cannot sample from an empty population");
  if (num == 0) return 0;
  if (!prob) {
    if (replace) {
      uniform_replace(num, population, (uint64_t)seed, out);
    } else {
      uniform_no_replace(num, population, (uint64_t)seed, out);
    }
  } else if (dtype == kF64) {
    if (replace) {
      weighted_replace<double>(num, population, (const double*)prob, (uint64_t)seed, out);
    } else {
      weighted_no_replace<double>(num, population, (const double*)prob, (uint64_t)seed, out);
    }
  } else {
    DGLB_DTYPE_SWITCH(dtype, T, {
      if (replace) {
        weighted_replace<T>(num, population, (const T*)prob, (uint64_t)seed, out);
      } else {
        weighted_no_replace<T>(num, population, (const T*)prob, (uint64_t)seed, out);
      }
    });
  }
  DGLB_CATCH
}
