// CPU cache-replacement policies for feature caches: S3-FIFO, SIEVE, LRU, CLOCK, plus a
// partitioned (sharded, multi-threaded) wrapper.
// Behavioural spec: /root/reference/graphbolt/src/cache_policy.{h,cc} (BaseCachePolicy :143,
// S3FifoCachePolicy :256, SieveCachePolicy :389, LruCachePolicy :485, ClockCachePolicy :579) and
// partitioned_cache_policy.cc. Fresh implementation: every policy manages `capacity` slots addressed
// by a dense slot id (the row of the cache tensor); keys map to slots through one open-addressing
// table per shard; batch query/replace calls are sharded by key hash and run under OpenMP.
#include <deque>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "common.h"

namespace dglb {

struct SlotMap {  // key -> slot, open addressing with tombstone-free backward-shift deletion
  std::vector<int64_t> keys;
  std::vector<int32_t> vals;
  uint64_t mask;
  explicit SlotMap(int64_t cap) {
    uint64_t c = 16;
    while ((int64_t)c < 2 * cap + 2) c <<= 1;
    keys.assign(c, -1);
    vals.assign(c, -1);
    mask = c - 1;
  }
  static inline uint64_t h(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    return k;
  }
  int32_t find(int64_t k) const {
    uint64_t i = h((uint64_t)k) & mask;
    while (keys[i] != -1) {
      if (keys[i] == k) return vals[i];
      i = (i + 1) & mask;
    }
    return -1;
  }
  void put(int64_t k, int32_t v) {
    uint64_t i = h((uint64_t)k) & mask;
    while (keys[i] != -1 && keys[i] != k) i = (i + 1) & mask;
    keys[i] = k;
    vals[i] = v;
  }
  void erase(int64_t k) {
    uint64_t i = h((uint64_t)k) & mask;
    while (keys[i] != -1 && keys[i] != k) i = (i + 1) & mask;
    if (keys[i] == -1) return;
    uint64_t j = i;
    while (true) {
      j = (j + 1) & mask;
      if (keys[j] == -1) break;
      uint64_t home = h((uint64_t)keys[j]) & mask;
      // can entry j move to hole i?
      if ((i <= j) ? (home <= i || home > j) : (home <= i && home > j)) {
        keys[i] = keys[j];
        vals[i] = vals[j];
        i = j;
      }
    }
    keys[i] = -1;
    vals[i] = -1;
  }
};

class Policy {
 public:
  explicit Policy(int64_t cap) : cap_(cap), map_(cap), slot_key_(cap, -1) {}
  virtual ~Policy() = default;
  // returns slot or -1; updates recency/frequency metadata on hit
  virtual int32_t Query(int64_t key) = 0;
  // ensures key is cached; returns its slot (evicting if needed)
  virtual int32_t Insert(int64_t key) = 0;
  int64_t capacity() const { return cap_; }

 protected:
  int64_t cap_;
  SlotMap map_;
  std::vector<int64_t> slot_key_;
};

class LruPolicy : public Policy {
 public:
  explicit LruPolicy(int64_t cap) : Policy(cap), prev_(cap, -1), next_(cap, -1), used_(0) { head_ = tail_ = -1; }
  int32_t Query(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0) Touch(s);
    return s;
  }
  int32_t Insert(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0) { Touch(s); return s; }
    if (used_ < cap_) {
      s = (int32_t)used_++;
    } else {
      s = tail_;
      Unlink(s);
      map_.erase(slot_key_[s]);
    }
    slot_key_[s] = key;
    map_.put(key, s);
    PushFront(s);
    return s;
  }

 private:
  void Unlink(int32_t s) {
    if (prev_[s] >= 0) next_[prev_[s]] = next_[s]; else head_ = next_[s];
    if (next_[s] >= 0) prev_[next_[s]] = prev_[s]; else tail_ = prev_[s];
    prev_[s] = next_[s] = -1;
  }
  void PushFront(int32_t s) {
    prev_[s] = -1;
    next_[s] = head_;
    if (head_ >= 0) prev_[head_] = s;
    head_ = s;
    if (tail_ < 0) tail_ = s;
  }
  void Touch(int32_t s) {
    if (head_ == s) return;
    Unlink(s);
    PushFront(s);
  }
  std::vector<int32_t> prev_, next_;
  int32_t head_, tail_;
  int64_t used_;
};

class ClockPolicy : public Policy {
 public:
  explicit ClockPolicy(int64_t cap) : Policy(cap), ref_(cap, 0), hand_(0), used_(0) {}
  int32_t Query(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0) ref_[s] = 1;
    return s;
  }
  int32_t Insert(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0) { ref_[s] = 1; return s; }
    if (used_ < cap_) {
      s = (int32_t)used_++;
    } else {
      while (ref_[hand_]) {
        ref_[hand_] = 0;
        hand_ = (hand_ + 1) % cap_;
      }
      s = (int32_t)hand_;
      hand_ = (hand_ + 1) % cap_;
      map_.erase(slot_key_[s]);
    }
    slot_key_[s] = key;
    ref_[s] = 0;
    map_.put(key, s);
    return s;
  }

 private:
  std::vector<uint8_t> ref_;
  int64_t hand_, used_;
};

// SIEVE (Zhang et al., NSDI'24): FIFO order + visited bit + a hand moving from tail to head.
class SievePolicy : public Policy {
 public:
  explicit SievePolicy(int64_t cap) : Policy(cap), prev_(cap, -1), next_(cap, -1), visited_(cap, 0), used_(0) {
    head_ = tail_ = hand_ = -1;
  }
  int32_t Query(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0) visited_[s] = 1;
    return s;
  }
  int32_t Insert(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0) { visited_[s] = 1; return s; }
    if (used_ < cap_) {
      s = (int32_t)used_++;
    } else {
      int32_t o = hand_ >= 0 ? hand_ : tail_;
      while (visited_[o]) {
        visited_[o] = 0;
        o = prev_[o] >= 0 ? prev_[o] : tail_;
      }
      hand_ = prev_[o];  // may be -1 -> restart from tail next time
      // unlink o
      if (prev_[o] >= 0) next_[prev_[o]] = next_[o]; else head_ = next_[o];
      if (next_[o] >= 0) prev_[next_[o]] = prev_[o]; else tail_ = prev_[o];
      map_.erase(slot_key_[o]);
      s = o;
    }
    slot_key_[s] = key;
    visited_[s] = 0;
    map_.put(key, s);
    prev_[s] = -1;
    next_[s] = head_;
    if (head_ >= 0) prev_[head_] = s;
    head_ = s;
    if (tail_ < 0) tail_ = s;
    return s;
  }

 private:
  std::vector<int32_t> prev_, next_;
  std::vector<uint8_t> visited_;
  int32_t head_, tail_, hand_;
  int64_t used_;
};

// S3-FIFO (Yang et al., SOSP'23): small FIFO (10%), main FIFO (90%), ghost FIFO of recently
// evicted keys; 2-bit frequency counters.
class S3FifoPolicy : public Policy {
 public:
  explicit S3FifoPolicy(int64_t cap)
      : Policy(cap), freq_(cap, 0), where_(cap, 0), small_cap_(std::max<int64_t>(1, cap / 10)), ghost_(cap) {
    for (int64_t i = cap - 1; i >= 0; --i) free_.push_back((int32_t)i);
    ghost_cap_ = cap;
  }
  int32_t Query(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0 && freq_[s] < 3) ++freq_[s];
    return s;
  }
  int32_t Insert(int64_t key) override {
    int32_t s = map_.find(key);
    if (s >= 0) {
      if (freq_[s] < 3) ++freq_[s];
      return s;
    }
    const bool in_ghost = ghost_.find(key) >= 0;
    while (free_.empty()) Evict();
    s = free_.back();
    free_.pop_back();
    slot_key_[s] = key;
    freq_[s] = 0;
    map_.put(key, s);
    if (in_ghost) {
      where_[s] = 1;
      main_.push_back(s);
      ghost_.erase(key);
    } else {
      where_[s] = 0;
      small_.push_back(s);
    }
    return s;
  }

 private:
  void Evict() {
    if ((int64_t)small_.size() >= small_cap_ || main_.empty()) EvictSmall(); else EvictMain();
  }
  void EvictSmall() {
    while (!small_.empty()) {
      int32_t s = small_.front();
      small_.pop_front();
      if (freq_[s] > 1) {  // promoted
        where_[s] = 1;
        freq_[s] = 0;
        main_.push_back(s);
        if ((int64_t)main_.size() > cap_ - small_cap_) { EvictMain(); return; }
      } else {
        GhostInsert(slot_key_[s]);
        map_.erase(slot_key_[s]);
        slot_key_[s] = -1;
        free_.push_back(s);
        return;
      }
    }
    if (free_.empty()) EvictMain();
  }
  void EvictMain() {
    while (!main_.empty()) {
      int32_t s = main_.front();
      main_.pop_front();
      if (freq_[s] > 0) {
        --freq_[s];
        main_.push_back(s);
      } else {
        map_.erase(slot_key_[s]);
        slot_key_[s] = -1;
        free_.push_back(s);
        return;
      }
    }
  }
  void GhostInsert(int64_t key) {
    ghost_q_.push_back(key);
    ghost_.put(key, 1);
    while ((int64_t)ghost_q_.size() > ghost_cap_) {
      ghost_.erase(ghost_q_.front());
      ghost_q_.pop_front();
    }
  }
  std::vector<uint8_t> freq_, where_;
  int64_t small_cap_, ghost_cap_;
  SlotMap ghost_;
  std::deque<int32_t> small_, main_;
  std::deque<int64_t> ghost_q_;
  std::vector<int32_t> free_;
};

struct PartitionedCache {
  std::vector<std::unique_ptr<Policy>> parts;
  std::vector<int64_t> offsets;  // slot offset of each shard in the global cache tensor
  int64_t hits = 0, queries = 0;
  std::mutex mu;
};

static std::unique_ptr<Policy> make_policy(int kind, int64_t cap) {
  switch (kind) {
    case 0: return std::make_unique<S3FifoPolicy>(cap);
    case 1: return std::make_unique<SievePolicy>(cap);
    case 2: return std::make_unique<LruPolicy>(cap);
    case 3: return std::make_unique<ClockPolicy>(cap);
    default: throw std::runtime_error("unknown cache policy");
  }
}

static inline int shard_of(int64_t key, int nparts) {
  return (int)((SlotMap::h((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 17) % (uint64_t)nparts);
}

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cache_create(int policy, int64_t capacity, int num_parts, int64_t* handle) {
  DGLB_TRY
  DGLB_CHECK(capacity > 0 && num_parts > 0, "capacity and num_parts must be positive");
  auto* c = new PartitionedCache();
  int64_t off = 0;
  // never more shards than slots: every shard owns at least one slot and the slot ranges tile [0, capacity)
  if ((int64_t)num_parts > capacity) num_parts = (int)capacity;
  for (int p = 0; p < num_parts; ++p) {
    int64_t cap = capacity / num_parts + (p < capacity % num_parts ? 1 : 0);
    c->parts.push_back(make_policy(policy, cap));
    c->offsets.push_back(off);
    off += cap;
  }
  *handle = (int64_t)(intptr_t)c;
  DGLB_CATCH
}

DGLB_API int dglb200_cache_destroy(int64_t handle) {
  DGLB_TRY
  delete (PartitionedCache*)(intptr_t)handle;
  DGLB_CATCH
}

// positions[i] = global slot of keys[i] or -1. mode 0 = query (read-only w.r.t. membership),
// mode 1 = replace (insert missing keys, evicting per policy).
static int cache_op(int64_t handle, const int64_t* keys, int64_t n, int64_t* positions, int mode, int64_t* num_hit) {
  auto* c = (PartitionedCache*)(intptr_t)handle;
  const int P = (int)c->parts.size();
  std::vector<std::vector<int64_t>> buckets(P);
  for (int64_t i = 0; i < n; ++i) buckets[shard_of(keys[i], P)].push_back(i);
  int64_t hits = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : hits) if (P > 1 && n > 4096)
  for (int p = 0; p < P; ++p) {
    Policy* pol = c->parts[p].get();
    const int64_t off = c->offsets[p];
    for (int64_t i : buckets[p]) {
      int32_t s = mode == 0 ? pol->Query(keys[i]) : pol->Insert(keys[i]);
      positions[i] = s >= 0 ? off + s : -1;
      hits += s >= 0;
    }
  }
  if (mode == 0) {
    std::lock_guard<std::mutex> lk(c->mu);
    c->hits += hits;
    c->queries += n;
  }
  if (num_hit) *num_hit = hits;
  return 0;
}

DGLB_API int dglb200_cache_query(int64_t handle, const void* keys, int64_t n, void* positions, void* num_hit) {
  DGLB_TRY
  cache_op(handle, (const int64_t*)keys, n, (int64_t*)positions, 0, (int64_t*)num_hit);
  DGLB_CATCH
}

DGLB_API int dglb200_cache_replace(int64_t handle, const void* keys, int64_t n, void* positions) {
  DGLB_TRY
  cache_op(handle, (const int64_t*)keys, n, (int64_t*)positions, 1, nullptr);
  DGLB_CATCH
}

DGLB_API int dglb200_cache_stats(int64_t handle, int64_t* hits, int64_t* queries) {
  DGLB_TRY
  auto* c = (PartitionedCache*)(intptr_t)handle;
  *hits = c->hits;
  *queries = c->queries;
  DGLB_CATCH
}

// multi-threaded row gather / scatter between a cache tensor and user buffers (FeatureCache)
DGLB_API int dglb200_cpu_gather_rows(const void* src, const void* idx, int64_t n, void* dst, int64_t row_bytes) {
  DGLB_TRY
  const int64_t* ix = (const int64_t*)idx;
  parallel_for(0, n, 256, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
      if (ix[i] >= 0) std::memcpy((char*)dst + i * row_bytes, (const char*)src + ix[i] * row_bytes, row_bytes);
  });
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_scatter_rows(void* dst, const void* idx, int64_t n, const void* src, int64_t row_bytes) {
  DGLB_TRY
  const int64_t* ix = (const int64_t*)idx;
  parallel_for(0, n, 256, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i)
      if (ix[i] >= 0) std::memcpy((char*)dst + ix[i] * row_bytes, (const char*)src + i * row_bytes, row_bytes);
  });
  DGLB_CATCH
}
