// Host geometry: kd-tree k-nearest-neighbour search (segmented), farthest-point sampling and
// greedy neighbour matching for graph coarsening.
//
// Behavioural reference: src/graph/transform/cpu/knn.cc (nanoflann kd-tree path),
// src/geometry/cpu/geometry_op_impl.cc (FarthestPointSampler, NeighborMatching).  The tree here
// is our own: implicit balanced tree over a permuted index array (median split on the widest
// dimension, 16-point leaves), queries run in parallel with a bounded max-heap per query.
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

#include "common.h"

namespace dglb {
namespace {

template <typename T>
struct KdTree {
  struct Node {
    int64_t lo, hi;   // range in `perm`
    int32_t left, right;  // children (-1 for leaves)
    int32_t dim;
    T split;
  };
  const T* pts;
  int64_t dim;
  std::vector<int64_t> perm;
  std::vector<Node> nodes;
  static constexpr int64_t kLeaf = 16;

  KdTree(const T* p, int64_t n, int64_t d) : pts(p), dim(d), perm(n) {
    std::iota(perm.begin(), perm.end(), int64_t(0));
    nodes.reserve(2 * (n / kLeaf + 1));
    if (n > 0) build(0, n);
  }

  int32_t build(int64_t lo, int64_t hi) {
    const int32_t id = (int32_t)nodes.size();
    nodes.push_back({lo, hi, -1, -1, 0, T(0)});
    if (hi - lo <= kLeaf) return id;
    // widest dimension of this cell
    int32_t best = 0;
    T spread = T(-1);
    for (int64_t d = 0; d < dim; ++d) {
      T mn = std::numeric_limits<T>::max(), mx = std::numeric_limits<T>::lowest();
      for (int64_t i = lo; i < hi; ++i) {
        const T v = pts[perm[i] * dim + d];
        mn = std::min(mn, v);
        mx = std::max(mx, v);
      }
      if (mx - mn > spread) {
        spread = mx - mn;
        best = (int32_t)d;
      }
    }
    if (spread <= T(0)) return id;  // all points coincide: keep as one (large) leaf
    const int64_t mid = (lo + hi) / 2;
    std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                     [&](int64_t a, int64_t b) { return pts[a * dim + best] < pts[b * dim + best]; });
    const T split = pts[perm[mid] * dim + best];
    const int32_t l = build(lo, mid);
    const int32_t r = build(mid, hi);
    nodes[id].left = l;
    nodes[id].right = r;
    nodes[id].dim = best;
    nodes[id].split = split;
    return id;
  }

  // heap: k (dist, index) pairs, max at front
  void search(int32_t id, const T* q, int64_t k, std::vector<std::pair<T, int64_t>>& heap) const {
    const Node& nd = nodes[id];
    if (nd.left < 0) {
      for (int64_t i = nd.lo; i < nd.hi; ++i) {
        const int64_t p = perm[i];
        T d2 = 0;
        for (int64_t d = 0; d < dim; ++d) {
          const T t = q[d] - pts[p * dim + d];
          d2 += t * t;
        }
        if ((int64_t)heap.size() < k) {
          heap.emplace_back(d2, p);
          std::push_heap(heap.begin(), heap.end());
        } else if (std::make_pair(d2, p) < heap.front()) {
          std::pop_heap(heap.begin(), heap.end());
          heap.back() = {d2, p};
          std::push_heap(heap.begin(), heap.end());
        }
      }
      return;
    }
    const T diff = q[nd.dim] - nd.split;
    const int32_t near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
    search(near, q, k, heap);
    if ((int64_t)heap.size() < k || diff * diff <= heap.front().first) search(far, q, k, heap);
  }
};

template <typename T>
void knn_segment(const T* data, int64_t n_data, const T* query, int64_t n_query, int64_t dim, int64_t k,
                 int64_t data_base, int64_t* out_idx, T* out_dist) {
  KdTree<T> tree(data, n_data, dim);
  const int64_t kk = std::min(k, n_data);
  parallel_for_dynamic(0, n_query, 256, [&](int64_t b, int64_t e) {
    std::vector<std::pair<T, int64_t>> heap;
    heap.reserve(kk + 1);
    for (int64_t qi = b; qi < e; ++qi) {
      heap.clear();
      if (kk > 0) tree.search(0, query + qi * dim, kk, heap);
      std::sort_heap(heap.begin(), heap.end());  // ascending distance, ties by index
      for (int64_t j = 0; j < k; ++j) {
        const bool ok = j < (int64_t)heap.size();
        out_idx[qi * k + j] = ok ? heap[j].second + data_base : -1;
        if (out_dist) out_dist[qi * k + j] = ok ? std::sqrt(heap[j].first) : std::numeric_limits<T>::infinity();
      }
    }
  });
}

template <typename T>
void knn(const T* data, const int64_t* data_off, const T* query, const int64_t* query_off, int64_t n_seg, int64_t dim,
         int64_t k, int64_t* out_idx, T* out_dist) {
  for (int64_t s = 0; s < n_seg; ++s) {
    const int64_t d0 = data_off[s], q0 = query_off[s];
    knn_segment<T>(data + d0 * dim, data_off[s + 1] - d0, query + q0 * dim, query_off[s + 1] - q0, dim, k, d0,
                   out_idx + q0 * k, out_dist ? out_dist + q0 * k : nullptr);
  }
}

template <typename T>
void fps(const T* pos, int64_t B, int64_t N, int64_t C, int64_t npoints, const int64_t* start, int64_t* out) {
  // clouds are independent: one thread per cloud when there are many, otherwise the distance
  // update of a single cloud is split across threads
  auto one_cloud = [&](int64_t b, bool inner_parallel) {
    const T* p = pos + b * N * C;
    std::vector<T> dist(N, std::numeric_limits<T>::max());
    int64_t cur = start[b];
    for (int64_t i = 0; i < npoints; ++i) {
      out[b * npoints + i] = cur;
      const T* c = p + cur * C;
      int64_t best = 0;
      T best_d = T(-1);
      if (inner_parallel && N >= 32768) {
#pragma omp parallel
        {
          int64_t lb = 0;
          T ld = T(-1);
#pragma omp for nowait
          for (int64_t j = 0; j < N; ++j) {
            T d2 = 0;
            for (int64_t d = 0; d < C; ++d) {
              const T t = p[j * C + d] - c[d];
              d2 += t * t;
            }
            if (d2 < dist[j]) dist[j] = d2;
            if (dist[j] > ld) {
              ld = dist[j];
              lb = j;
            }
          }
#pragma omp critical
          if (ld > best_d || (ld == best_d && lb < best)) {
            best_d = ld;
            best = lb;
          }
        }
      } else {
        for (int64_t j = 0; j < N; ++j) {
          T d2 = 0;
          for (int64_t d = 0; d < C; ++d) {
            const T t = p[j * C + d] - c[d];
            d2 += t * t;
          }
          if (d2 < dist[j]) dist[j] = d2;
          if (dist[j] > best_d) {
            best_d = dist[j];
            best = j;
          }
        }
      }
      cur = best;
    }
  };
  if (B >= omp_get_max_threads()) {
    parallel_for_dynamic(0, B, 1, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) one_cloud(i, false);
    });
  } else {
    for (int64_t b = 0; b < B; ++b) one_cloud(b, true);
  }
}

template <typename I>
void neighbor_matching(const I* indptr, const I* indices, const float* weight, int64_t n, const int64_t* order,
                       I* match) {
  // `match` arrives filled with -1.  Visit nodes in the given (random) order; an unmatched node takes its
  // heaviest unmatched neighbour (first unmatched neighbour when unweighted) or stays a singleton.
  for (int64_t t = 0; t < n; ++t) {
    const int64_t u = order[t];
    if (match[u] >= 0) continue;
    int64_t best = -1;
    float bw = -std::numeric_limits<float>::infinity();
    for (I p = indptr[u]; p < indptr[u + 1]; ++p) {
      const I v = indices[p];
      if (v == u || match[v] >= 0) continue;
      if (!weight) {
        best = v;
        break;
      }
      if (weight[p] > bw) {
        bw = weight[p];
        best = v;
      }
    }
    if (best >= 0) {
      match[u] = match[best] = (I)std::min<int64_t>(u, best);
    } else {
      match[u] = (I)u;
    }
  }
}

}  // namespace
}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_cpu_knn_kdtree(int dtype, const void* data, const int64_t* data_off, const void* query,
                                    const int64_t* query_off, int64_t n_seg, int64_t dim, int64_t k, int64_t* out_idx,
                                    void* out_dist) {
  DGLB_TRY
  if (dtype == kF32) {
    knn<float>((const float*)data, data_off, (const float*)query, query_off, n_seg, dim, k, out_idx, (float*)out_dist);
  } else if (dtype == kF64) {
    knn<double>((const double*)data, data_off, (const double*)query, query_off, n_seg, dim, k, out_idx,
                (double*)out_dist);
  } else {
    throw std::runtime_error("kd-tree kNN supports float32/float64 points");
  }
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_fps(int dtype, const void* pos, int64_t B, int64_t N, int64_t C, int64_t npoints,
                             const int64_t* start, int64_t* out) {
  DGLB_TRY
  DGLB_CHECK(npoints <= N, "cannot sample more points than the cloud holds");
  if (dtype == kF32) {
    fps<float>((const float*)pos, B, N, C, npoints, start, out);
  } else if (dtype == kF64) {
    fps<double>((const double*)pos, B, N, C, npoints, start, out);
  } else {
    throw std::runtime_error("farthest point sampling supports float32/float64 points");
  }
  DGLB_CATCH
}

DGLB_API int dglb200_cpu_neighbor_matching(int idtype, const void* indptr, const void* indices, const float* weight,
                                           int64_t n, const int64_t* order, void* match) {
  DGLB_TRY
  DGLB_ID_SWITCH(idtype, I, neighbor_matching<I>((const I*)indptr, (const I*)indices, weight, n, order, (I*)match));
  DGLB_CATCH
}
