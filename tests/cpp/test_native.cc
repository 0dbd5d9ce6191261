// Native unit tests of libdglb200_cpu.so through its C ABI (the counterpart of the reference's gtest suite
// /root/reference/tests/cpp/: test_spmm.cc, test_partition.cc, message_queue_test.cc, socket_communicator_test.cc).
// No gtest in this image: a 30-line harness, built and run by tests/test_cpp_native.py.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <thread>
#include <vector>

extern "C" {
const char* dglb200_last_error();
int dglb200_cpu_gspmm(int op, int red, int dtype, int idtype, const void* indptr, const void* indices, const void* eids,
                      int64_t num_rows, const void* ufeat, const void* efeat, void* out, void* arg_u, void* arg_e,
                      int64_t F, int64_t ldiv, int64_t lmod, int64_t rdiv, int64_t rmod);
int dglb200_cache_create(int policy, int64_t capacity, int num_parts, int64_t* handle);
int dglb200_cache_destroy(int64_t handle);
int dglb200_cache_query(int64_t handle, const void* keys, int64_t n, void* positions, void* num_hit);
int dglb200_cache_replace(int64_t handle, const void* keys, int64_t n, void* positions);
int dglb200_cache_stats(int64_t handle, int64_t* hits, int64_t* queries);
int dglb200_cpu_ldg_partition(int idtype, const void* indptr, const void* indices, int64_t n, int k, const void* vwgt,
                              int refine_iters, double imbalance, int64_t seed, void* part);
int dglb200_cpu_libra_partition(int idtype, const void* src, const void* dst, int64_t E, int64_t n, int k, void* edge_part);
int dglb200_cpu_csr_mm_count(int idtype, int64_t n_rows, const void* ap, const void* ai, const void* bp, const void* bi,
                             int64_t n_cols, int64_t* row_nnz);
int dglb200_cpu_csr_mm_fill(int idtype, int dtype, int64_t n_rows, const void* ap, const void* ai, const void* av,
                            const void* bp, const void* bi, const void* bv, int64_t n_cols, const int64_t* cp, void* ci,
                            void* cv);
int dglb200_cpu_csr_sum_count(int idtype, int64_t n_rows, const void* ap, const void* ai, const void* bp, const void* bi,
                              int64_t* row_nnz);
int dglb200_cpu_csr_sum_fill(int idtype, int dtype, int64_t n_rows, const void* ap, const void* ai, const void* av,
                             double alpha, const void* bp, const void* bi, const void* bv, double beta, const int64_t* cp,
                             void* ci, void* cv, void* src_a, void* src_b);
int dglb200_cpu_csr_rowdot(int idtype, int dtype, int64_t nnz, const void* rows, const void* cols, const void* xp,
                           const void* xi, const void* xv, const void* yp, const void* yi, const void* yv, void* out);
int dglb200_cpu_knn_kdtree(int dtype, const void* data, const int64_t* data_off, const void* query, const int64_t* query_off,
                           int64_t n_seg, int64_t dim, int64_t k, int64_t* out_idx, void* out_dist);
int dglb200_cpu_fps(int dtype, const void* pos, int64_t B, int64_t N, int64_t C, int64_t npoints, const int64_t* start,
                    int64_t* out);
int dglb200_cpu_choice(int64_t num, int64_t population, int dtype, const void* prob, int replace, int64_t seed, int64_t* out);
int dglb200_cpu_bfs(int idtype, const void* indptr, const void* indices, const void* eids, int64_t n, const void* src,
                    int64_t n_src, int want_edges, void* out_ids, int64_t* sections, int64_t* n_sections, int64_t* n_ids);
int dglb200_cpu_rcmk_coo(int idtype, const void* u, const void* v, int64_t E, int64_t n, int64_t* perm);
int dglb200_rpc_create(int rank, const char* ip, int port, int64_t* handle, int* bound_port);
int dglb200_rpc_connect(int64_t handle, int peer, const char* ip, int port, int max_try, int retry_ms);
int dglb200_rpc_send(int64_t handle, int peer, int tag, const void* data, int64_t n);
int dglb200_rpc_recv_peek(int64_t handle, int timeout_ms, int* src, int* tag, int64_t* size);
int dglb200_rpc_recv_take(void* dst);
int dglb200_rpc_destroy(int64_t handle);
}

static std::atomic<int> g_failed{0}, g_checks{0};  // CHECK is used from helper threads too
#define CHECK(cond)                                                                  \
  do {                                                                               \
    ++g_checks;                                                                      \
    if (!(cond)) {                                                                   \
      ++g_failed;                                                                    \
      std::printf("FAILED %s:%d  %s   (last error: %s)\n", __FILE__, __LINE__, #cond, dglb200_last_error()); \
    }                                                                                \
  } while (0)

static void test_spmm() {
  // 4 destination rows, CSC: row r sums its in-neighbours' features
  std::vector<int64_t> indptr = {0, 2, 2, 5, 6}, indices = {1, 3, 0, 1, 2, 3};
  const int F = 3;
  std::vector<float> x(4 * F), w(6), out(4 * F, -1.f);
  for (int i = 0; i < 4 * F; ++i) x[i] = 0.5f * i - 1.f;
  for (int i = 0; i < 6; ++i) w[i] = 1.f + i;
  // copy_u sum
  CHECK(dglb200_cpu_gspmm(4, 0, 0, 1, indptr.data(), indices.data(), nullptr, 4, x.data(), nullptr, out.data(), nullptr,
                          nullptr, F, 1, F, 1, 1) == 0);
  for (int r = 0; r < 4; ++r)
    for (int f = 0; f < F; ++f) {
      float ref = 0;
      for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) ref += x[indices[p] * F + f];
      CHECK(std::fabs(out[r * F + f] - ref) < 1e-6);
    }
  // u_mul_e sum with a scalar per edge
  CHECK(dglb200_cpu_gspmm(2, 0, 0, 1, indptr.data(), indices.data(), nullptr, 4, x.data(), w.data(), out.data(), nullptr,
                          nullptr, F, 1, F, 1, 1) == 0);
  for (int r = 0; r < 4; ++r)
    for (int f = 0; f < F; ++f) {
      float ref = 0;
      for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) ref += x[indices[p] * F + f] * w[p];
      CHECK(std::fabs(out[r * F + f] - ref) < 1e-5);
    }
  // copy_u max with arg
  std::vector<int64_t> arg(4 * F, -7);
  CHECK(dglb200_cpu_gspmm(4, 1, 0, 1, indptr.data(), indices.data(), nullptr, 4, x.data(), nullptr, out.data(), arg.data(),
                          nullptr, F, 1, F, 1, 1) == 0);
  CHECK(arg[0] == 3 && out[0] == x[3 * F]);   // row 0: neighbours {1,3} -> node 3 has the larger feature
  CHECK(out[1 * F] == 0.f);                   // empty row -> 0
}

static void test_cache(int policy) {
  int64_t h = 0;
  CHECK(dglb200_cache_create(policy, 64, 2, &h) == 0);
  std::vector<int64_t> keys(40), pos(40);
  for (int i = 0; i < 40; ++i) keys[i] = i * 3;
  int64_t nhit = -1;
  CHECK(dglb200_cache_query(h, keys.data(), 40, pos.data(), &nhit) == 0);
  CHECK(nhit == 0);
  CHECK(dglb200_cache_replace(h, keys.data(), 40, pos.data()) == 0);
  std::set<int64_t> slots(pos.begin(), pos.end());
  CHECK(slots.size() == 40);  // distinct slots
  for (auto s : slots) CHECK(s >= 0 && s < 64);
  std::vector<int64_t> pos2(40);
  CHECK(dglb200_cache_query(h, keys.data(), 40, pos2.data(), &nhit) == 0);
  CHECK(nhit == 40);
  // the query reorders hits first; every returned slot must be one that was assigned
  for (int i = 0; i < 40; ++i) CHECK(slots.count(pos2[i]) == 1);
  // overflow the capacity: old keys get evicted, the structure stays consistent
  std::vector<int64_t> more(200), mpos(200);
  for (int i = 0; i < 200; ++i) more[i] = 1000 + i;
  for (int b = 0; b < 200; b += 50) {
    CHECK(dglb200_cache_query(h, more.data() + b, 50, mpos.data(), &nhit) == 0);
    CHECK(dglb200_cache_replace(h, more.data() + b + 0, 50, mpos.data()) == 0);
  }
  int64_t hits = 0, queries = 0;
  CHECK(dglb200_cache_stats(h, &hits, &queries) == 0);
  CHECK(queries >= 280 && hits >= 40);
  CHECK(dglb200_cache_destroy(h) == 0);
  // fewer slots than requested shards: slot ids must still stay inside the cache tensor
  int64_t tiny = 0;
  CHECK(dglb200_cache_create(policy, 3, 8, &tiny) == 0);
  std::vector<int64_t> tk(100), tp(100);
  for (int i = 0; i < 100; ++i) tk[i] = 7 * i + 1;
  for (int b = 0; b < 100; b += 2) {
    CHECK(dglb200_cache_replace(tiny, tk.data() + b, 2, tp.data() + b) == 0);
    for (int j = b; j < b + 2; ++j) CHECK(tp[j] >= -1 && tp[j] < 3);
  }
  CHECK(dglb200_cache_destroy(tiny) == 0);
}

static void test_partitioners() {
  // two 50-node cliques joined by one edge: a 2-way partition must separate them
  const int n = 100;
  std::vector<int64_t> indptr(n + 1, 0), indices;
  for (int v = 0; v < n; ++v) {
    const int lo = v < 50 ? 0 : 50;
    for (int u = lo; u < lo + 50; ++u)
      if (u != v) indices.push_back(u);
    if (v == 49) indices.push_back(50);
    if (v == 50) indices.push_back(49);
    indptr[v + 1] = (int64_t)indices.size();
  }
  std::vector<int64_t> part(n, -1);
  CHECK(dglb200_cpu_ldg_partition(1, indptr.data(), indices.data(), n, 2, nullptr, 10, 1.05, 7, part.data()) == 0);
  int a = 0, cut = 0;
  for (int v = 0; v < n; ++v) a += part[v] == part[0];
  for (int v = 0; v < n; ++v)
    for (int64_t p = indptr[v]; p < indptr[v + 1]; ++p) cut += part[v] != part[indices[p]];
  CHECK(a >= 45 && a <= 55);
  CHECK(cut <= 2 * 60);  // far below the ~2500 of a random split
  // Libra vertex cut on a sparse pseudo-random graph: every edge gets a partition in range, loads stay balanced
  const int64_t n2 = 2000, E2 = 16000;
  std::vector<int64_t> src(E2), dst(E2);
  uint64_t st = 12345;
  for (int64_t e = 0; e < E2; ++e) {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    src[e] = (int64_t)((st >> 33) % n2);
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    dst[e] = (int64_t)((st >> 33) % n2);
  }
  std::vector<int64_t> ep(E2, -1), load(4, 0);
  CHECK(dglb200_cpu_libra_partition(1, src.data(), dst.data(), E2, n2, 4, ep.data()) == 0);
  for (auto p : ep) { CHECK(p >= 0 && p < 4); if (p >= 0 && p < 4) ++load[p]; }
  for (auto l : load) CHECK(l > E2 / 8 && l < E2 / 2);
}

static void test_rpc_loopback() {
  int64_t a = 0, b = 0;
  int pa = 0, pb = 0;
  CHECK(dglb200_rpc_create(0, "127.0.0.1", 0, &a, &pa) == 0);
  CHECK(dglb200_rpc_create(1, "127.0.0.1", 0, &b, &pb) == 0);
  CHECK(pa > 0 && pb > 0);
  CHECK(dglb200_rpc_connect(a, 1, "127.0.0.1", pb, 50, 20) == 0);
  std::vector<char> payload(1 << 20);
  for (size_t i = 0; i < payload.size(); ++i) payload[i] = (char)(i * 31);
  std::thread sender([&] {
    for (int k = 0; k < 4; ++k) CHECK(dglb200_rpc_send(a, 1, 100 + k, payload.data(), (int64_t)payload.size() - k) == 0);
    CHECK(dglb200_rpc_send(a, 1, 7, nullptr, 0) == 0);  // empty message
  });
  for (int k = 0; k < 5; ++k) {
    int src = -1, tag = -1;
    int64_t size = -2;
    CHECK(dglb200_rpc_recv_peek(b, 5000, &src, &tag, &size) == 0);
    CHECK(src == 0);
    std::vector<char> buf(size > 0 ? size : 1);
    CHECK(dglb200_rpc_recv_take(buf.data()) == 0);
    if (k < 4) {
      CHECK(tag == 100 + k && size == (int64_t)payload.size() - k);
      CHECK(std::memcmp(buf.data(), payload.data(), (size_t)size) == 0);
    } else {
      CHECK(tag == 7 && size == 0);
    }
  }
  sender.join();
  int src, tag;
  int64_t size = 0;
  CHECK(dglb200_rpc_recv_peek(b, 50, &src, &tag, &size) == 0);
  CHECK(size == -1);  // timeout, nothing pending
  CHECK(dglb200_rpc_destroy(a) == 0);
  CHECK(dglb200_rpc_destroy(b) == 0);
}

// csr_mm / csr_sum / masked row-dot against a dense evaluation (reference tests/cpp/test_csrmm.cc)
static void test_csr_arith() {
  const int M = 5, K = 4, N = 6;
  // A (5x4) and B (4x6) with sorted rows; row 3 of A is empty
  std::vector<int64_t> ap = {0, 2, 3, 5, 5, 7}, ai = {0, 2, 1, 0, 3, 1, 2};
  std::vector<float> av = {1, 2, 3, 4, 5, 6, 7};
  std::vector<int64_t> bp = {0, 2, 4, 5, 7}, bi = {1, 4, 0, 4, 5, 2, 3};
  std::vector<float> bv = {1, -1, 2, 3, 4, 5, 6};
  auto dense = [](int rows, int cols, const std::vector<int64_t>& p, const std::vector<int64_t>& i,
                  const std::vector<float>& v) {
    std::vector<float> d(rows * cols, 0.f);
    for (int r = 0; r < rows; ++r)
      for (int64_t e = p[r]; e < p[r + 1]; ++e) d[r * cols + i[e]] += v[e];
    return d;
  };
  std::vector<float> da = dense(M, K, ap, ai, av), db = dense(K, N, bp, bi, bv), dc(M * N, 0.f);
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < K; ++k)
      for (int c = 0; c < N; ++c) dc[r * N + c] += da[r * K + k] * db[k * N + c];
  std::vector<int64_t> nnz(M), cp(M + 1, 0);
  CHECK(dglb200_cpu_csr_mm_count(1, M, ap.data(), ai.data(), bp.data(), bi.data(), N, nnz.data()) == 0);
  for (int r = 0; r < M; ++r) cp[r + 1] = cp[r] + nnz[r];
  CHECK(nnz[3] == 0);
  std::vector<int64_t> ci(cp[M]);
  std::vector<float> cv(cp[M]);
  CHECK(dglb200_cpu_csr_mm_fill(1, 0, M, ap.data(), ai.data(), av.data(), bp.data(), bi.data(), bv.data(), N, cp.data(),
                                ci.data(), cv.data()) == 0);
  std::vector<float> got = dense(M, N, cp, ci, cv);
  for (int i = 0; i < M * N; ++i) CHECK(std::fabs(got[i] - dc[i]) < 1e-5);
  for (int r = 0; r < M; ++r)
    for (int64_t e = cp[r] + 1; e < cp[r + 1]; ++e) CHECK(ci[e - 1] < ci[e]);  // sorted, unique
  // structural product (null values = all ones) counts paths
  CHECK(dglb200_cpu_csr_mm_fill(1, 0, M, ap.data(), ai.data(), nullptr, bp.data(), bi.data(), nullptr, N, cp.data(),
                                ci.data(), cv.data()) == 0);
  for (int64_t e = 0; e < cp[M]; ++e) CHECK(cv[e] >= 1.f && cv[e] == std::floor(cv[e]));

  // C2 = 2*A - A' where A' shares the shape of A
  std::vector<int64_t> a2p = {0, 1, 3, 3, 4, 6}, a2i = {2, 0, 1, 3, 0, 2};
  std::vector<float> a2v = {10, 20, 30, 40, 50, 60};
  std::vector<int64_t> sn(M), sp(M + 1, 0);
  CHECK(dglb200_cpu_csr_sum_count(1, M, ap.data(), ai.data(), a2p.data(), a2i.data(), sn.data()) == 0);
  for (int r = 0; r < M; ++r) sp[r + 1] = sp[r] + sn[r];
  std::vector<int64_t> si(sp[M]), sa(sp[M]), sb(sp[M]);
  std::vector<float> sv(sp[M]);
  CHECK(dglb200_cpu_csr_sum_fill(1, 0, M, ap.data(), ai.data(), av.data(), 2.0, a2p.data(), a2i.data(), a2v.data(), -1.0,
                                 sp.data(), si.data(), sv.data(), sa.data(), sb.data()) == 0);
  std::vector<float> ds = dense(M, K, sp, si, sv), d2 = dense(M, K, a2p, a2i, a2v);
  for (int i = 0; i < M * K; ++i) CHECK(std::fabs(ds[i] - (2 * da[i] - d2[i])) < 1e-5);
  for (int64_t e = 0; e < sp[M]; ++e) {
    CHECK(sa[e] >= 0 || sb[e] >= 0);
    if (sa[e] >= 0) CHECK(ai[sa[e]] == si[e]);
    if (sb[e] >= 0) CHECK(a2i[sb[e]] == si[e]);
  }

  // row-dot: <A[r,:], A'[c,:]> for a few (r, c) pairs
  std::vector<int64_t> rows = {0, 1, 4, 3, 2}, cols = {0, 1, 4, 2, 4};
  std::vector<float> out(rows.size());
  CHECK(dglb200_cpu_csr_rowdot(1, 0, (int64_t)rows.size(), rows.data(), cols.data(), ap.data(), ai.data(), av.data(),
                               a2p.data(), a2i.data(), a2v.data(), out.data()) == 0);
  for (size_t t = 0; t < rows.size(); ++t) {
    float ref = 0;
    for (int k = 0; k < K; ++k) ref += da[rows[t] * K + k] * d2[cols[t] * K + k];
    CHECK(std::fabs(out[t] - ref) < 1e-4);
  }
}

// Larger inputs so the OpenMP paths (dynamic chunks, per-thread scratch) really run multi-threaded; this is what the
// ThreadSanitizer build of script/run_sanitizers.sh --tsan looks at.
static void test_parallel_paths() {
  const int64_t n = 3000;
  // random sparse matrix with ~8 nnz per row, sorted rows
  std::vector<int64_t> ap(n + 1, 0), ai;
  std::vector<float> av;
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (int64_t r = 0; r < n; ++r) {
    std::set<int64_t> cols;
    while (cols.size() < 8) cols.insert((int64_t)(rnd() % n));
    for (auto c : cols) { ai.push_back(c); av.push_back(1.f + (float)(rnd() % 7)); }
    ap[r + 1] = (int64_t)ai.size();
  }
  std::vector<int64_t> nnz(n), cp(n + 1, 0);
  CHECK(dglb200_cpu_csr_mm_count(1, n, ap.data(), ai.data(), ap.data(), ai.data(), n, nnz.data()) == 0);
  for (int64_t r = 0; r < n; ++r) cp[r + 1] = cp[r] + nnz[r];
  std::vector<int64_t> ci(cp[n]);
  std::vector<float> cv(cp[n]);
  CHECK(dglb200_cpu_csr_mm_fill(1, 0, n, ap.data(), ai.data(), av.data(), ap.data(), ai.data(), av.data(), n, cp.data(),
                                ci.data(), cv.data()) == 0);
  // spot-check a few rows against a direct evaluation
  for (int64_t r = 0; r < n; r += 997) {
    std::vector<double> acc(n, 0.0);
    for (int64_t x = ap[r]; x < ap[r + 1]; ++x)
      for (int64_t y = ap[ai[x]]; y < ap[ai[x] + 1]; ++y) acc[ai[y]] += (double)av[x] * av[y];
    for (int64_t e = cp[r]; e < cp[r + 1]; ++e) CHECK(std::fabs(cv[e] - acc[ci[e]]) < 1e-3);
  }
  // kNN: every point's nearest neighbour is itself
  std::vector<float> pts(3 * 5000);
  for (auto& p : pts) p = (float)(rnd() % 100000) / 1000.f;
  std::vector<int64_t> off = {0, 5000}, idx(5000 * 4);
  CHECK(dglb200_cpu_knn_kdtree(0, pts.data(), off.data(), pts.data(), off.data(), 1, 3, 4, idx.data(), nullptr) == 0);
  for (int64_t i = 0; i < 5000; i += 13) CHECK(idx[i * 4] == i || (pts[3 * i] == pts[3 * idx[i * 4]]));
  // FPS over many clouds (thread per cloud) and one big cloud (threads inside the cloud)
  std::vector<int64_t> start(64, 0), out(64 * 8);
  std::vector<float> clouds(64 * 200 * 3);
  for (auto& p : clouds) p = (float)(rnd() % 1000) / 10.f;
  CHECK(dglb200_cpu_fps(0, clouds.data(), 64, 200, 3, 8, start.data(), out.data()) == 0);
  for (int b = 0; b < 64; ++b) { std::set<int64_t> u(out.begin() + b * 8, out.begin() + b * 8 + 8); CHECK(u.size() == 8); }
  std::vector<float> big(40000 * 3);
  for (auto& p : big) p = (float)(rnd() % 100000) / 100.f;
  std::vector<int64_t> bout(16), bstart(1, 0);
  CHECK(dglb200_cpu_fps(0, big.data(), 1, 40000, 3, 16, bstart.data(), bout.data()) == 0);
  { std::set<int64_t> u(bout.begin(), bout.end()); CHECK(u.size() == 16); }
  // parallel uniform draws
  std::vector<int64_t> draws(200000);
  CHECK(dglb200_cpu_choice(200000, 1000, 0, nullptr, 1, 7, draws.data()) == 0);
  for (int64_t i = 0; i < 200000; i += 101) CHECK(draws[i] >= 0 && draws[i] < 1000);
  // BFS + RCM on the matrix above viewed as a graph
  std::vector<int64_t> ids(n), sections(n + 1), counts(2, 0), src = {0};
  CHECK(dglb200_cpu_bfs(1, ap.data(), ai.data(), nullptr, n, src.data(), 1, 0, ids.data(), sections.data(), &counts[0],
                        &counts[1]) == 0);
  CHECK(counts[1] > n / 2);
  std::vector<int64_t> eu, ev, perm(n);
  for (int64_t r = 0; r < n; ++r)
    for (int64_t e = ap[r]; e < ap[r + 1]; ++e) { eu.push_back(r); ev.push_back(ai[e]); }
  CHECK(dglb200_cpu_rcmk_coo(1, eu.data(), ev.data(), (int64_t)eu.size(), n, perm.data()) == 0);
  { std::set<int64_t> u(perm.begin(), perm.end()); CHECK((int64_t)u.size() == n); }
}

int main() {
  test_spmm();
  test_parallel_paths();
  test_csr_arith();
  for (int policy = 0; policy < 4; ++policy) test_cache(policy);  // S3-FIFO, SIEVE, LRU, CLOCK
  test_partitioners();
  test_rpc_loopback();
  std::printf("%d checks, %d failed\n", g_checks.load(), g_failed.load());
  return g_failed ? 1 : 0;
}
