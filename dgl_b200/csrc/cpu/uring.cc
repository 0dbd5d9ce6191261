// Row reads from a file through io_uring, without liburing: the rings are set up with the raw syscalls
// (io_uring_setup / io_uring_enter) and mapped by hand from <linux/io_uring.h>.
//
// One ring per worker thread, QD requests in flight per ring: a request is one IORING_OP_READ of a feature row
// (or the unread tail of a row after a short read) straight into the caller's (pinned) destination buffer, so the
// kernel overlaps thousands of NVMe / page-cache reads per call while the CPU only fills SQEs and reaps CQEs.
// Behavioural counterpart of the reference's OnDiskNpyArray::IndexSelectIOUring
// (/root/reference/graphbolt/src/cnumpy.cc:146-330), which batches preads through liburing with a semaphore-bounded queue.
// Containers often forbid io_uring (seccomp / sysctl kernel.io_uring_disabled): dglb200_uring_available() probes it
// once and the Python side falls back to the OpenMP pread path in diskio.cc.
#include <linux/io_uring.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstring>
#include <vector>

#include "common.h"

using namespace dglb;

namespace {

int sys_setup(unsigned entries, io_uring_params* p) { return (int)syscall(__NR_io_uring_setup, entries, p); }
int sys_enter(int fd, unsigned to_submit, unsigned min_complete, unsigned flags) {
  return (int)syscall(__NR_io_uring_enter, fd, to_submit, min_complete, flags, nullptr, 0);
}

template <typename T>
T load_acquire(const T* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
template <typename T>
void store_release(T* p, T v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

class Ring {
 public:
  explicit Ring(unsigned depth) {
    io_uring_params p;
    std::memset(&p, 0, sizeof(p));
    fd_ = sys_setup(depth, &p);
    if (fd_ < 0) { err_ = errno; return; }
    sq_bytes_ = p.sq_off.array + p.sq_entries * sizeof(unsigned);
    cq_bytes_ = p.cq_off.cqes + p.cq_entries * sizeof(io_uring_cqe);
    const bool single = p.features & IORING_FEAT_SINGLE_MMAP;
    if (single) sq_bytes_ = cq_bytes_ = std::max(sq_bytes_, cq_bytes_);
    sq_ = map(sq_bytes_, IORING_OFF_SQ_RING);
    cq_ = single ? sq_ : map(cq_bytes_, IORING_OFF_CQ_RING);
    sqe_bytes_ = p.sq_entries * sizeof(io_uring_sqe);
    sqes_ = (io_uring_sqe*)map(sqe_bytes_, IORING_OFF_SQES);
    if (!sq_ || !cq_ || !sqes_) { err_ = errno ? errno : ENOMEM; return; }
    char* s = (char*)sq_;
    sq_tail_ = (unsigned*)(s + p.sq_off.tail);
    sq_mask_ = *(unsigned*)(s + p.sq_off.ring_mask);
    sq_array_ = (unsigned*)(s + p.sq_off.array);
    char* c = (char*)cq_;
    cq_head_ = (unsigned*)(c + p.cq_off.head);
    cq_tail_ = (unsigned*)(c + p.cq_off.tail);
    cq_mask_ = *(unsigned*)(c + p.cq_off.ring_mask);
    cqes_ = (io_uring_cqe*)(c + p.cq_off.cqes);
    depth_ = p.sq_entries;
  }
  ~Ring() {
    if (sqes_) munmap(sqes_, sqe_bytes_);
    if (cq_ && cq_ != sq_) munmap(cq_, cq_bytes_);
    if (sq_) munmap(sq_, sq_bytes_);
    if (fd_ >= 0) close(fd_);
  }
  int error() const { return err_; }
  unsigned depth() const { return depth_; }

  // queue one read; the caller guarantees fewer than depth() requests are outstanding
  void push_read(int file, void* dst, unsigned len, uint64_t off, uint64_t tag) {
    const unsigned tail = *sq_tail_, slot = tail & sq_mask_;
    io_uring_sqe* e = &sqes_[slot];
    std::memset(e, 0, sizeof(*e));
    e->opcode = IORING_OP_READ;
    e->fd = file;
    e->addr = (uint64_t)(uintptr_t)dst;
    e->len = len;
    e->off = off;
    e->user_data = tag;
    sq_array_[slot] = slot;
    store_release(sq_tail_, tail + 1);
    ++pending_submit_;
  }
  // hand queued SQEs to the kernel and (optionally) wait for at least `wait` completions
  int flush(unsigned wait) {
    while (pending_submit_ > 0 || wait > 0) {
      int r = sys_enter(fd_, pending_submit_, wait, wait ? IORING_ENTER_GETEVENTS : 0);
      if (r < 0) {
        if (errno == EINTR || errno == EAGAIN || errno == EBUSY) continue;
        return errno;
      }
      pending_submit_ -= (unsigned)r;
      if (pending_submit_ == 0) break;
    }
    return 0;
  }
  // pop one completion if there is one
  bool pop(uint64_t* tag, int* res) {
    const unsigned head = *cq_head_;
    if (head == load_acquire(cq_tail_)) return false;
    const io_uring_cqe& c = cqes_[head & cq_mask_];
    *tag = c.user_data;
    *res = c.res;
    store_release(cq_head_, head + 1);
    return true;
  }

 private:
  void* map(size_t bytes, off_t off) {
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd_, off);
    return p == MAP_FAILED ? nullptr : p;
  }
  int fd_ = -1, err_ = 0;
  unsigned depth_ = 0, pending_submit_ = 0, sq_mask_ = 0, cq_mask_ = 0;
  size_t sq_bytes_ = 0, cq_bytes_ = 0, sqe_bytes_ = 0;
  void *sq_ = nullptr, *cq_ = nullptr;
  io_uring_sqe* sqes_ = nullptr;
  io_uring_cqe* cqes_ = nullptr;
  unsigned *sq_tail_ = nullptr, *sq_array_ = nullptr, *cq_head_ = nullptr, *cq_tail_ = nullptr;
};

// rows [lo, hi) of the request through one ring; returns 0 or an errno
int read_span(int file, int64_t data_offset, int64_t row_bytes, const int64_t* ix, int64_t lo, int64_t hi, char* dst,
              unsigned depth) {
  Ring ring(depth);
  if (ring.error()) return ring.error();
  std::vector<int64_t> done((size_t)(hi - lo), 0);   // bytes of each row already landed (short reads are resumed)
  int64_t next = lo, finished = lo;
  unsigned inflight = 0;
  auto submit_row = [&](int64_t i) {
    const int64_t got = done[(size_t)(i - lo)];
    const int64_t want = std::min<int64_t>(row_bytes - got, 1 << 30);
    ring.push_read(file, dst + i * row_bytes + got, (unsigned)want, (uint64_t)(data_offset + ix[i] * row_bytes + got), (uint64_t)i);
    ++inflight;
  };
  while (finished < hi) {
    while (next < hi && inflight < ring.depth()) submit_row(next++);
    if (int e = ring.flush(1)) return e;
    uint64_t tag;
    int res;
    while (ring.pop(&tag, &res)) {
      --inflight;
      const int64_t i = (int64_t)tag;
      if (res < 0) {
        if (res == -EAGAIN || res == -EINTR) { submit_row(i); continue; }
        return -res;
      }
      if (res == 0) return EIO;   // EOF inside a row: the index is out of range for this file
      done[(size_t)(i - lo)] += res;
      if (done[(size_t)(i - lo)] < row_bytes) submit_row(i);
      else ++finished;
    }
  }
  return 0;
}

}  // namespace

// *ok = 1 if the kernel lets this process create a ring, else 0 with the errno of io_uring_setup in *reason
DGLB_API int dglb200_uring_available(int* ok, int* reason) {
  static int cached = -1, why = 0;
  if (cached < 0) {
    Ring probe(4);
    why = probe.error();
    cached = why == 0 ? 1 : 0;
  }
  *ok = cached;
  if (reason) *reason = why;
  return 0;
}

// dst[i] = file[data_offset + idx[i] * row_bytes : + row_bytes], io_uring with `queue_depth` requests per ring and
// `num_threads` rings (each on its own OpenMP thread, contiguous spans of the request).
DGLB_API int dglb200_disk_read_rows_uring(int64_t fd, int64_t data_offset, int64_t row_bytes, const void* idx, int64_t n,
                                          void* dst, int num_threads, int queue_depth) {
  DGLB_TRY
  if (n <= 0 || row_bytes <= 0) return 0;
  if (queue_depth <= 0) queue_depth = 128;
  if (num_threads <= 0) num_threads = std::min<int>(8, std::max(1, omp_get_max_threads()));
  num_threads = (int)std::min<int64_t>(num_threads, std::max<int64_t>(1, n / 64));
  std::atomic<int> err{0};
  const int64_t span = (n + num_threads - 1) / num_threads;
#pragma omp parallel for schedule(static, 1) num_threads(num_threads)
  for (int t = 0; t < num_threads; ++t) {
    const int64_t lo = t * span, hi = std::min<int64_t>(n, lo + span);
    if (lo >= hi) continue;
    int e = read_span((int)fd, data_offset, row_bytes, (const int64_t*)idx, lo, hi, (char*)dst, (unsigned)queue_depth);
    if (e) err.store(e);
  }
  if (err.load()) throw std::runtime_error(std::string("io_uring read failed: ") + strerror(err.load()));
  DGLB_CATCH
}
