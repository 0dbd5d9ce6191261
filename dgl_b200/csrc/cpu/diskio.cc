// Disk-backed feature reads (OnDiskNpyArray equivalent) and POSIX shared-memory helpers.
// Reference behaviour: /root/reference/graphbolt/src/cnumpy.cc (io_uring reads of selected rows of a
// .npy file), src/runtime/shared_mem.cc (named POSIX shm for graphs/tensors shared between the
// server/sampler/trainer processes of one host). liburing is not available in this image, so row
// reads are issued as parallel pread(2) calls from an OpenMP team (each thread keeps many 4 KiB-
// aligned requests in flight through the page cache / NVMe queue).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>

#include "common.h"

using namespace dglb;

DGLB_API int dglb200_disk_open(const char* path, int64_t* fd_out) {
  DGLB_TRY
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) throw std::runtime_error(std::string("open failed: ") + path + ": " + strerror(errno));
  *fd_out = fd;
  DGLB_CATCH
}

DGLB_API int dglb200_disk_close(int64_t fd) {
  DGLB_TRY
  ::close((int)fd);
  DGLB_CATCH
}

// dst[i] = file[data_offset + idx[i]*row_bytes : +row_bytes]
DGLB_API int dglb200_disk_read_rows(int64_t fd, int64_t data_offset, int64_t row_bytes, const void* idx, int64_t n,
                                    void* dst, int num_threads) {
  DGLB_TRY
  const int64_t* ix = (const int64_t*)idx;
  int err = 0;
  if (num_threads <= 0) num_threads = std::min(64, omp_get_max_threads() * 4);
#pragma omp parallel for schedule(dynamic, 16) num_threads(num_threads)
  for (int64_t i = 0; i < n; ++i) {
    char* out = (char*)dst + i * row_bytes;
    int64_t off = data_offset + ix[i] * row_bytes, left = row_bytes;
    while (left > 0) {
      ssize_t r = ::pread((int)fd, out, (size_t)left, (off_t)off);
      if (r <= 0) {
#pragma omp atomic write
        err = errno ? errno : EIO;
        break;
      }
      out += r; off += r; left -= r;
    }
  }
  if (err) throw std::runtime_error(std::string("pread failed: ") + strerror(err));
  DGLB_CATCH
}

// ---- POSIX shared memory --------------------------------------------------------------------
DGLB_API int dglb200_shm_create(const char* name, int64_t bytes, void** ptr_out) {
  DGLB_TRY
  int fd = ::shm_open(name, O_CREAT | O_RDWR, 0600);
  if (fd < 0) throw std::runtime_error(std::string("shm_open(create) failed: ") + name + ": " + strerror(errno));
  if (::ftruncate(fd, (off_t)bytes) != 0) {
    ::close(fd);
    throw std::runtime_error(std::string("ftruncate failed: ") + strerror(errno));
  }
  void* p = ::mmap(nullptr, (size_t)std::max<int64_t>(bytes, 1), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  ::close(fd);
  if (p == MAP_FAILED) throw std::runtime_error(std::string("mmap failed: ") + strerror(errno));
  *ptr_out = p;
  DGLB_CATCH
}

DGLB_API int dglb200_shm_open(const char* name, int64_t bytes, void** ptr_out) {
  DGLB_TRY
  int fd = ::shm_open(name, O_RDWR, 0600);
  if (fd < 0) throw std::runtime_error(std::string("shm_open failed: ") + name + ": " + strerror(errno));
  void* p = ::mmap(nullptr, (size_t)std::max<int64_t>(bytes, 1), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  ::close(fd);
  if (p == MAP_FAILED) throw std::runtime_error(std::string("mmap failed: ") + strerror(errno));
  *ptr_out = p;
  DGLB_CATCH
}

DGLB_API int dglb200_shm_exists(const char* name) {
  int fd = ::shm_open(name, O_RDONLY, 0600);
  if (fd < 0) return 0;
  ::close(fd);
  return 1;
}

DGLB_API int dglb200_shm_unmap(void* p, int64_t bytes) {
  DGLB_TRY
  ::munmap(p, (size_t)std::max<int64_t>(bytes, 1));
  DGLB_CATCH
}

DGLB_API int dglb200_shm_unlink(const char* name) {
  DGLB_TRY
  ::shm_unlink(name);
  DGLB_CATCH
}
