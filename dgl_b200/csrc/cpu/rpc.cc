// Message-oriented TCP transport for the multi-node RPC layer (DistGraph sampling / KVStore).
// Reference behaviour: /root/reference/src/rpc/network/socket_communicator.{h,cc} (per-connection
// sender threads, receiver threads, blocking message queues), tcp_socket.cc, msg_queue.cc.
// Fresh design: one Endpoint per process owns a listening socket, an epoll-driven receiver thread
// that reassembles length-prefixed frames into a single MPSC inbox, and per-peer send mutexes
// (sends are synchronous writev calls; large tensors go out zero-copy as extra iovecs).
// Frame: [u64 total_len][i32 src_rank][i32 tag][payload bytes...]
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/epoll.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "common.h"

namespace dglb {

// upper bound of one frame (payload + 8-byte header); DGLB200_RPC_MAX_FRAME_MB overrides the 1 GiB default
static const uint64_t kMaxFrameBytes = [] {
  const char* e = std::getenv("DGLB200_RPC_MAX_FRAME_MB");
  const uint64_t mb = e ? std::strtoull(e, nullptr, 10) : 1024;
  return (mb ? mb : 1024) << 20;
}();

struct Message {
  int32_t src, tag;
  std::vector<char> data;
};

struct Conn {
  int fd = -1;
  std::vector<char> buf;  // partial frame
  std::mutex send_mu;
};

class Endpoint {
 public:
  Endpoint(int rank, const std::string& ip, int port) : rank_(rank) {
    listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (listen_fd_ < 0) throw std::runtime_error("socket() failed");
    int one = 1;
    setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    if (inet_pton(AF_INET, ip.c_str(), &addr.sin_addr) != 1) throw std::runtime_error("bad ip " + ip);
    if (::bind(listen_fd_, (sockaddr*)&addr, sizeof(addr)) != 0)
      throw std::runtime_error(std::string("bind failed: ") + strerror(errno));
    if (::listen(listen_fd_, 1024) != 0) throw std::runtime_error("listen failed");
    socklen_t len = sizeof(addr);
    getsockname(listen_fd_, (sockaddr*)&addr, &len);
    port_ = ntohs(addr.sin_port);
    ep_ = epoll_create1(0);
    epoll_event ev{};
    ev.events = EPOLLIN;
    ev.data.fd = listen_fd_;
    epoll_ctl(ep_, EPOLL_CTL_ADD, listen_fd_, &ev);
    running_ = true;
    recv_thread_ = std::thread([this] { RecvLoop(); });
  }
  ~Endpoint() { Shutdown(); }

  int port() const { return port_; }

  void Connect(int peer, const std::string& ip, int port, int max_try, int retry_ms) {
    for (int t = 0; t < max_try; ++t) {
      int fd = ::socket(AF_INET, SOCK_STREAM, 0);
      sockaddr_in addr{};
      addr.sin_family = AF_INET;
      addr.sin_port = htons((uint16_t)port);
      inet_pton(AF_INET, ip.c_str(), &addr.sin_addr);
      if (::connect(fd, (sockaddr*)&addr, sizeof(addr)) == 0) {
        int one = 1;
        setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        auto c = std::make_shared<Conn>();
        c->fd = fd;
        std::lock_guard<std::mutex> lk(mu_);
        out_[peer] = c;
        return;
      }
      ::close(fd);
      std::this_thread::sleep_for(std::chrono::milliseconds(retry_ms));
    }
    throw std::runtime_error("connect to peer " + std::to_string(peer) + " (" + ip + ":" + std::to_string(port) +
                             ") failed after " + std::to_string(max_try) + " tries");
  }

  void Send(int peer, int tag, const void* data, int64_t n) {
    std::shared_ptr<Conn> c;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = out_.find(peer);
      if (it == out_.end()) throw std::runtime_error("no connection to peer " + std::to_string(peer));
      c = it->second;
    }
    uint64_t total = 8 + (uint64_t)n;
    int32_t hdr[2] = {rank_, tag};
    iovec iov[3] = {{&total, 8}, {hdr, 8}, {const_cast<void*>(data), (size_t)n}};
    std::lock_guard<std::mutex> lk(c->send_mu);
    int idx = 0;
    size_t left = 16 + (size_t)n;
    while (left > 0) {
      ssize_t w = ::writev(c->fd, iov + idx, 3 - idx);
      if (w < 0) {
        if (errno == EINTR) continue;
        throw std::runtime_error(std::string("send failed: ") + strerror(errno));
      }
      left -= (size_t)w;
      while (idx < 3 && (size_t)w >= iov[idx].iov_len) { w -= iov[idx].iov_len; ++idx; }
      if (idx < 3 && w > 0) { iov[idx].iov_base = (char*)iov[idx].iov_base + w; iov[idx].iov_len -= w; }
    }
  }

  // blocks up to timeout_ms; returns false on timeout
  bool Recv(Message* out, int timeout_ms) {
    std::unique_lock<std::mutex> lk(q_mu_);
    if (!q_cv_.wait_for(lk, std::chrono::milliseconds(timeout_ms), [this] { return !inbox_.empty() || !running_; }))
      return false;
    if (inbox_.empty()) return false;
    *out = std::move(inbox_.front());
    inbox_.pop_front();
    return true;
  }

  void Shutdown() {
    bool was = running_.exchange(false);
    if (!was) return;
    q_cv_.notify_all();
    if (recv_thread_.joinable()) recv_thread_.join();
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : out_) ::close(kv.second->fd);
    for (auto& kv : in_) ::close(kv.first);
    out_.clear();
    in_.clear();
    ::close(listen_fd_);
    ::close(ep_);
  }

 private:
  void RecvLoop() {
    std::vector<epoll_event> evs(64);
    std::vector<char> tmp(1 << 20);
    while (running_) {
      int n = epoll_wait(ep_, evs.data(), (int)evs.size(), 100);
      for (int i = 0; i < n; ++i) {
        int fd = evs[i].data.fd;
        if (fd == listen_fd_) {
          int cfd = ::accept(listen_fd_, nullptr, nullptr);
          if (cfd < 0) continue;
          int one = 1;
          setsockopt(cfd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
          epoll_event ev{};
          ev.events = EPOLLIN;
          ev.data.fd = cfd;
          epoll_ctl(ep_, EPOLL_CTL_ADD, cfd, &ev);
          std::lock_guard<std::mutex> lk(mu_);
          in_[cfd] = std::make_shared<Conn>();
          in_[cfd]->fd = cfd;
          continue;
        }
        std::shared_ptr<Conn> c;
        {
          std::lock_guard<std::mutex> lk(mu_);
          auto it = in_.find(fd);
          if (it == in_.end()) continue;
          c = it->second;
        }
        ssize_t r = ::read(fd, tmp.data(), tmp.size());
        if (r <= 0) {
          epoll_ctl(ep_, EPOLL_CTL_DEL, fd, nullptr);
          ::close(fd);
          std::lock_guard<std::mutex> lk(mu_);
          in_.erase(fd);
          continue;
        }
        c->buf.insert(c->buf.end(), tmp.data(), tmp.data() + r);
        size_t pos = 0;
        bool bad_frame = false;
        while (c->buf.size() - pos >= 8) {
          uint64_t total;
          std::memcpy(&total, c->buf.data() + pos, 8);
          // the length prefix comes from the network: a frame carries at least (src, tag) = 8 bytes and at most
          // kMaxFrameBytes; anything else is a protocol violation (or an attack) and drops the connection instead of
          // running assign() with first > last or growing the buffer without bound
          if (total < 8 || total > kMaxFrameBytes) {
            bad_frame = true;
            break;
          }
          if (c->buf.size() - pos < 8 + total) break;
          Message m;
          std::memcpy(&m.src, c->buf.data() + pos + 8, 4);
          std::memcpy(&m.tag, c->buf.data() + pos + 12, 4);
          m.data.assign(c->buf.begin() + pos + 16, c->buf.begin() + pos + 8 + total);
          pos += 8 + total;
          {
            std::lock_guard<std::mutex> lk(q_mu_);
            inbox_.push_back(std::move(m));
          }
          q_cv_.notify_one();
        }
        if (bad_frame) {
          epoll_ctl(ep_, EPOLL_CTL_DEL, fd, nullptr);
          ::close(fd);
          std::lock_guard<std::mutex> lk(mu_);
          in_.erase(fd);
          continue;
        }
        if (pos > 0) c->buf.erase(c->buf.begin(), c->buf.begin() + pos);
      }
    }
  }

  int rank_, listen_fd_, ep_, port_;
  std::atomic<bool> running_{false};
  std::thread recv_thread_;
  std::mutex mu_, q_mu_;
  std::condition_variable q_cv_;
  std::map<int, std::shared_ptr<Conn>> out_;
  std::map<int, std::shared_ptr<Conn>> in_;
  std::deque<Message> inbox_;
};

}  // namespace dglb

using namespace dglb;

DGLB_API int dglb200_rpc_create(int rank, const char* ip, int port, int64_t* handle, int* bound_port) {
  DGLB_TRY
  auto* e = new Endpoint(rank, ip, port);
  *handle = (int64_t)(intptr_t)e;
  *bound_port = e->port();
  DGLB_CATCH
}

DGLB_API int dglb200_rpc_connect(int64_t handle, int peer, const char* ip, int port, int max_try, int retry_ms) {
  DGLB_TRY
  ((Endpoint*)(intptr_t)handle)->Connect(peer, ip, port, max_try, retry_ms);
  DGLB_CATCH
}

DGLB_API int dglb200_rpc_send(int64_t handle, int peer, int tag, const void* data, int64_t n) {
  DGLB_TRY
  ((Endpoint*)(intptr_t)handle)->Send(peer, tag, data, n);
  DGLB_CATCH
}

// two-phase receive: first call returns src/tag/size and parks the message; second call copies it out
static thread_local Message g_parked;
static thread_local bool g_has_parked = false;

DGLB_API int dglb200_rpc_recv_peek(int64_t handle, int timeout_ms, int* src, int* tag, int64_t* size) {
  DGLB_TRY
  if (!g_has_parked) {
    if (!((Endpoint*)(intptr_t)handle)->Recv(&g_parked, timeout_ms)) {
      *size = -1;
      return 0;
    }
    g_has_parked = true;
  }
  *src = g_parked.src;
  *tag = g_parked.tag;
  *size = (int64_t)g_parked.data.size();
  DGLB_CATCH
}

DGLB_API int dglb200_rpc_recv_take(void* dst) {
  DGLB_TRY
  DGLB_CHECK(g_has_parked, "no parked message");
  if (!g_parked.data.empty()) std::memcpy(dst, g_parked.data.data(), g_parked.data.size());
  g_parked = Message();
  g_has_parked = false;
  DGLB_CATCH
}

DGLB_API int dglb200_rpc_destroy(int64_t handle) {
  DGLB_TRY
  delete (Endpoint*)(intptr_t)handle;
  DGLB_CATCH
}
