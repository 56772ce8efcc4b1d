// Small pieces shared by the drop-in command lines: page-locked buffers and a blocking queue for the
// reader -> codec -> writer pipeline (the reference tools are single threaded, EncodeStream.cpp:452-770).
#ifndef VC2_HOST_PIPELINE_H
#define VC2_HOST_PIPELINE_H
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "vc2_cabi.h"

namespace vc2cli {

// host buffer for vc2_codec_encode_host / _decode_host: pinned when the driver grants it, ordinary memory otherwise
class HostBuf {
 public:
  HostBuf() : p_(nullptr), n_(0), pinned_(false) {}
  explicit HostBuf(size_t n) : p_(nullptr), n_(0), pinned_(false) { resize(n); }
  HostBuf(HostBuf&& o) : p_(o.p_), n_(o.n_), pinned_(o.pinned_) { o.p_ = nullptr; o.n_ = 0; }
  HostBuf& operator=(HostBuf&& o) {
    if (this != &o) { release(); p_ = o.p_; n_ = o.n_; pinned_ = o.pinned_; o.p_ = nullptr; o.n_ = 0; }
    return *this;
  }
  ~HostBuf() { release(); }
  void resize(size_t n) {
    release();
    p_ = static_cast<uint8_t*>(vc2_host_alloc(n));
    pinned_ = p_ != nullptr;
    if (!p_) p_ = static_cast<uint8_t*>(std::malloc(n ? n : 1));
    if (!p_) throw std::bad_alloc();
    n_ = n;
  }
  uint8_t* data() { return p_; }
  const uint8_t* data() const { return p_; }
  size_t size() const { return n_; }
  bool pinned() const { return pinned_; }
 private:
  HostBuf(const HostBuf&);
  HostBuf& operator=(const HostBuf&);
  void release() {
    if (!p_) return;
    if (pinned_) vc2_host_free(p_); else std::free(p_);
    p_ = nullptr;
  }
  uint8_t* p_;
  size_t n_;
  bool pinned_;
};

// blocking FIFO of small values (round indices)
template <class T>
class Channel {
 public:
  void push(const T& v) {
    { std::lock_guard<std::mutex> l(m_); q_.push_back(v); }
    cv_.notify_one();
  }
  T pop() {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !q_.empty(); });
    T v = q_.front();
    q_.pop_front();
    return v;
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<T> q_;
};

// Output of the pipelined modes.  A regular file is written with pwrite() at offsets the caller computes, large pieces by
// several threads at once (a single thread copies into the page cache at 3-4 GB/s, less than one GPU produces); anything else
// (a pipe, standard output) is written in order.
class PositionalWriter {
 public:
  PositionalWriter() : fd_(-1), seekable_(false), pos_(0), failed_(false) {}
  ~PositionalWriter() { close(); }
  // name "-" = standard output.  false when the file cannot be opened
  bool open(const char* name) {
    if (name[0] == '-' && name[1] == 0) { fd_ = 1; seekable_ = false; return true; }
    fd_ = ::open(name, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd_ < 0) return false;
    struct stat st;
    seekable_ = fstat(fd_, &st) == 0 && S_ISREG(st.st_mode);
    return true;
  }
  struct Piece { const void* p; size_t n; };
  // append the pieces in order; large ones are split over the threads of a seekable output
  void append(const std::vector<Piece>& pieces, int threads = 4) {
    if (!seekable_) {
      for (const Piece& q : pieces) writeAll(q.p, q.n);
      return;
    }
    struct Job { const uint8_t* p; size_t n; long long at; };
    std::vector<Job> jobs;
    const size_t chunk = 8u << 20;
    for (const Piece& q : pieces) {
      for (size_t o = 0; o < q.n; o += chunk) jobs.push_back({static_cast<const uint8_t*>(q.p) + o, std::min(chunk, q.n - o), pos_ + (long long)o});
      pos_ += (long long)q.n;
    }
    if (jobs.size() < 4 || threads < 2) { for (const Job& j : jobs) pwriteAll(j.p, j.n, j.at); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
      th.emplace_back([&, t]() { for (size_t i = t; i < jobs.size(); i += threads) pwriteAll(jobs[i].p, jobs[i].n, jobs[i].at); });
    for (auto& t : th) t.join();
  }
  bool ok() const { return !failed_; }
  void close() { if (fd_ > 1) ::close(fd_); fd_ = -1; }
 private:
  void writeAll(const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    while (n) {
      const ssize_t r = ::write(fd_, b, n);
      if (r <= 0) { failed_ = true; return; }
      b += r; n -= (size_t)r;
    }
  }
  void pwriteAll(const uint8_t* b, size_t n, long long at) {
    while (n) {
      const ssize_t r = ::pwrite(fd_, b, n, (off_t)at);
      if (r <= 0) { failed_ = true; return; }
      b += r; n -= (size_t)r; at += r;
    }
  }
  int fd_;
  bool seekable_;
  long long pos_;
  std::atomic<bool> failed_;
};

}  // namespace vc2cli
#endif
