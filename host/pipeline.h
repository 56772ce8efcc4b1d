// Small pieces shared by the drop-in command lines: page-locked buffers and a blocking queue for the
// reader -> codec -> writer pipeline (the reference tools are single threaded, EncodeStream.cpp:452-770).
#ifndef VC2_HOST_PIPELINE_H
#define VC2_HOST_PIPELINE_H
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <mutex>
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

}  // namespace vc2cli
#endif
