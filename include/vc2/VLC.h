// vc2/VLC.h - the interleaved exp-Golomb value classes of src/Library/VLC.h:17-46 (UnsignedVLC, SignedVLC: a value as
// (number of bits, code word) and back, VLC.cpp:21-94) and an MSB-first bit buffer for the places host code needs them (stream
// headers; the slice payloads are coded on the GPU).  Header only, own code; checked against the compiled reference in
// tests/test_host_helpers.py.  The reference's stream-state manipulators (vlc::bounded, flush, align, VLC.h:103-143) have no
// counterpart here: the bound of a slice component is an argument of the device coders (csrc/slices.cu, csrc/bitwriter.cuh).
#ifndef VC2_VLC_H
#define VC2_VLC_H
#include <cstddef>
#include <cstdint>
#include <vector>

namespace vc2 {

class UnsignedVLC {
 public:
  UnsignedVLC() : nBits_(1), bits_(1) {}
  // value -> code: the bits of value + 1 below its leading one, each behind a 0 "follow" bit, then the 1 that ends the code
  explicit UnsignedVLC(unsigned int value) : nBits_(1), bits_(1) {
    const unsigned int m = value + 1u;
    int k = 0;
    while ((m >> (k + 1)) != 0u) ++k;          // m has k bits below its leading one
    unsigned int code = 0;
    for (int i = k - 1; i >= 0; --i) code = (code << 2) | ((m >> i) & 1u);
    bits_ = (code << 1) | 1u;
    nBits_ = 2u * (unsigned)k + 1u;
  }
  UnsignedVLC(unsigned int numOfBits, unsigned int code) : nBits_(numOfBits), bits_(code) {}
  unsigned int numOfBits() const { return nBits_; }
  unsigned int code() const { return bits_; }
  // code -> value
  operator unsigned int() const {
    unsigned int m = 1;
    for (int pos = (int)nBits_ - 1; pos >= 2; pos -= 2) m = (m << 1) | ((bits_ >> (pos - 1)) & 1u);
    return m - 1u;
  }
 private:
  unsigned int nBits_, bits_;
};

class SignedVLC {
 public:
  SignedVLC() : nBits_(1), bits_(1) {}
  // the unsigned code of |value|, and a sign bit (1 = negative) behind it when the value is not zero
  explicit SignedVLC(int value) : nBits_(1), bits_(1) {
    if (value != 0) {
      const UnsignedVLC u((unsigned int)(value < 0 ? -value : value));
      bits_ = (u.code() << 1) | (value < 0 ? 1u : 0u);
      nBits_ = u.numOfBits() + 1u;
    }
  }
  SignedVLC(unsigned int numOfBits, unsigned int code) : nBits_(numOfBits), bits_(code) {}
  unsigned int numOfBits() const { return nBits_; }
  unsigned int code() const { return bits_; }
  operator int() const {
    if (nBits_ <= 1) return 0;
    const int mag = (int)(unsigned int)UnsignedVLC(nBits_ - 1, bits_ >> 1);
    return (bits_ & 1u) ? -mag : mag;
  }
 private:
  unsigned int nBits_, bits_;
};

// MSB-first bit buffer (putBits / getBit of VLC.cpp:96-150 on a byte vector instead of a stream)
class BitBuffer {
 public:
  BitBuffer() : wpos_(0), rpos_(0) {}
  explicit BitBuffer(const std::vector<uint8_t>& bytes) : v_(bytes), wpos_(8 * bytes.size()), rpos_(0) {}
  void put(unsigned int nBits, unsigned int code) {
    for (int i = (int)nBits - 1; i >= 0; --i) {
      if ((wpos_ & 7) == 0) v_.push_back(0);
      if ((code >> i) & 1u) v_.back() |= (uint8_t)(0x80u >> (wpos_ & 7));
      ++wpos_;
    }
  }
  void put(const UnsignedVLC& c) { put(c.numOfBits(), c.code()); }
  void put(const SignedVLC& c) { put(c.numOfBits(), c.code()); }
  void align() { wpos_ = (wpos_ + 7) & ~(size_t)7; rpos_ = (rpos_ + 7) & ~(size_t)7; }
  bool getBit() {                                   // past the end: ones, like a bounded read (VLC.cpp:182-185)
    if (rpos_ >= 8 * v_.size()) { ++rpos_; return true; }
    const bool b = (v_[rpos_ >> 3] >> (7 - (rpos_ & 7))) & 1u;
    ++rpos_;
    return b;
  }
  UnsignedVLC getUnsigned() {
    unsigned int n = 0, code = 0;
    while (!getBit()) { code = (code << 2) | (getBit() ? 1u : 0u); n += 2; }
    return UnsignedVLC(n + 1, (code << 1) | 1u);
  }
  SignedVLC getSigned() {
    const UnsignedVLC u = getUnsigned();
    if (u.numOfBits() == 1) return SignedVLC(1, 1);
    return SignedVLC(u.numOfBits() + 1, (u.code() << 1) | (getBit() ? 1u : 0u));
  }
  const std::vector<uint8_t>& bytes() const { return v_; }
  size_t bitsWritten() const { return wpos_; }
 private:
  std::vector<uint8_t> v_;
  size_t wpos_, rpos_;
};

}  // namespace vc2
#endif
