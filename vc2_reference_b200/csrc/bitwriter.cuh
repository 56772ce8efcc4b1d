// The HQ packer's bit writer, in a header of its own so that the host-side fuzz test (tests/test_bitwriter_host.py,
// tests/bitwriter_fuzz.cpp) compiles the very same code against a bit-by-bit model - no GPU needed.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define VC2_HD __host__ __device__ __forceinline__
#else
#define VC2_HD inline
#endif

namespace vc2 {

VC2_HD int bw_min(int a, int b) { return a < b ? a : b; }
VC2_HD void bw_store4(uint32_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {   // p is 16-byte aligned
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
#else
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

// The HQ packer's writer: the same bit accumulator, but finished words leave the thread FOUR AT A TIME.  The slice
// images of a warp's lanes are 4 KB apart, so every lane of a store instruction opens its own 32-byte sector: word
// stores cost one L2 request per word (tools/stream_probe.cu: 4.1 ms per 16.6 M warp-wide scattered word stores, most
// of the packer's time), 16-byte stores a quarter of that.  b0..b3 always hold the last four words pushed (b3 the
// newest); they are stored when the word cursor crosses a 16-byte boundary of the (16-byte aligned) slice image, and
// the up to three words behind the last boundary ("pending") exist only in registers until spill().
struct WideBitWriter {
  uint32_t* w;    // first word of the slice image (16-byte aligned: staging_words() is a multiple of four)
  unsigned long long acc;   // the low (bp & 31) bits are pending
  unsigned bp;    // bits written so far: ONE cursor, word index bp >> 5, so that "a word is full" and "four words are
                  // full" are comparisons on old ^ new cursor and the packer's marks are the cursor itself
  uint32_t b0, b1, b2, b3;
  VC2_HD void init(uint32_t* words) { w = words; acc = 0; bp = 0u; b0 = b1 = b2 = b3 = 0u; }
  VC2_HD int wc() const { return (int)(bp >> 5); }
  VC2_HD int pos() const { return (int)bp; }
  VC2_HD unsigned mark() const { return bp; }
  VC2_HD int unmark(unsigned m) const { return (int)m; }
  VC2_HD int pending() const { return (int)((bp >> 5) & 3u); }
  VC2_HD void put(uint32_t code, int nb) {   // nb in 0..32: at most one word boundary is crossed
    acc = (acc << nb) | code;
    const unsigned nbp = bp + (unsigned)nb;
    const unsigned x = nbp ^ bp;
    const bool full = x >= 32u;   // branch free: the lanes of a warp (one slice each) fill their words at different coefficients
    const uint32_t out = (uint32_t)(acc >> (nbp & 31u));
    b0 = full ? b1 : b0;
    b1 = full ? b2 : b1;
    b2 = full ? b3 : b2;
    b3 = full ? out : b3;
    bp = nbp;
    if (x >= 128u) bw_store4(w + (nbp >> 5) - 4, b0, b1, b2, b3);
  }
  // make the slice image in memory complete up to the cursor / fetch the pending words of a moved cursor back
  VC2_HD void spill() {
    const int r = pending();
    uint32_t* wp = w + wc();
    if (r >= 1) wp[-1] = b3;
    if (r >= 2) wp[-2] = b2;
    if (r >= 3) wp[-3] = b1;
  }
  VC2_HD void reload() {
    const int r = pending();
    const uint32_t* wp = w + wc();
    if (r >= 1) b3 = wp[-1];
    if (r >= 2) b2 = wp[-2];
    if (r >= 3) b1 = wp[-3];
  }
  VC2_HD void seek(int target) {   // see BitWriter::seek
    while ((int)bp < target) put(0u, bw_min(target - (int)bp, 32));
    if (target < (int)bp) {
      const int twc = target >> 5, tb = target & 31;
      if (twc == wc()) acc >>= ((int)(bp & 31u) - tb);
      else {
        spill();
        acc = (unsigned long long)w[twc] >> (32 - tb);
        bp = (unsigned)target;
        reload();
      }
      bp = (unsigned)target;
    }
  }
  VC2_HD void patch_byte(int bitpos, uint32_t value) {   // see BitWriter::patch_byte
    const int idx = bitpos >> 5;
    if (idx < wc()) {
      spill();
      w[idx] |= value << (24 - (bitpos & 31));
      reload();
    } else acc |= (unsigned long long)value << ((int)(bp & 31u) - (bitpos & 31) - 8);
  }
  VC2_HD void finish() {
    if (bp & 31u) put(0u, 32 - (int)(bp & 31u));
    spill();
  }
};

}  // namespace vc2
