// Host-side fuzz of the HQ packer's bit writer (vc2_reference_b200/csrc/bitwriter.cuh, compiled here as plain C++)
// against a bit-by-bit model of what the reference's stream does (VLC.cpp:119-185: MSB-first bits, vlc::bounded's
// truncation / zero padding = a cursor move, Slices.cpp:478-530: a length byte written after its component).
// The sequence of calls mirrors hq_pack_kernel: prefix bytes, qindex, then per component a zero length byte, codes,
// a seek to the component's byte length (backwards over dropped trailing codes, or forwards = zero padding) and the
// late patch of the length byte; finish.  Exit code 0 = every trial identical.   usage: bitwriter_fuzz [trials] [seed]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../vc2_reference_b200/csrc/bitwriter.cuh"

static uint64_t rng_state;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }

struct Model {
  std::vector<uint8_t> bits;
  void put(uint32_t code, int nb) { for (int i = nb - 1; i >= 0; --i) bits.push_back((code >> i) & 1u); }
  void seek(size_t target) { bits.resize(target, 0); }
  void patch_byte(size_t pos, uint32_t v) { for (int i = 0; i < 8; ++i) bits[pos + i] |= (v >> (7 - i)) & 1u; }
};

int main(int argc, char** argv) {
  const int trials = argc > 1 ? atoi(argv[1]) : 20000;
  rng_state = argc > 2 ? strtoull(argv[2], 0, 10) : 12345;
  alignas(16) static uint32_t words[1 << 14];
  for (int t = 0; t < trials; ++t) {
    memset(words, 0xA5, sizeof(words));   // stale staging contents must never show
    vc2::WideBitWriter W;
    Model M;
    W.init(words);
    const int prefix = rnd() % 4;
    for (int i = 0; i < prefix; ++i) { W.put(0u, 8); M.put(0u, 8); }
    const uint32_t qi = rnd() & 0xFF;
    W.put(qi, 8); M.put(qi, 8);
    const int style = rnd() % 4;   // 0 short codes, 1 mixed, 2 long codes, 3 mostly '1' codes (zero coefficients)
    for (int c = 0; c < 3; ++c) {
      const int len_pos = W.pos();
      if ((size_t)len_pos != M.bits.size()) { printf("trial %d: cursor mismatch\n", t); return 1; }
      W.put(0u, 8); M.put(0u, 8);
      const int data_start = W.pos();
      const int ncodes = rnd() % (style == 2 ? 40 : 300);
      int last = data_start;
      for (int i = 0; i < ncodes; ++i) {
        int nb;
        switch (style) {
          case 0: nb = 1 + rnd() % 6; break;
          case 1: nb = 1 + rnd() % 32; break;
          case 2: nb = 20 + rnd() % 13; break;
          default: nb = (rnd() % 8) ? 1 : 2 + rnd() % 10; break;
        }
        if (rnd() % 50 == 0) nb = 32;
        uint32_t code = rnd();
        if (nb < 32) code &= (1u << nb) - 1u;
        if (nb == 1) code = 1u;
        W.put(code, nb); M.put(code, nb);
        if (nb > 1 || rnd() % 16 == 0) last = W.pos();   // "behind the last non-zero coefficient"
      }
      const int scalar = 1 << (rnd() % 4);
      int L = (((last - data_start) + 7) / 8 + scalar - 1) / scalar * scalar;
      if (rnd() % 5 == 0) L += scalar * (rnd() % 40);      // HQ_CBR: the last component takes what is left
      if (rnd() % 11 == 0) L = 0;                           // everything dropped
      W.seek(data_start + 8 * L); M.seek((size_t)data_start + 8 * (size_t)L);
      W.patch_byte(len_pos, (uint32_t)(L / scalar) & 0xFFu); M.patch_byte(len_pos, (uint32_t)(L / scalar) & 0xFFu);
    }
    const size_t total = M.bits.size();
    W.finish();
    for (size_t i = 0; i < total; ++i) {
      const uint32_t bit = (words[i >> 5] >> (31 - (i & 31))) & 1u;
      if (bit != M.bits[i]) { printf("trial %d: bit %zu of %zu differs\n", t, i, total); return 1; }
    }
    for (size_t i = total; i < ((total + 31) & ~(size_t)31); ++i)
      if ((words[i >> 5] >> (31 - (i & 31))) & 1u) { printf("trial %d: padding bit %zu set\n", t, i); return 1; }
  }
  printf("%d trials identical\n", trials);
  return 0;
}
