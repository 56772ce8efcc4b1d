// Slice coding kernels (see slices.cuh for the reference line citations).
#include "slices.cuh"
#include "bitwriter.cuh"

namespace vc2 {

__constant__ QuantTables c_qt;


// ---- SignedVLC code tables (VLC.cpp:21-52, 78-85, 283-317), built on the host, staged into shared memory by
// the slice coders: the arithmetic pipe is the bottleneck of both coders, a table look-up runs on the LSU.
//   d_enc_lut[2 * |v| + (v < 0)] = (code << 5) | bits                 for |v| < ENC_LUT_MAG
//   d_dec_lut[top DEC_LUT_BITS bits of the window] = value * 16 + bits, or 0 when the code is longer than that
constexpr int ENC_LUT_MAG = 256;
constexpr int DEC_LUT_BITS = 12;
__device__ uint32_t d_enc_lut[2 * ENC_LUT_MAG];
__device__ int16_t d_dec_lut[1 << DEC_LUT_BITS];
// the same for the narrow parser: (t << 4) | bits with t = 2 * |value| + (value < 0), the word the narrow block stores
__device__ uint16_t d_dec_lut_sm[1 << DEC_LUT_BITS];

cudaError_t upload_quant_tables(const QuantTables& t) {
  cudaError_t e = cudaMemcpyToSymbol(c_qt, &t, sizeof(t));
  if (e != cudaSuccess) return e;
  static uint32_t enc[2 * ENC_LUT_MAG];
  static int16_t dec[1 << DEC_LUT_BITS];
  static uint16_t dec_sm[1 << DEC_LUT_BITS];
  for (int mag = 0; mag < ENC_LUT_MAG; ++mag)
    for (int neg = 0; neg < 2; ++neg) {
      uint32_t code = 1, nb = 1;
      if (mag) {
        const uint32_t m = (uint32_t)mag + 1u;
        int k = 0;
        while ((m >> (k + 1)) != 0) ++k;
        code = 0;
        for (int i = k - 1; i >= 0; --i) code = (code << 2) | ((m >> i) & 1u);   // 0 b(k-1) 0 b(k-2) ... 0 b0
        code = (code << 2) | 2u | (uint32_t)neg;                                 // 1 s
        nb = 2 * k + 2;
      }
      enc[2 * mag + neg] = (code << 5) | nb;
    }
  for (int i = 0; i < (1 << DEC_LUT_BITS); ++i) {
    int pos = DEC_LUT_BITS - 1, m = 1, entry = 0;
    for (int k = 0; pos >= 0; ++k) {
      const int follow = (i >> pos) & 1;
      --pos;
      if (follow) {   // stop bit: magnitude complete, a sign bit follows when it is not zero
        int v = m - 1, len = 2 * k + 1;
        if (v) {
          if (pos < 0) break;
          if ((i >> pos) & 1) v = -v;
          ++len;
        }
        entry = v * 16 + len;
        break;
      }
      if (pos < 0) break;
      m = (m << 1) | ((i >> pos) & 1);
      --pos;
    }
    dec[i] = (int16_t)entry;
    const int val = entry >> 4, len = entry & 15;   // arithmetic shift: the value is signed
    dec_sm[i] = entry ? (uint16_t)(((2 * (val < 0 ? -val : val) + (val < 0 ? 1 : 0)) << 4) | len) : (uint16_t)0;
  }
  e = cudaMemcpyToSymbol(d_enc_lut, enc, sizeof(enc));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(d_dec_lut_sm, dec_sm, sizeof(dec_sm));
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(d_dec_lut, dec, sizeof(dec));
}

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

// ---- exact dead-zone quantiser (Quantisation.cpp:69-76): sign(v) * ((|v| << 2) / qf) --------------
// The truncating division by the table constant uses the round-up multiply-shift of Granlund &
// Montgomery, exact for every 32-bit dividend:  t = mulhi(m', a);  q = (t + ((a - t) >> 1)) >> (l - 1)
struct QParam {
  uint32_t qf, qo, qm, ql, qm31, ql31, qmul16, qsh16;
};
__device__ __forceinline__ QParam qparam(int q) {
  QParam r;
  q = min(max(q, 0), 127);
  r.qf = c_qt.qf[q];
  r.qo = c_qt.qo[q];
  r.qm = c_qt.qm[q];
  r.ql = c_qt.ql[q];
  r.qm31 = c_qt.qm31[q];
  r.ql31 = c_qt.ql31[q];
  r.qmul16 = c_qt.qmul16[q];
  r.qsh16 = c_qt.qsh16[q];
  return r;
}
__device__ __forceinline__ uint32_t udiv_magic(uint32_t a, uint32_t m, uint32_t l) {
  const uint32_t t = __umulhi(m, a);
  return (t + ((a - t) >> 1)) >> (l - 1);
}
__device__ __forceinline__ int quant_one(int v, uint32_t qm, uint32_t ql) {
  const uint32_t a = (uint32_t)abs(v) << 2;
  const int q = (int)udiv_magic(a, qm, ql);
  return v < 0 ? -q : q;
}
// inverse quantiser (Quantisation.cpp:86-95)
__device__ __forceinline__ int scale_one(int v, uint32_t qf, uint32_t qo) {
  if (v == 0) return 0;
  const uint32_t m = ((uint32_t)abs(v) * qf + qo + 2u) >> 2;
  return v < 0 ? -(int)m : (int)m;
}
// SignedVLC length (VLC.cpp:78-85): 1 for zero, else 2*floor(log2(|v|+1)) + 2
__device__ __forceinline__ int vlc_bits(int v) {
  if (v == 0) return 1;
  const int k = 31 - __clz(abs(v) + 1);
  return 2 * k + 2;
}
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__device__ __forceinline__ uint32_t compress16(uint32_t x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0x0000FFFFu;
  return x;
}
// SignedVLC code word (VLC.cpp:21-52, 78-85), valid for |v| < 65535 (reference VLC is 32-bit)
__device__ __forceinline__ void vlc_code(int v, uint32_t& code, int& nb) {
  if (v == 0) { code = 1; nb = 1; return; }
  const uint32_t m = (uint32_t)abs(v) + 1u;
  const int k = 31 - __clz(m);
  const uint32_t low = m & ((1u << k) - 1u);
  code = (spread16(low) << 2) | 2u | (v < 0 ? 1u : 0u);
  nb = 2 * k + 2;
}

// quantiser parameters of one band for one slice (per lane: the slices of a warp may use different indices)
struct BandP {
  uint32_t qm, ql, qf, qo;   // one-multiply magic (exact below 2^31) and its shift, quant_factor, quant_offset + 2
  uint32_t mul16, sh16;      // full-rate form for |v| < VC2_NARROW_FAST_MAX (mul16 == 0: not available)
};
// adjusted index max(q - qmatrix[b], 0) (Quantisation.cpp:16-20); bad when the reference would throw (:60-63)
__device__ __forceinline__ BandP band_params(int q, int qmat, bool& bad) {
  const int aq = max(q - qmat, 0);
  if (aq > 119) bad = true;
  const QParam p = qparam(aq);
  BandP r;
  r.qm = p.qm31; r.ql = p.ql31; r.qf = p.qf; r.qo = p.qo + 2u;
  r.mul16 = p.qmul16; r.sh16 = p.qsh16;
  return r;
}
__device__ __forceinline__ int quant_band(int v, const BandP& bp) {
  const int q = (int)(__umulhi(bp.qm, (uint32_t)abs(v) << 2) >> bp.ql);
  return v < 0 ? -q : q;
}
__device__ __forceinline__ int scale_band(int v, const BandP& bp) {   // bp.qo already holds quant_offset + 2
  const uint32_t a = (uint32_t)abs(v);
  const uint32_t m = (a * bp.qf + (a ? bp.qo : 0u)) >> 2;   // 0 -> 0 without a branch (lanes differ)
  return v < 0 ? -(int)m : (int)m;
}

// ------------------------------------------------------------------------------------------
// Walk the NC-long coefficient list of one slice component in coding order, one 16-byte piece
// (four coefficients) per load.  All lanes of a warp are at the same list position, so the band
// bookkeeping is warp uniform; the per-band quantiser parameters are per lane.
//   src  : this lane's first piece of the component (pieces are 32 int4 apart, see vc2_common.cuh)
//   op(v, bp) is called once per coefficient, in coding order
// ------------------------------------------------------------------------------------------
#ifndef VC2_WALK_UNROLL
#define VC2_WALK_UNROLL
#endif
template <class Op>
__device__ __forceinline__ void walk_component(const int4* __restrict__ base, size_t first, const SliceGeom& g, int c, int q, bool& badq, Op& op) {
  const int n = g.band_start[c][g.nbands];
  int k = 0, b = 0, bend = g.band_start[c][1];
  BandP bp = band_params(q, g.qmatrix[0], badq);
  // three pieces in flight: the loads of a thread are 512 bytes apart (one piece of each of the 32 slices of the
  // group in between), every one a fresh line, and one piece of work is too short to cover an HBM round trip
  const int np = n >> 2;
  int4 n0 = __ldg(base + first), n1 = n0, n2 = n0;
  if (np > 1) n1 = __ldg(base + first + 32);
  if (np > 2) n2 = __ldg(base + first + 64);
  int piece = 0;
  while (piece < np) {
    while (k == bend) {   // coefficient k exists (piece < np), so does its band
      ++b;
      bend = g.band_start[c][b + 1];
      bp = band_params(q, g.qmatrix[b], badq);
    }
    // the whole pieces inside the current band are one inner loop: band bookkeeping (constant-bank look-ups, the
    // quantiser parameters) stays outside it
    const int run = min((bend - k) >> 2, np - piece);
    if (run > 0) {
      k += 4 * run;
      size_t pf = first + (size_t)(piece + 3) * 32;
      int ahead = np - (piece + 3);   // pieces that can still be prefetched
      piece += run;
      VC2_WALK_UNROLL
      for (int i = 0; i < run; ++i) {
        const int4 v4 = n0;
        n0 = n1; n1 = n2;
        if (ahead > 0) n2 = __ldg(base + pf);
        pf += 32; --ahead;
        op.pair(v4.x, v4.y, bp);
        op.pair(v4.z, v4.w, bp);
      }
    } else {
      const int4 v4 = n0;
      n0 = n1; n1 = n2;
      if (piece + 3 < np) n2 = __ldg(base + first + (size_t)(piece + 3) * 32);
      ++piece;
      const int v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        while (k == bend) {
          ++b;
          bend = g.band_start[c][b + 1];
          bp = band_params(q, g.qmatrix[b], badq);
        }
        op(v[e], bp);
        ++k;
      }
    }
  }
}

// SignedVLC of a value of magnitude mag (sign neg): table look-up for |v| < ENC_LUT_MAG, else computed with
// m = |v| + 1 clamped to the 32-bit code domain (VLC.h:27-28; bigor collects the m that leave it)
__device__ __forceinline__ void vlc_of(const uint32_t* __restrict__ lut, uint32_t mag, bool neg, unsigned& bigor, uint32_t& code, int& nb) {
  if (mag < (uint32_t)ENC_LUT_MAG) {
    const uint32_t e = lut[2u * mag + (neg ? 1u : 0u)];
    nb = (int)(e & 31u);
    code = e >> 5;
  } else {
    uint32_t m = mag + 1u;
    bigor |= m;
    m = min(m, 65535u);
    const int k = 31 - __clz(m);
    nb = 2 * k + 2;
    code = (spread16(m ^ (1u << k)) << 2) | 2u | (neg ? 1u : 0u);
  }
}
__device__ __forceinline__ uint32_t quant_mag(int v, const BandP& bp) {   // |quant(v, q)|
  return __umulhi(bp.qm, (uint32_t)abs(v) << 2) >> bp.ql;
}
// stage a table of n 32-bit words from global into shared memory (whole CTA; caller synchronises)
__device__ __forceinline__ void stage_table(uint32_t* dst, const uint32_t* src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// MSB-first bit writer into this slice's staging words (one thread owns the whole slice)
struct BitWriter {
  uint32_t* w;    // first word of the slice image
  uint32_t* wp;   // next word to write
  unsigned long long acc;   // the low nacc bits are pending
  int nacc;
  __device__ __forceinline__ void init(uint32_t* words) { w = wp = words; acc = 0; nacc = 0; }
  __device__ __forceinline__ int wc() const { return (int)(wp - w); }
  __device__ __forceinline__ int pos() const { return 32 * wc() + nacc; }
  // cheap monotonic cursor (one multiply-add): pos() + 8 * (low address bits of w); compare / subtract only
  __device__ __forceinline__ unsigned mark() const { return 8u * (unsigned)reinterpret_cast<uintptr_t>(wp) + (unsigned)nacc; }
  __device__ __forceinline__ int unmark(unsigned m) const { return (int)(m - 8u * (unsigned)reinterpret_cast<uintptr_t>(w)); }
  __device__ __forceinline__ void put(uint32_t code, int nb) {   // nb in 0..32
    acc = (acc << nb) | code;
    nacc += nb;
    // branch-free flush: the lanes of a warp (one slice each) reach 32 pending bits at different coefficients
    const bool full = nacc >= 32;
    const uint32_t out = (uint32_t)(acc >> ((nacc - 32) & 31));
    if (full) *wp = out;
    wp += full ? 1 : 0;
    nacc -= full ? 32 : 0;
  }
  // move the cursor to absolute bit position target: pad with zeros, or drop what was written beyond it
  // (only the '1' codes of trailing zero coefficients can be dropped, VLC.cpp:151-155)
  __device__ __forceinline__ void seek(int target) {
    int cur = pos();
    while (cur < target) {
      const int t = min(target - cur, 32);
      put(0u, t);
      cur += t;
    }
    if (target < cur) {
      const int twc = target >> 5, tb = target & 31;
      if (twc == wc()) acc >>= (nacc - tb);
      else { acc = (unsigned long long)w[twc] >> (32 - tb); wp = w + twc; }
      nacc = tb;
    }
  }
  // overwrite the (zero) byte at byte-aligned bit position bitpos < pos()
  __device__ __forceinline__ void patch_byte(int bitpos, uint32_t value) {
    const int idx = bitpos >> 5;
    if (idx < wc()) w[idx] |= value << (24 - (bitpos & 31));
    else acc |= (unsigned long long)value << (nacc - (bitpos - 32 * wc()) - 8);
  }
  __device__ __forceinline__ void finish() {
    if (nacc > 0) { *wp++ = (uint32_t)(acc << (32 - nacc)); nacc = 0; }
  }
};

struct CountOp {      // component_slice_bytes' bit count of quantise(v) (rate-control probe)
  const uint32_t* lut;
  int bits, last;
  unsigned bigor;
  const uint8_t* blut;   // SignedVLC(v).numOfBits() for |v| < ENC_LUT_MAG as bytes (shared memory): the probes only need the length
  __device__ __forceinline__ void operator()(int v, const BandP& bp) { count(quant_mag(v, bp)); }
  __device__ __forceinline__ void count(uint32_t mag) {
    if (mag < (uint32_t)ENC_LUT_MAG) bits += (int)blut[mag];
    else {
      const uint32_t m = mag + 1u;
      bigor |= m;
      bits += 2 * (31 - __clz(min(m, 65535u))) + 2;
    }
    if (mag) last = bits;
  }
  // The probes of the rate control quantise every coefficient seven times.  Nearly all pairs take the first branch: both
  // coefficients small enough for the full-rate multiply (instead of the multiply-high) and both codes in the table - one test
  // for the pair, then straight-line code: multiply, shift, byte look-up, add, and the move that remembers the last non-zero.
  __device__ __forceinline__ void pair(int v0, int v1, const BandP& bp) { mags((uint32_t)abs(v0), (uint32_t)abs(v1), bp); }
  // the same on magnitudes (the bit count does not depend on the sign)
  __device__ __forceinline__ void mag1(uint32_t a, const BandP& bp) { count(__umulhi(bp.qm, a << 2) >> bp.ql); }
  __device__ __forceinline__ void mags(uint32_t a0, uint32_t a1, const BandP& bp) {
    if (bp.mul16 != 0u && (a0 | a1) < (uint32_t)VC2_NARROW_FAST_MAX) {
      const uint32_t m0 = (a0 * bp.mul16) >> bp.sh16, m1 = (a1 * bp.mul16) >> bp.sh16;
      if ((m0 | m1) < (uint32_t)ENC_LUT_MAG) {
        bits += (int)blut[m0];
        last = m0 ? bits : last;
        bits += (int)blut[m1];
        last = m1 ? bits : last;
      } else {
        count(m0);
        count(m1);
      }
    } else {
      mag1(a0, bp);
      mag1(a1, bp);
    }
  }
};
struct SseOp {        // yss_for_slice (Quantisation.cpp:627-642): product in int, sum in long long
  long long acc;
  __device__ __forceinline__ void operator()(int v, const BandP& bp) {
    const int d = v - scale_band(quant_band(v, bp), bp);
    acc += (long long)(int)((unsigned)d * (unsigned)d);
  }
  __device__ __forceinline__ void pair(int v0, int v1, const BandP& bp) { (*this)(v0, bp); (*this)(v1, bp); }
  // on magnitudes: v - scale(quant(v)) only changes sign with v, its square does not
  __device__ __forceinline__ void mag1(uint32_t a, const BandP& bp) {
    const uint32_t m = __umulhi(bp.qm, a << 2) >> bp.ql;
    const uint32_t r = (m * bp.qf + (m ? bp.qo : 0u)) >> 2;
    const unsigned d = a - r;
    acc += (long long)(int)(d * d);
  }
  __device__ __forceinline__ void mags(uint32_t a0, uint32_t a1, const BandP& bp) { mag1(a0, bp); mag1(a1, bp); }
};
template <bool QUANT, class Writer>
struct EmitOp {
  const uint32_t* lut;
  Writer* W;
  unsigned last;   // BitWriter::mark() behind the last non-zero coefficient
  unsigned bigor;
  __device__ __forceinline__ void operator()(int v, const BandP& bp) {
    const uint32_t mag = QUANT ? quant_mag(v, bp) : (uint32_t)abs(v);
    uint32_t code;
    int nb;
    vlc_of(lut, mag, v < 0, bigor, code, nb);
    W->put(code, nb);
    last = mag ? W->mark() : last;
  }
  // two coefficients of the same band: when both codes come from the table and are at most 16 bits long they go
  // into the accumulator as ONE append (one shift, one flush test)
  __device__ __forceinline__ void pair(int v0, int v1, const BandP& bp) {
    const uint32_t m0 = QUANT ? quant_mag(v0, bp) : (uint32_t)abs(v0), m1 = QUANT ? quant_mag(v1, bp) : (uint32_t)abs(v1);
    if ((m0 | m1) < 128u) {
      const uint32_t e0 = lut[2u * m0 + (v0 < 0 ? 1u : 0u)], e1 = lut[2u * m1 + (v1 < 0 ? 1u : 0u)];
      const int nb0 = (int)(e0 & 31u), nb1 = (int)(e1 & 31u);
      const unsigned mid = W->mark() + (unsigned)nb0;        // the cursor behind the first code (mark() is linear in the bit position)
      W->put(((e0 >> 5) << nb1) | (e1 >> 5), nb0 + nb1);
      last = m1 ? W->mark() : (m0 ? mid : last);
    } else {
      (*this)(v0, bp);
      (*this)(v1, bp);
    }
  }
};

// component_slice_bytes (Slices.cpp:114-118): whole scalar units; err when the length byte overflows
__device__ __forceinline__ int scaled_bytes(int count, int scalar, bool& too_big) {
  const int units = ((count + 7) / 8 + scalar - 1) / scalar;
  if (units > 0xFF) too_big = true;
  return units * scalar;
}

// ---- rate control of one slice: literal replay of quantIndicesCBR (EncodeStream.cpp:85-122).  walk(c, q, badq, op) feeds the
// coefficients of component c, in coding order, to op.  Returns the index (0 when the search died; flags say why).
template <class Walk>
__device__ __forceinline__ int search_slice(const PackParams& p, const SliceGeom& g, int s, const uint32_t* s_enc, const uint8_t* s_bits,
                                            unsigned& flags, Walk& walk) {
  const int avail = p.slice_bytes[s] - 4;
  int trialQ = 63, q = 127, delta = 64;
  bool dead = false;
  while (delta > 0) {
    delta >>= 1;
    int need = 0;
    bool too_big = false, badq = false;
    for (int c = 0; c < 3; ++c) {
      CountOp op = {s_enc, 0, 0, 0u, s_bits};
      walk(c, trialQ, badq, op);
      need += scaled_bytes(op.last, g.scalar, too_big);
    }
    if (badq) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; break; }
    if (too_big) { flags |= VC2_FLAG_SCALAR_TOO_SMALL | VC2_FLAG_SEARCH_PHASE; dead = true; break; }
    if (need <= avail) { if (trialQ < q) q = trialQ; trialQ -= delta; }
    else trialQ += delta;
  }
  if (!dead) {
    // "try a few higher quantisers": keep going while the luma squared error strictly drops
    trialQ = q;
    bool badq = false;
    SseOp prev = {0};
    walk(0, trialQ, badq, prev);
    while (!badq) {
      ++trialQ;
      SseOp cur = {0};
      walk(0, trialQ, badq, cur);
      if (badq) break;
      const long long d = cur.acc - prev.acc;
      prev = cur;
      if (!(d < 0)) break;
    }
    if (badq) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; }
    q = trialQ - 1;
  }
  return dead ? 0 : q;
}

// The coefficient list of a slice component as 16-bit MAGNITUDES in shared memory (hq_search_kernel): piece i of this lane at
// sm[i * stride], four magnitudes per piece.  Same band bookkeeping as walk_component; the ops take magnitudes.
template <class Op>
__device__ __forceinline__ void walk_component_mag(const uint2* sm, int stride, const SliceGeom& g, int c, int q, bool& badq, Op& op) {
  const int np = g.band_start[c][g.nbands] >> 2;
  int k = 0, b = 0, bend = g.band_start[c][1];
  BandP bp = band_params(q, g.qmatrix[0], badq);
  int piece = 0;
  while (piece < np) {
    while (k == bend) {
      ++b;
      bend = g.band_start[c][b + 1];
      bp = band_params(q, g.qmatrix[b], badq);
    }
    const int run = min((bend - k) >> 2, np - piece);
    if (run > 0) {
      k += 4 * run;
      const uint2* q2 = sm + (size_t)piece * stride;
      piece += run;
#pragma unroll 2
      for (int i = 0; i < run; ++i) {
        const uint2 w = *q2;
        q2 += stride;
        op.mags(w.x & 0xFFFFu, w.x >> 16, bp);
        op.mags(w.y & 0xFFFFu, w.y >> 16, bp);
      }
    } else {
      const uint2 w = sm[(size_t)piece * stride];
      ++piece;
      const uint32_t a[4] = {w.x & 0xFFFFu, w.x >> 16, w.y & 0xFFFFu, w.y >> 16};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        while (k == bend) {
          ++b;
          bend = g.band_start[c][b + 1];
          bp = band_params(q, g.qmatrix[b], badq);
        }
        op.mag1(a[e], bp);
        ++k;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// HQ_CBR rate control with the slice in shared memory (optional, see search_in_smem()).  The seven probes and the squared-error walk of a slice read its
// coefficients ten times; hq_pack_kernel streams them from DRAM every time (7.2 x the block, and the load latency is the first
// stall reason, ncu).  Here a CTA first copies the magnitudes of its slices - neither the bit count nor the squared error
// depends on the sign - as 16-bit words into shared memory, [piece][lane] so that the lanes of a warp read consecutive 8-byte
// words, and every walk runs from there.  A slice with a magnitude beyond 16 bits keeps the walks in global memory.
// blockDim.x slices per CTA, dynamic shared memory = blockDim.x * NC * 2 bytes.
// ------------------------------------------------------------------------------------------
__global__ void hq_search_kernel(const PackParams p) {
  extern __shared__ uint2 s_mag[];
  __shared__ uint8_t s_bits[ENC_LUT_MAG];
  for (int i = threadIdx.x; i < ENC_LUT_MAG; i += blockDim.x) s_bits[i] = (uint8_t)(d_enc_lut[2 * i] & 31u);
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y, spc = blockDim.x;
  const int s = blockIdx.x * spc + threadIdx.x;
  const int nc4 = g.comp_start[3] >> 2;
  const int4* base = reinterpret_cast<const int4*>(p.coef + (long long)pic * g.coef_pic_stride);
  const size_t src = (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  bool wide = false;
  if (s < nslices) {
    uint2* dst = s_mag + threadIdx.x;
    const int4* q4 = base + src;
#pragma unroll 4
    for (int i = 0; i < nc4; ++i) {
      const int4 v = __ldg(q4 + (size_t)i * 32);
      const uint32_t a0 = (uint32_t)abs(v.x), a1 = (uint32_t)abs(v.y), a2 = (uint32_t)abs(v.z), a3 = (uint32_t)abs(v.w);
      wide |= ((a0 | a1 | a2 | a3) >> 16) != 0u;
      dst[(size_t)i * spc] = make_uint2(a0 | (a1 << 16), a2 | (a3 << 16));
    }
  }
  __syncthreads();
  if (s >= nslices) return;
  const long long sidx = (long long)pic * nslices + s;
  unsigned flags = 0;
  int qi;
  if (!wide) {
    auto walk = [&](int c, int q, bool& badq, auto& op) { walk_component_mag(s_mag + (size_t)(g.comp_start[c] >> 2) * spc + threadIdx.x, spc, g, c, q, badq, op); };
    qi = search_slice(p, g, s, nullptr, s_bits, flags, walk);
  } else {
    auto walk = [&](int c, int q, bool& badq, auto& op) { walk_component(base, src + (size_t)(g.comp_start[c] >> 2) * 32, g, c, q, badq, op); };
    qi = search_slice(p, g, s, nullptr, s_bits, flags, walk);
  }
  p.qidx[sidx] = qi;
  p.err_flags[sidx] = flags;
}

// ------------------------------------------------------------------------------------------
// HQ_CBR rate control, ONE WARP PER SLICE, the slice in registers (quantIndicesCBR, EncodeStream.cpp:73-125); optional, see
// search_in_registers().
// The seven probes and the squared-error walk read a slice ten times; with a thread per slice that is ten passes over
// memory (hq_pack_kernel).  Here lane l of the warp loads pieces l, l + 32, ... of the slice's coefficient list ONCE - R
// pieces of four magnitudes per lane; neither the bit count nor the squared error depends on the sign - and every probe is
// arithmetic on those registers plus two warp reductions per component:
//   bits of a component up to its last non-zero coefficient = (sum of all its code lengths) - (coefficients behind the last
//   non-zero one), since a zero costs exactly one bit; the sum is a REDUX add, the position of the last non-zero a REDUX max.
// The quantiser parameters of the bands for the probed index sit in a small per-warp table in shared memory (one lane per
// band fills it), the band of every coefficient is worked out once.  The search itself is warp uniform.
// Eight slices per CTA: the eight warps read the same 128-byte lines of the group-interleaved block.
// ------------------------------------------------------------------------------------------
struct WarpBand {                 // per warp: the parameters of every band for the index being probed, as three 8-byte tables
  uint2 fast[VC2_MAX_BANDS];      // (qmul16, qsh16): |q| = (|v| * mul) >> sh for small |v|
  uint2 slow[VC2_MAX_BANDS];      // (qm31, ql31):    |q| = mulhi(4 |v|, m) >> l
  uint2 back[VC2_MAX_BANDS];      // (quant_factor, quant_offset + 2)
};

template <int R>
__global__ void __launch_bounds__(256) hq_search_warp_kernel(const PackParams p) {
  __shared__ uint8_t s_bits[ENC_LUT_MAG];
  __shared__ WarpBand s_band[8];
  // the quantiser tables by index, out of the constant bank: the lanes of set_index() look up different indices, which the
  // constant cache serves one distinct address at a time
  __shared__ uint2 s_qfast[128], s_qslow[128], s_qback[128];
  for (int i = threadIdx.x; i < ENC_LUT_MAG; i += blockDim.x) s_bits[i] = (uint8_t)(d_enc_lut[2 * i] & 31u);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    s_qfast[i] = make_uint2(c_qt.qmul16[i], c_qt.qsh16[i]);
    s_qslow[i] = make_uint2(c_qt.qm31[i], c_qt.ql31[i]);
    s_qback[i] = make_uint2(c_qt.qf[i], c_qt.qo[i] + 2u);
  }
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * 8 + warp;
  if (s >= nslices) return;   // whole warps leave
  const int nc4 = g.comp_start[3] >> 2;
  const int4* base = reinterpret_cast<const int4*>(p.coef + (long long)pic * g.coef_pic_stride) + (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  // ---- the slice: magnitudes, component, position inside the component and band of every coefficient this lane holds
  uint32_t a[R][4];
  int comp[R], kc0[R];      // component of the piece (-1: no such piece), index of its first coefficient inside the component
  uint32_t bands[R];        // the bands of its four coefficients, 8 bits each
  bool fastmag = true;      // every magnitude is below VC2_NARROW_FAST_MAX (then the full-rate multiply is exact)
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int piece = j * 32 + lane;
    comp[j] = -1; kc0[j] = 0; bands[j] = 0;
    a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0u;
    if (piece < nc4) {
      const int4 v = __ldg(base + (size_t)piece * 32);
      a[j][0] = (uint32_t)abs(v.x); a[j][1] = (uint32_t)abs(v.y); a[j][2] = (uint32_t)abs(v.z); a[j][3] = (uint32_t)abs(v.w);
      fastmag = fastmag && (a[j][0] | a[j][1] | a[j][2] | a[j][3]) < (uint32_t)VC2_NARROW_FAST_MAX;
      const int k = 4 * piece;
      const int c = k >= g.comp_start[2] ? 2 : (k >= g.comp_start[1] ? 1 : 0);
      comp[j] = c;
      kc0[j] = k - g.comp_start[c];
      int b = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        while (kc0[j] + e >= g.band_start[c][b + 1]) ++b;   // bands are in coding order: b only grows
        bands[j] |= (uint32_t)b << (8 * e);
      }
    }
  }
  fastmag = __all_sync(FULL, fastmag);
  // which lanes hold component c in slot j (warp uniform; zero: nobody)
  unsigned cmask[R][3];
#pragma unroll
  for (int j = 0; j < R; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) cmask[j][c] = __ballot_sync(FULL, comp[j] == c);
  WarpBand& tab = s_band[warp];
  int my_qmat = 0;                       // lane b: the quantisation matrix entry of band b
  if (lane < g.nbands) my_qmat = g.qmatrix[lane];
  // the quantiser parameters of every band for index q; false when the reference would throw (Quantisation.cpp:60-63)
  auto set_index = [&](int q, bool& allfast) -> bool {
    bool bad = false, fast = true;
    __syncwarp();
    if (lane < g.nbands) {
      const int aq = max(q - my_qmat, 0);
      bad = aq > 119;
      const int i = min(aq, 127);
      const uint2 f = s_qfast[i];
      fast = f.x != 0u;
      tab.fast[lane] = f; tab.slow[lane] = s_qslow[i]; tab.back[lane] = s_qback[i];
    }
    __syncwarp();
    allfast = fastmag && __all_sync(FULL, fast);
    return !__any_sync(FULL, bad);
  };
  // bytes the three components need at the index in the table (component_slice_bytes, Slices.cpp:97-119)
  auto need_bytes = [&](bool allfast, bool& too_big) -> int {
    int need = 0;
    int total[3] = {0, 0, 0}, lastpos[3] = {-1, -1, -1};
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (__all_sync(FULL, comp[j] < 0)) continue;
      uint32_t m[4];
      if (allfast) {   // warp uniform
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint2 w = tab.fast[(bands[j] >> (8 * e)) & 0xFFu];
          m[e] = (a[j][e] * w.x) >> w.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint2 w = tab.slow[(bands[j] >> (8 * e)) & 0xFFu];
          m[e] = __umulhi(w.x, a[j][e] << 2) >> w.y;
        }
      }
      int bits = 0, last = -1;
      if (!__any_sync(FULL, (m[0] | m[1] | m[2] | m[3]) >= (uint32_t)ENC_LUT_MAG)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { bits += (int)s_bits[m[e]]; last = m[e] ? kc0[j] + e : last; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          bits += m[e] < (uint32_t)ENC_LUT_MAG ? (int)s_bits[m[e]] : 2 * (31 - __clz(min(m[e] + 1u, 65535u))) + 2;
          last = m[e] ? kc0[j] + e : last;
        }
      }
      if (comp[j] < 0) { bits = 0; last = -1; }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (cmask[j][c] == 0u) continue;          // warp uniform
        int t = 0, l = -1;
        if (comp[j] == c) { t = __reduce_add_sync(cmask[j][c], bits); l = __reduce_max_sync(cmask[j][c], last); }
        const int src = __ffs(cmask[j][c]) - 1;
        total[c] += __shfl_sync(FULL, t, src);
        lastpos[c] = max(lastpos[c], __shfl_sync(FULL, l, src));
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int n = g.band_start[c][g.nbands];
      const int count = lastpos[c] < 0 ? 0 : total[c] - (n - 1 - lastpos[c]);
      need += scaled_bytes(count, g.scalar, too_big);
    }
    return need;
  };
  // yss_for_slice (Quantisation.cpp:627-642) at the index in the table: luma only
  auto luma_sse = [&]() -> long long {
    long long acc = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (comp[j] != 0) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned b = (bands[j] >> (8 * e)) & 0xFFu;
        const uint2 ws = tab.slow[b], wb = tab.back[b];
        const uint32_t m = __umulhi(ws.x, a[j][e] << 2) >> ws.y;
        const uint32_t r = (m * wb.x + (m ? wb.y : 0u)) >> 2;
        const unsigned d = a[j][e] - r;
        acc += (long long)(int)(d * d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    return acc;
  };

  // ---- the search, warp uniform: literal replay of EncodeStream.cpp:85-122
  const int avail = p.slice_bytes[s] - 4;
  unsigned flags = 0;
  int trialQ = 63, q = 127, delta = 64;
  bool dead = false;
  while (delta > 0) {
    delta >>= 1;
    bool too_big = false, allfast = false;
    if (!set_index(trialQ, allfast)) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; break; }
    const int need = need_bytes(allfast, too_big);
    if (too_big) { flags |= VC2_FLAG_SCALAR_TOO_SMALL | VC2_FLAG_SEARCH_PHASE; dead = true; break; }
    if (need <= avail) { if (trialQ < q) q = trialQ; trialQ -= delta; }
    else trialQ += delta;
  }
  if (!dead) {
    trialQ = q;
    bool allfast = false;
    bool ok = set_index(trialQ, allfast);
    long long prev = ok ? luma_sse() : 0;
    while (ok) {
      ++trialQ;
      ok = set_index(trialQ, allfast);
      if (!ok) break;
      const long long cur = luma_sse();
      const long long d = cur - prev;
      prev = cur;
      if (!(d < 0)) break;
    }
    if (!ok) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; }
    q = trialQ - 1;
  }
  if (lane == 0) {
    const long long sidx = (long long)pic * nslices + s;
    p.qidx[sidx] = dead ? 0 : q;
    p.err_flags[sidx] = flags;
  }
}

// ------------------------------------------------------------------------------------------
// HQ_CBR rate control, EIGHT LANES PER SLICE (a warp = four slices, a CTA of eight warps = one group of 32 slices), the slice in
// registers; optional, see search_in_subwarps().  The middle way between hq_pack_kernel (a thread per slice: ten passes over memory) and hq_search_warp_kernel (a
// warp per slice: the per-slice work costs a whole warp instruction per slice): lane l of a slice's eight loads pieces l, l + 8,
// ... ONCE, R pieces of four magnitudes; a probe is arithmetic on those registers, per-lane sums per component, and one
// three-step butterfly per component.  Needs every slot (eight consecutive pieces) to lie inside one component - the host checks.
// ------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) hq_search_sub_kernel(const PackParams p) {
  __shared__ uint8_t s_bits[ENC_LUT_MAG];
  __shared__ WarpBand s_band[32];                 // one table per slice of the CTA
  __shared__ uint2 s_qfast[128], s_qslow[128], s_qback[128];
  for (int i = threadIdx.x; i < ENC_LUT_MAG; i += blockDim.x) s_bits[i] = (uint8_t)(d_enc_lut[2 * i] & 31u);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    s_qfast[i] = make_uint2(c_qt.qmul16[i], c_qt.qsh16[i]);
    s_qslow[i] = make_uint2(c_qt.qm31[i], c_qt.ql31[i]);
    s_qback[i] = make_uint2(c_qt.qf[i], c_qt.qo[i] + 2u);
  }
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, l8 = lane & 7;                     // slice inside the warp, lane inside the slice
  const int in_cta = warp * 4 + sub;
  const int s = blockIdx.x * 32 + in_cta;
  const bool live = s < nslices;                                // a dead slice walks along on zeros and writes nothing
  const int nc4 = g.comp_start[3] >> 2;
  const int4* base = reinterpret_cast<const int4*>(p.coef + (long long)pic * g.coef_pic_stride) + (size_t)blockIdx.x * nc4 * 32 + in_cta;
  uint32_t a[R][4];
  uint32_t bands[R];        // the bands of the four coefficients of a piece, 8 bits each
  int kc0[R];               // index of the piece's first coefficient inside its component
  int cslot[R];             // component of slot j (the same for every lane and slice: warp uniform), -1: no such slot
  bool fastmag = true;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int piece = j * 8 + l8, first = 4 * (j * 8);
    cslot[j] = j * 8 >= nc4 ? -1 : (first >= g.comp_start[2] ? 2 : (first >= g.comp_start[1] ? 1 : 0));
    a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0u;
    bands[j] = 0; kc0[j] = 0;
    if (piece < nc4) {
      if (live) {
        const int4 v = __ldg(base + (size_t)piece * 32);
        a[j][0] = (uint32_t)abs(v.x); a[j][1] = (uint32_t)abs(v.y); a[j][2] = (uint32_t)abs(v.z); a[j][3] = (uint32_t)abs(v.w);
      }
      fastmag = fastmag && (a[j][0] | a[j][1] | a[j][2] | a[j][3]) < (uint32_t)VC2_NARROW_FAST_MAX;
      const int c = cslot[j];
      kc0[j] = 4 * piece - g.comp_start[c];
      int b = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        while (kc0[j] + e >= g.band_start[c][b + 1]) ++b;
        bands[j] |= (uint32_t)b << (8 * e);
      }
    } else if (cslot[j] >= 0) {
      kc0[j] = -1 << 20;   // a lane behind the end of the list in the last slot: zeros that are never the last non-zero
    }
  }
  fastmag = __all_sync(FULL, fastmag);
  WarpBand& tab = s_band[in_cta];
  const int ncomp[3] = {g.band_start[0][g.nbands], g.band_start[1][g.nbands], g.band_start[2][g.nbands]};
  // the band parameters of this slice for its index q (the four slices of a warp probe different indices)
  auto set_index = [&](int q, bool& bad) -> bool {   // returns: every band of every slice of the warp has the full-rate form
    bool fast = true;
    bad = false;
    __syncwarp();
    for (int b = l8; b < g.nbands; b += 8) {
      const int aq = max(q - g.qmatrix[b], 0);
      bad = bad || aq > 119;
      const int i = min(aq, 127);
      const uint2 f = s_qfast[i];
      fast = fast && f.x != 0u;
      tab.fast[b] = f; tab.slow[b] = s_qslow[i]; tab.back[b] = s_qback[i];
    }
    __syncwarp();
    // bad is per slice: any of its eight lanes
    bad = (__ballot_sync(FULL, bad) >> (8 * sub) & 0xFFu) != 0u;
    return fastmag && __all_sync(FULL, fast);
  };
  auto need_bytes = [&](bool allfast, bool& too_big) -> int {
    int total[3] = {0, 0, 0}, lastpos[3] = {-1, -1, -1};
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (cslot[j] < 0) continue;   // warp uniform
      uint32_t m[4];
      if (allfast) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint2 w = tab.fast[(bands[j] >> (8 * e)) & 0xFFu];
          m[e] = (a[j][e] * w.x) >> w.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint2 w = tab.slow[(bands[j] >> (8 * e)) & 0xFFu];
          m[e] = __umulhi(w.x, a[j][e] << 2) >> w.y;
        }
      }
      int bits = 0, last = -1;
      if (!__any_sync(FULL, (m[0] | m[1] | m[2] | m[3]) >= (uint32_t)ENC_LUT_MAG)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { bits += (int)s_bits[m[e]]; last = m[e] ? kc0[j] + e : last; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          bits += m[e] < (uint32_t)ENC_LUT_MAG ? (int)s_bits[m[e]] : 2 * (31 - __clz(min(m[e] + 1u, 65535u))) + 2;
          last = m[e] ? kc0[j] + e : last;
        }
      }
      if (kc0[j] < 0) bits = 0;     // no piece here (the tail of the last slot)
      // the slot's component is warp uniform: these are plain branches
      if (cslot[j] == 0) { total[0] += bits; lastpos[0] = max(lastpos[0], last); }
      else if (cslot[j] == 1) { total[1] += bits; lastpos[1] = max(lastpos[1], last); }
      else { total[2] += bits; lastpos[2] = max(lastpos[2], last); }
    }
    int need = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {   // butterfly over the eight lanes of the slice
        total[c] += __shfl_xor_sync(FULL, total[c], o);
        lastpos[c] = max(lastpos[c], __shfl_xor_sync(FULL, lastpos[c], o));
      }
      const int count = lastpos[c] < 0 ? 0 : total[c] - (ncomp[c] - 1 - lastpos[c]);
      if (g.scalar == 1) {
        const int units = (count + 7) >> 3;
        if (units > 0xFF) too_big = true;
        need += units;
      } else {
        need += scaled_bytes(count, g.scalar, too_big);
      }
    }
    return need;
  };
  auto luma_sse = [&]() -> long long {
    long long acc = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (cslot[j] != 0) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned b = (bands[j] >> (8 * e)) & 0xFFu;
        const uint2 ws = tab.slow[b], wb = tab.back[b];
        const uint32_t m = __umulhi(ws.x, a[j][e] << 2) >> ws.y;
        const uint32_t r = (m * wb.x + (m ? wb.y : 0u)) >> 2;
        const unsigned d = a[j][e] - r;
        acc += (long long)(int)(d * d);
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    return acc;
  };

  // ---- the search of EncodeStream.cpp:85-122, four slices in lockstep: the seven probes are seven iterations for every slice,
  // a slice that has died (or finished its squared-error walk) idles through the rest
  const int avail = live ? p.slice_bytes[s] - 4 : 0;
  unsigned flags = 0;
  int trialQ = 63, q = 127, delta = 64;
  bool dead = false;
  while (delta > 0) {
    delta >>= 1;
    bool too_big = false, bad = false;
    const bool allfast = set_index(trialQ, bad);
    const int need = need_bytes(allfast, too_big);
    if (!dead) {
      if (bad) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; }
      else if (too_big) { flags |= VC2_FLAG_SCALAR_TOO_SMALL | VC2_FLAG_SEARCH_PHASE; dead = true; }
      else if (need <= avail) { if (trialQ < q) q = trialQ; trialQ -= delta; }
      else trialQ += delta;
    }
  }
  // "try a few higher quantisers": every slice walks until its squared error stops dropping; the warp until all four have stopped
  bool walking = !dead;
  trialQ = q;
  long long prev = 0;
  {
    bool bad = false;
    set_index(trialQ, bad);
    const long long first = luma_sse();
    if (walking && bad) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; walking = false; }
    prev = first;
  }
  while (__any_sync(FULL, walking)) {
    const int tq = walking ? trialQ + 1 : trialQ;
    bool bad = false;
    set_index(tq, bad);
    const long long cur = luma_sse();
    if (walking) {
      trialQ = tq;
      if (bad) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; walking = false; }
      else {
        const long long d = cur - prev;
        prev = cur;
        if (!(d < 0)) walking = false;
      }
    }
  }
  if (live && l8 == 0) {
    const long long sidx = (long long)pic * nslices + s;
    p.qidx[sidx] = dead ? 0 : trialQ - 1;
    p.err_flags[sidx] = flags;
  }
}

// ------------------------------------------------------------------------------------------
// HQ slice encoder: ONE THREAD PER SLICE, a warp = one group of 32 consecutive slices.
//   stream the slice's coefficient list (coalesced through the group-interleaved layout) ->
//   [quantIndicesCBR, literal replay] -> quantise -> SignedVLC -> MSB-first words of the slice image
//   (prefix | qindex | len Y | Y | len U | U | len V | V, Slices.cpp:478-530) in the staging buffer.
// The slice sizes are scanned and the images gathered into the payload by assemble_kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) hq_pack_kernel(const PackParams p) {
  __shared__ uint32_t s_enc[2 * ENC_LUT_MAG];
  __shared__ uint8_t s_bits[ENC_LUT_MAG];   // code lengths alone, for the rate-control probes
  stage_table(s_enc, d_enc_lut, 2 * ENC_LUT_MAG);
  for (int i = threadIdx.x; i < ENC_LUT_MAG; i += blockDim.x) s_bits[i] = (uint8_t)(d_enc_lut[2 * i] & 31u);
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  const int nc4 = g.comp_start[3] >> 2;
  // this picture's block and the lane's first piece in it
  const int4* base = reinterpret_cast<const int4*>(p.coef + (long long)pic * g.coef_pic_stride);
  const size_t src = (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  const long long sidx = (long long)pic * nslices + s;
  unsigned flags = 0;
  int qi = 0;

  if (p.search) {
    auto walk = [&](int c, int q, bool& badq, auto& op) { walk_component(base, src + (size_t)(g.comp_start[c] >> 2) * 32, g, c, q, badq, op); };
    qi = search_slice(p, g, s, s_enc, s_bits, flags, walk);
    p.qidx[sidx] = qi;
  } else if (p.const_q >= 0) {
    qi = p.const_q;
    p.qidx[sidx] = qi;
  } else {
    qi = p.qidx[sidx];
    if (p.after_search) flags = p.err_flags[sidx];
  }

  if (p.emit) {
    WideBitWriter W;
    W.init(p.staging + sidx * p.wcap);
    int total = 0;
    if (!(flags & VC2_FLAG_SEARCH_PHASE)) {
      for (int i = 0; i < g.prefix; ++i) W.put(0u, 8);
      W.put((uint32_t)qi & 0xFFu, 8);
      bool too_big = false, badq = false;
      unsigned bigor = 0;
      int lensum = 0;
      for (int c = 0; c < 3; ++c) {
        const int len_pos = W.pos();
        W.put(0u, 8);
        const int data_start = W.pos();
        const size_t csrc = src + (size_t)(g.comp_start[c] >> 2) * 32;
        int last;
        if (p.quantise) {
          EmitOp<true, WideBitWriter> op = {s_enc, &W, W.mark(), 0u};
          walk_component(base, csrc, g, c, qi, badq, op);
          last = W.unmark(op.last); bigor |= op.bigor;
        } else {
          EmitOp<false, WideBitWriter> op = {s_enc, &W, W.mark(), 0u};
          walk_component(base, csrc, g, c, qi, badq, op);
          last = W.unmark(op.last); bigor |= op.bigor;
        }
        int L = scaled_bytes(last - data_start, g.scalar, too_big);
        if (c == 2 && p.mode == VC2_HQ_CBR) {
          // V takes all remaining bytes (Slices.cpp:355-366)
          const int vBytes = p.slice_bytes[s] - 4 - lensum;
          if (vBytes < L) flags |= VC2_FLAG_CBR_TOO_MANY_BYTES;
          else if (vBytes / g.scalar > 255) flags |= VC2_FLAG_CBR_COMP_LENGTH;
          else L = vBytes;
        }
        if (too_big) { flags |= VC2_FLAG_SCALAR_TOO_SMALL; break; }
        lensum += L;
        W.seek(data_start + 8 * L);
        W.patch_byte(len_pos, (uint32_t)(L / g.scalar) & 0xFFu);
      }
      if (p.quantise && badq) flags |= VC2_FLAG_QUANT_INDEX;
      if (bigor >> 16) flags |= VC2_FLAG_VLC_RANGE;
      W.finish();
      total = g.prefix + 4 + lensum;
    }
    if (flags & ~VC2_FLAG_VLC_RANGE) total = 0;
    p.sizes[sidx] = (uint32_t)total;
  }
  p.err_flags[sidx] = flags;
}

// ------------------------------------------------------------------------------------------
// HQ slice encoder over the NARROW coefficient block (HQ_ConstQ): the lifting kernels have already quantised, a
// coefficient arrives as the 16-bit word t = 2 * |q| + (q < 0), which IS the index of the code table.  No band
// bookkeeping is left: a component is one flat run of 8-byte pieces (four coefficients), 256 bytes apart per lane.
// Same slice syntax, writer and staging images as hq_pack_kernel.
// ------------------------------------------------------------------------------------------
template <bool RO>
__device__ __forceinline__ void copy_slice_image(const uint32_t* img, uint8_t* dst, int total, int lane, uint32_t first0, uint32_t first1);

struct NarrowEmit {
  unsigned lut;    // shared-memory address of the code table (kept in a register: the look-up is one LDS behind a shift-add)
  WideBitWriter* W;
  unsigned last;   // cursor behind the last non-zero coefficient
  __device__ __forceinline__ uint32_t entry(uint32_t t) const {
    uint32_t e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(lut + 4u * t));
    return e;
  }
  __device__ __forceinline__ void one(uint32_t t) {
    uint32_t code;
    int nb;
    if (t < 2u * (uint32_t)ENC_LUT_MAG) {
      const uint32_t e = entry(t);
      nb = (int)(e & 31u);
      code = e >> 5;
    } else {   // magnitudes up to 32767: m = |q| + 1 <= 2^15, 2 * 15 + 2 = 32 bits at most
      const uint32_t m = (t >> 1) + 1u;
      const int k = 31 - __clz(m);
      nb = 2 * k + 2;
      code = (spread16(m ^ (1u << k)) << 2) | 2u | (t & 1u);
    }
    W->put(code, nb);
    last = t >= 2u ? W->mark() : last;
  }
  // two coefficients (the two halves of w) whose codes come from the table and are at most 16 bits long: ONE append
  __device__ __forceinline__ void pair(uint32_t w) {
    const uint32_t t0 = w & 0xFFFFu, t1 = w >> 16;
    if ((w & 0xFF00FF00u) == 0u) {
      const uint32_t e0 = entry(t0), e1 = entry(t1);
      const int nb0 = (int)(e0 & 31u), nb1 = (int)(e1 & 31u);
      const unsigned mid = W->mark() + (unsigned)nb0;
      W->put(((e0 >> 5) << nb1) | (e1 >> 5), nb0 + nb1);
      last = t1 >= 2u ? W->mark() : (t0 >= 2u ? mid : last);
    } else {
      one(t0);
      one(t1);
    }
  }
};

__global__ void __launch_bounds__(128) hq_pack_narrow_kernel(const PackParams p) {
  __shared__ uint32_t s_enc[2 * ENC_LUT_MAG];
  __shared__ unsigned s_tile, s_warp_sum[4], s_before;
  const int pic = blockIdx.y;
  if (p.fuse && threadIdx.x == 0) s_tile = atomicAdd(p.tile_ticket + pic, 1u);
  stage_table(s_enc, d_enc_lut, 2 * ENC_LUT_MAG);
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int tile = p.fuse ? (int)s_tile : (int)blockIdx.x;
  const int s = tile * blockDim.x + threadIdx.x;
  if (!p.fuse && s >= nslices) return;
  unsigned size = 0;
  if (s < nslices) {
  const int nc4 = g.comp_start[3] >> 2;
  const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.coef) + (long long)pic * g.coef_pic_stride) +
                     (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  const long long sidx = (long long)pic * nslices + s;
  unsigned flags = 0;
  const int qi = p.const_q;
  p.qidx[sidx] = qi;
  // the reference rejects an index beyond its quantiser table when it quantises (Quantisation.cpp:60-63)
  for (int b = 0; b < g.nbands; ++b) if (max(qi - g.qmatrix[b], 0) > 119) flags |= VC2_FLAG_QUANT_INDEX;
  WideBitWriter W;
  W.init(p.staging + sidx * p.wcap);
  for (int i = 0; i < g.prefix; ++i) W.put(0u, 8);
  W.put((uint32_t)qi & 0xFFu, 8);
  bool too_big = false;
  int lensum = 0;
  for (int c = 0; c < 3; ++c) {
    const int len_pos = W.pos();
    W.put(0u, 8);
    const int data_start = W.pos();
    const uint2* csrc = src + (size_t)(g.comp_start[c] >> 2) * 32;
    const int np = g.band_start[c][g.nbands] >> 2;
    NarrowEmit op = {(unsigned)__cvta_generic_to_shared(s_enc), &W, W.mark()};
    // three pieces in registers (every one a fresh line, see walk_component), the lines of the pieces PF ahead on their
    // way into L1: a piece is 8 bytes per lane, the warp walks its group 256 bytes at a time
    constexpr int PF = 8;
    uint2 n0 = __ldg(csrc), n1 = n0, n2 = n0;
    if (np > 1) n1 = __ldg(csrc + 32);
    if (np > 2) n2 = __ldg(csrc + 64);
    const uint2* pf = csrc + 96;
    for (int i = 0; i < np; ++i) {
      const uint2 w = n0;
      n0 = n1; n1 = n2;
      if (i + 3 < np) n2 = __ldg(pf);
      if (i + 3 + PF < np) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + 32 * PF));
      pf += 32;
      op.pair(w.x);
      op.pair(w.y);
    }
    const int L = scaled_bytes(W.unmark(op.last) - data_start, g.scalar, too_big);
    if (too_big) { flags |= VC2_FLAG_SCALAR_TOO_SMALL; break; }
    lensum += L;
    W.seek(data_start + 8 * L);
    W.patch_byte(len_pos, (uint32_t)(L / g.scalar) & 0xFFu);
  }
  W.finish();
  size = flags ? 0u : (uint32_t)(g.prefix + 4 + lensum);
  p.sizes[sidx] = size;
  p.err_flags[sidx] = flags;
  }
  if (!p.fuse) return;

  // ---- where do this CTA's slices go?  scan inside the CTA, then the bytes of all the tiles in front of it
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = size;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned v = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  __syncthreads();
  unsigned before_warp = 0, cta_total = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    if (w < warp) before_warp += s_warp_sum[w];
    cta_total += s_warp_sum[w];
  }
  if (threadIdx.x == 0) {
    volatile unsigned long long* st = p.tile_state + (long long)pic * p.tiles;
    unsigned before = 0;
    if (tile == 0) st[0] = (2ull << 32) | cta_total;
    else {
      st[tile] = (1ull << 32) | cta_total;
      for (int j = tile - 1;; --j) {
        unsigned long long v;
        do { v = st[j]; } while ((v >> 32) == 0);   // tile j started before this one (tickets): it will publish
        before += (unsigned)v;
        if ((v >> 32) == 2) break;
      }
      st[tile] = (2ull << 32) | (before + cta_total);
    }
    s_before = before;
  }
  __syncthreads();
  const unsigned offset = s_before + before_warp + incl - size;
  uint32_t* so = p.slice_off + (long long)pic * (nslices + 1);
  if (s < nslices) so[s] = offset;
  if (s == nslices - 1) {
    so[nslices] = offset + size;
    if (p.total_len) p.total_len[pic] = offset + size;
  }
  // ---- gather: the warp copies its 32 slice images (written a moment ago: L2 hits) to their place in the payload
  __threadfence_block();
  __syncwarp();
  const long long sidx0 = (long long)pic * nslices + (s - lane);
  uint8_t* out = p.out + (long long)pic * p.out_pic_stride;
  for (int j = 0; j < 32; ++j) {
    const unsigned off_j = __shfl_sync(FULL, offset, j);
    const int total = (int)__shfl_sync(FULL, size, j);
    if (total <= 0) continue;
    if ((long long)off_j + total > p.out_capacity) {
      if (lane == 0) atomicOr(&p.err_flags[sidx0 + j], VC2_FLAG_STREAM);
      continue;
    }
    const uint32_t* img = p.staging + (sidx0 + j) * p.wcap;
    copy_slice_image<false>(img, out + off_j, total, lane, img[lane], img[lane + 1]);
  }
}

// ------------------------------------------------------------------------------------------
// LD encoder.  The reference picks the slice quantisers in raster order (quantIndicesLD,
// EncodeStream.cpp:193-245): seven probes of a binary search per slice, each probe quantising the whole
// slice - with the LL band predicted from its already decoded neighbours, which is what chains the slices
// together.  Only the few LL samples of a slice take part in that chain, so the work is split in three:
//   1. ld_ac_bits_kernel   every slice x every index 0..127, in parallel: bits up to the last non-zero
//                          coefficient of the bands other than LL (luma list; U/V interleaved list)
//   2. ld_rate_kernel      one CTA per picture walks the slice anti-diagonals (a slice needs its W, N and NW
//                          neighbours): the reference's search with the LL samples quantised for real and the
//                          rest looked up; leaves the chosen index, the locally decoded LL band and the
//                          quantised LL residuals
//   3. ld_pack_kernel      one thread per slice: quantise + the LD slice syntax (Slices.cpp:195-244)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_intlog2(int v) {   // utils::intlog2 (Utils.cpp:40-48)
  int lg = 0;
  --v;
  while (v > 0) { v >>= 1; ++lg; }
  return lg;
}
__device__ __forceinline__ int lut_bits(const uint32_t* lut, uint32_t mag) {   // SignedVLC(v).numOfBits(), v of magnitude mag
  if (mag < (uint32_t)ENC_LUT_MAG) return (int)(lut[2u * mag] & 31u);
  return 2 * (31 - __clz(min(mag + 1u, 65535u))) + 2;
}

__global__ void __launch_bounds__(128) ld_ac_bits_kernel(const LdEncParams p) {
  __shared__ uint32_t s_enc[2 * ENC_LUT_MAG];
  stage_table(s_enc, d_enc_lut, 2 * ENC_LUT_MAG);
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * 32 + lane;        // the four warps of a CTA share one group of 32 slices
  if (s >= nslices) return;
  const int nc4 = g.comp_start[3] >> 2;
  const int4* src = reinterpret_cast<const int4*>(p.coef + (long long)pic * g.coef_pic_stride) + (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  uint32_t* tab = p.acbits + ((long long)pic * nslices + s) * 256;
  // q is warp uniform: blockIdx.y * 32 .. + 31, eight per warp
  for (int qi = 0; qi < 8; ++qi) {
    const int q = blockIdx.y * 32 + warp * 8 + qi;
    bool badq = false;
    unsigned res[2];
    for (int cls = 0; cls < 2; ++cls) {          // luma list, then the U/V interleaved list
      const int c = cls;                           // band geometry of Y, or of the chroma planes
      const int n = g.band_start[c][g.nbands], nll = g.band_start[c][1];
      const int4* a = src + (size_t)(g.comp_start[cls == 0 ? 0 : 1] >> 2) * 32;
      const int4* bsrc = src + (size_t)(g.comp_start[2] >> 2) * 32;
      int gross = 0, last = 0, k = 0, b = 0, bend = g.band_start[c][1];
      BandP bp = band_params(q, g.qmatrix[0], badq);
      for (int piece = 0; piece < (n >> 2); ++piece) {
        const int4 u4 = __ldg(a + (size_t)piece * 32);
        int4 v4 = u4;
        if (cls) v4 = __ldg(bsrc + (size_t)piece * 32);
        const int u[4] = {u4.x, u4.y, u4.z, u4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e, ++k) {
          while (k == bend) { ++b; bend = g.band_start[c][b + 1]; bp = band_params(q, g.qmatrix[b], badq); }
          if (k < nll) continue;                   // the LL band belongs to the rate kernel
          uint32_t mag = quant_mag(u[e], bp);
          gross += lut_bits(s_enc, mag);
          if (mag) last = gross;
          if (cls) {
            mag = quant_mag(v[e], bp);
            gross += lut_bits(s_enc, mag);
            if (mag) last = gross;
          }
        }
      }
      res[cls] = (unsigned)last;
    }
    tab[2 * q] = badq ? 0xFFFFFFFFu : res[0];
    tab[2 * q + 1] = badq ? 0xFFFFFFFFu : res[1];
  }
}

// the LL samples of one slice for trial index q: quantise against the prediction, restore, count bits.
// Returns false when the index is outside the quantiser table (Quantisation.cpp:60-63).
struct LdLL {
  const LdEncParams* p;
  int32_t* rest;               // this picture's restored LL planes
  const int32_t* coefpic;      // this picture's coefficient block
  int32_t* qpic;               // this picture's quantised block (LL residuals are stored here)
  int nc4;
  __device__ __forceinline__ int predict(const int32_t* r, int w, int y, int x) const {   // predictDC (Quantisation.cpp:191-208)
    if (y > 0 && x > 0) {
      const int sum = r[(y - 1) * w + x - 1] + r[(y - 1) * w + x] + r[y * w + x - 1];
      return sum >= 0 ? (sum + 1) / 3 : (sum - 1) / 3;
    }
    if (y > 0) return r[(y - 1) * w + x];
    if (x > 0) return r[y * w + x - 1];
    return 0;
  }
  // class 0: luma LL samples; class 1: U and V LL samples interleaved.  gross / last as in luma_slice_bits / chroma_slice_bits
  __device__ __forceinline__ void run(const uint32_t* lut, int s, int sy, int sx, int q, int cls, bool store, int& gross, int& last) const {
    const SliceGeom& g = p->g;
    const int aq = max(q - g.qmatrix[0], 0);
    const QParam qp = qparam(aq);
    const int c0 = cls == 0 ? 0 : 1, ncomp = cls == 0 ? 1 : 2;
    const int bh = g.part_h[c0][0], bw = g.part_w[c0][0];
    gross = 0; last = 0;
    for (int ly = 0; ly < bh; ++ly)
      for (int lx = 0; lx < bw; ++lx)
        for (int j = 0; j < ncomp; ++j) {
          const int c = c0 + j;
          int32_t* r = rest + p->ll_off[c];
          const int w = p->ll_w[c], y = sy * bh + ly, x = sx * bw + lx;
          const long long idx = coef_index(s, g.comp_start[c] + ly * bw + lx, nc4);
          const int pred = predict(r, w, y, x);
          const int qv = quant_one(coefpic[idx] - pred, qp.qm, qp.ql);
          r[y * w + x] = scale_one(qv, qp.qf, qp.qo) + pred;
          if (store) qpic[idx] = qv;
          gross += lut_bits(lut, (uint32_t)abs(qv));
          if (qv) last = gross;
        }
  }
};

__global__ void __launch_bounds__(256) ld_rate_kernel(const LdEncParams p) {
  __shared__ uint32_t s_enc[2 * ENC_LUT_MAG];
  stage_table(s_enc, d_enc_lut, 2 * ENC_LUT_MAG);
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y, pic = blockIdx.x;
  LdLL ll;
  ll.p = &p;
  ll.rest = p.restored + (long long)pic * p.ll_stride;
  ll.coefpic = p.coef + (long long)pic * g.coef_pic_stride;
  ll.qpic = p.qcoef + (long long)pic * g.coef_pic_stride;
  ll.nc4 = g.comp_start[3] >> 2;
  for (int diag = 0; diag < g.slices_y + g.slices_x - 1; ++diag) {
    const int vlo = max(0, diag - (g.slices_x - 1)), vhi = min(g.slices_y - 1, diag);
    for (int v = vlo + (int)threadIdx.x; v <= vhi; v += blockDim.x) {
      const int h = diag - v, s = v * g.slices_x + h;
      const long long sidx = (long long)pic * nslices + s;
      const uint32_t* tab = p.acbits + sidx * 256;
      const int bytes = p.slice_bytes[s];
      const int avail = 8 * bytes - 7 - ld_intlog2(8 * bytes - 7);
      unsigned flags = 0;
      auto probe = [&](int q, bool store, int& bits) -> bool {   // bits the slice needs at index q
        const unsigned acy = tab[2 * q], acuv = tab[2 * q + 1];
        if (acy == 0xFFFFFFFFu || max(q - g.qmatrix[0], 0) > 119) return false;
        int gy, ly, guv, luv;
        ll.run(s_enc, s, v, h, q, 0, store, gy, ly);
        ll.run(s_enc, s, v, h, q, 1, store, guv, luv);
        bits = (acy ? gy + (int)acy : ly) + (acuv ? guv + (int)acuv : luv);
        return true;
      };
      int trialQ = 63, q = 127, delta = 64, bits = 0;
      bool dead = false;
      while (delta > 0) {
        delta >>= 1;
        if (!probe(trialQ, false, bits)) { dead = true; break; }
        if (bits <= avail) { if (trialQ < q) q = trialQ; trialQ -= delta; }
        else trialQ += delta;
      }
      // the slice is quantised again with the chosen index: the neighbours predict from that state (:232-236)
      if (!dead && !probe(q, true, bits)) dead = true;
      if (dead) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; q = 0; }
      p.qidx[sidx] = q;
      p.err_flags[sidx] = flags;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) ld_pack_kernel(const LdEncParams p) {
  __shared__ uint32_t s_enc[2 * ENC_LUT_MAG];
  stage_table(s_enc, d_enc_lut, 2 * ENC_LUT_MAG);
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  const int nc4 = g.comp_start[3] >> 2;
  const size_t lane_off = (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  const int4* src = reinterpret_cast<const int4*>(p.coef + (long long)pic * g.coef_pic_stride) + lane_off;
  int4* qdst = reinterpret_cast<int4*>(p.qcoef + (long long)pic * g.coef_pic_stride) + lane_off;
  const long long sidx = (long long)pic * nslices + s;
  unsigned flags = p.err_flags[sidx];
  const int qi = p.qidx[sidx];
  const int size = p.slice_bytes[s];
  if (flags) { p.sizes[sidx] = 0; return; }
  bool badq = false;
  // pass 1: quantise (the LL residuals are already in qcoef), keep the quantised values, count the bits
  int ybits = 0, uvbits_needed = 0;
  unsigned bigor = 0;
  for (int cls = 0; cls < 2; ++cls) {
    const int c = cls;
    const int n = g.band_start[c][g.nbands], nll = g.band_start[c][1];
    const size_t o0 = (size_t)(g.comp_start[cls == 0 ? 0 : 1] >> 2) * 32, o1 = (size_t)(g.comp_start[2] >> 2) * 32;
    int gross = 0, last = 0, k = 0, b = 0, bend = g.band_start[c][1];
    BandP bp = band_params(qi, g.qmatrix[0], badq);
    for (int piece = 0; piece < (n >> 2); ++piece) {
      const int4 u4 = __ldg(src + o0 + (size_t)piece * 32);
      const int4 uq4 = qdst[o0 + (size_t)piece * 32];
      int4 v4 = u4, vq4 = uq4;
      if (cls) { v4 = __ldg(src + o1 + (size_t)piece * 32); vq4 = qdst[o1 + (size_t)piece * 32]; }
      const int u[4] = {u4.x, u4.y, u4.z, u4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
      int uq[4] = {uq4.x, uq4.y, uq4.z, uq4.w}, vq[4] = {vq4.x, vq4.y, vq4.z, vq4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e, ++k) {
        while (k == bend) { ++b; bend = g.band_start[c][b + 1]; bp = band_params(qi, g.qmatrix[b], badq); }
        if (k >= nll && !p.prequantised) uq[e] = quant_band(u[e], bp);
        uint32_t mag = (uint32_t)abs(uq[e]);
        if (mag >= (uint32_t)ENC_LUT_MAG) bigor |= mag + 1u;
        gross += lut_bits(s_enc, mag);
        if (mag) last = gross;
        if (cls) {
          if (k >= nll && !p.prequantised) vq[e] = quant_band(v[e], bp);
          mag = (uint32_t)abs(vq[e]);
          if (mag >= (uint32_t)ENC_LUT_MAG) bigor |= mag + 1u;
          gross += lut_bits(s_enc, mag);
          if (mag) last = gross;
        }
      }
      qdst[o0 + (size_t)piece * 32] = make_int4(uq[0], uq[1], uq[2], uq[3]);
      if (cls) qdst[o1 + (size_t)piece * 32] = make_int4(vq[0], vq[1], vq[2], vq[3]);
    }
    if (cls == 0) ybits = last; else uvbits_needed = last;
  }
  // Slices.cpp:203-213
  const int split = ld_intlog2(8 * size - 7);
  const int uvbits = 8 * size - 7 - split - ybits;
  if (uvbits < uvbits_needed) flags |= VC2_FLAG_LD_TOO_MANY_BYTES;
  if (badq && !p.prequantised) flags |= VC2_FLAG_QUANT_INDEX;   // the writer alone never evaluates quant_factor
  if (bigor >> 16) flags |= VC2_FLAG_VLC_RANGE;
  if (flags & ~VC2_FLAG_VLC_RANGE) { p.err_flags[sidx] = flags; p.sizes[sidx] = 0; return; }
  // pass 2: qindex (7 bits) | luma length | luma, bounded to its exact length | U/V interleaved, bounded to the rest
  BitWriter W;
  W.init(p.staging + sidx * p.wcap);
  W.put((uint32_t)qi & 0x7Fu, 7);
  W.put((uint32_t)ybits, split);
  for (int cls = 0; cls < 2; ++cls) {
    const int c = cls;
    const int n = g.band_start[c][g.nbands];
    const size_t o0 = (size_t)(g.comp_start[cls == 0 ? 0 : 1] >> 2) * 32, o1 = (size_t)(g.comp_start[2] >> 2) * 32;
    const int start = W.pos();
    for (int piece = 0; piece < (n >> 2); ++piece) {
      const int4 uq4 = qdst[o0 + (size_t)piece * 32];
      int4 vq4 = uq4;
      if (cls) vq4 = qdst[o1 + (size_t)piece * 32];
      const int uq[4] = {uq4.x, uq4.y, uq4.z, uq4.w}, vq[4] = {vq4.x, vq4.y, vq4.z, vq4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t code;
        int nb;
        unsigned dummy = 0;
        vlc_of(s_enc, (uint32_t)abs(uq[e]), uq[e] < 0, dummy, code, nb);
        W.put(code, nb);
        if (cls) {
          vlc_of(s_enc, (uint32_t)abs(vq[e]), vq[e] < 0, dummy, code, nb);
          W.put(code, nb);
        }
      }
    }
    // vlc::bounded + flush (VLC.cpp:151-185): ones beyond the bound are dropped, zeros fill up to it
    W.seek(start + (cls == 0 ? ybits : uvbits));
  }
  W.finish();
  p.err_flags[sidx] = flags;
  p.sizes[sidx] = (uint32_t)size;
}

// ------------------------------------------------------------------------------------------
// Slice offsets = exclusive scan of the slice sizes in raster order (one CTA per picture), or the
// a-priori table in CBR mode.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) slice_scan_kernel(const AssembleParams p) {
  __shared__ unsigned s_part[1024];
  const int pic = blockIdx.x, n = p.nslices, t = threadIdx.x;
  uint32_t* so = p.slice_off + (long long)pic * (n + 1);
  if (p.fixed_off) {
    for (int i = t; i <= n; i += blockDim.x) so[i] = p.fixed_off[i];
    if (t == 0 && p.total_len) p.total_len[pic] = p.fixed_off[n];
    return;
  }
  const uint32_t* sz = p.sizes + (long long)pic * n;
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int i0 = min(t * per, n), i1 = min(i0 + per, n);
  unsigned sum = 0;
  for (int i = i0; i < i1; ++i) sum += sz[i];
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {   // Hillis-Steele inclusive scan of the per-thread sums
    const unsigned v = t >= o ? s_part[t - o] : 0u;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  unsigned run = s_part[t] - sum;
  for (int i = i0; i < i1; ++i) { so[i] = run; run += sz[i]; }
  if (t == (int)blockDim.x - 1) {
    so[n] = s_part[t];
    if (p.total_len) p.total_len[pic] = s_part[t];
  }
}

// ------------------------------------------------------------------------------------------
// Gather: one warp copies one slice image from the staging words (MSB-first) to its byte offset in
// the payload.  Head bytes up to 4-byte alignment, aligned 32-bit words, tail bytes.
// ------------------------------------------------------------------------------------------
// one warp copies `total` bytes of a slice image (MSB-first words) to dst: head bytes up to 4-byte alignment, aligned
// 32-bit words, tail bytes.  first0 / first1 = image words lane, lane + 1 (fetched by the caller, early).  RO: the images were
// written by an earlier kernel (read-only path); else by this kernel (plain loads)
template <bool RO>
__device__ __forceinline__ void copy_slice_image(const uint32_t* img, uint8_t* dst, int total, int lane, uint32_t first0, uint32_t first1) {
  const int head = min((int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3), total);
  const uint32_t img0 = __shfl_sync(FULL, first0, 0);
  if (lane < head) dst[lane] = (uint8_t)(img0 >> (24 - 8 * lane));
  const int nwords = (total - head) >> 2;
  uint32_t* __restrict__ dw = reinterpret_cast<uint32_t*>(dst + head);
  // image bytes head + 4m .. head + 4m + 3 = the funnel of words m, m + 1 (shift 0 when the slice starts aligned)
  if (lane < nwords) dw[lane] = __byte_perm(__funnelshift_l(first1, first0, 8 * head), 0, 0x0123);
#pragma unroll 2
  for (int m = lane + 32; m < nwords; m += 32) {
    const uint32_t be = RO ? __funnelshift_l(__ldg(img + m + 1), __ldg(img + m), 8 * head) : __funnelshift_l(img[m + 1], img[m], 8 * head);
    dw[m] = __byte_perm(be, 0, 0x0123);
  }
  const int done = head + 4 * nwords, tail = total - done;
  if (lane < tail) {
    const int i = done + lane;
    dst[i] = (uint8_t)(img[i >> 2] >> (24 - 8 * (i & 3)));
  }
}

__global__ void __launch_bounds__(256) assemble_kernel(const AssembleParams p) {
  const int pic = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= p.nslices) return;
  const long long sidx = (long long)pic * p.nslices + s;
  const uint32_t* __restrict__ img = p.staging + sidx * p.wcap;
  // the first 33 words of the image do not depend on where the slice goes: fetch them together with its offset and
  // size instead of one DRAM round trip later (the kernel is latency bound: a slice is six 128-byte rows)
  const uint32_t first0 = __ldg(img + lane), first1 = __ldg(img + lane + 1);
  const uint32_t* so = p.slice_off + (long long)pic * (p.nslices + 1);
  const unsigned offset = so[s];
  int total = (int)p.sizes[sidx];
  if (p.fixed_off && total > 0) total = (int)(so[s + 1] - offset);
  if (total <= 0) return;
  if ((long long)offset + total > p.out_capacity) {
    if (lane == 0) atomicOr(&p.err_flags[sidx], VC2_FLAG_STREAM);
    return;
  }
  copy_slice_image<true>(img, p.out + (long long)pic * p.out_pic_stride + offset, total, lane, first0, first1);
}

// ------------------------------------------------------------------------------------------
// MSB-first bit reader over global memory, one thread per slice: 64-bit window (hi:lo), refilled
// 32 bits at a time from aligned words with one word of prefetch.  The reader is BOUNDED
// (vlc::bounded, VLC.cpp:182-185): bits at and beyond `bound` read as ones; they are forced to one
// when a word enters the window, so the per-code path carries no bound arithmetic.
// ------------------------------------------------------------------------------------------
struct BitReader {
  const uint32_t* p;   // next aligned word to fetch
  uint32_t hi, lo;     // the window, top aligned; bits below the nb valid ones are zero
  uint32_t nxt;        // the word that follows the window, AS LOADED: it is byte swapped and bounded only when it enters the
                       // window, so that nothing touches the register between the load and the next refill (the swap
                       // right behind the load was half of all stall samples of the parser, ncu source view)
  int nb;              // valid bits in hi:lo
  int left;            // real stream bits behind nxt (<= 0: only ones follow)
  __device__ __forceinline__ static uint32_t be(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
  // Past the bound nothing is loaded: a component whose length byte is too small for its coefficients would otherwise walk
  // one word per 32 coefficients beyond its slice (and, for the last slice, beyond the payload buffer).
  __device__ __forceinline__ void fetch() {
    nxt = left > 0 ? __ldg(p) : 0xFFFFFFFFu;   // left <= 0: the word holds no real bit, next_word() makes it all ones anyway
    ++p;
    left -= 32;
  }
  // the prefetched word as window bits: ones from the bound on (left + 32 real bits were behind the window when it was fetched)
  __device__ __forceinline__ uint32_t next_word() const { return be(nxt) | __funnelshift_rc(0xFFFFFFFFu, 0u, max(left + 32, 0)); }
  // start reading at byte address a (any alignment); the first `bound` bits are real.  The buffers have slack
  // behind the data.  Leaves nb >= 33.
  __device__ __forceinline__ void init(const uint8_t* a, int bound) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(a);
    p = reinterpret_cast<const uint32_t*>(u & ~(uintptr_t)3);
    const int lead = 8 * (int)(u & 3);
    hi = be(__ldg(p)); lo = be(__ldg(p + 1));
    p += 2;
    const int real = lead + bound;   // real bits among the 64 just loaded
    if (real < 64) {
      const unsigned long long ones = real <= 0 ? ~0ull : (~0ull >> real);
      hi |= (uint32_t)(ones >> 32);
      lo |= (uint32_t)ones;
    }
    left = real - 64;
    fetch();
    hi = __funnelshift_l(lo, hi, lead);
    lo <<= lead;
    nb = 64 - lead;
  }
  // nb >= 33 afterwards.  Straight-line code: the lanes of a warp (one slice each) run dry at different codes.
  __device__ __forceinline__ void ensure() {
    if (nb <= 32) {   // lo is empty: append the prefetched word behind the nb valid bits of hi
      const uint32_t w = next_word();
      hi |= __funnelshift_rc(w, 0u, nb);
      lo = __funnelshift_lc(0u, w, 32 - nb);
      nb += 32;
      fetch();
    }
  }
  __device__ __forceinline__ void consume(int n) {   // n in 0..31, n <= nb
    hi = __funnelshift_l(lo, hi, n);
    lo <<= n;
    nb -= n;
  }
  __device__ __forceinline__ void consume32(int n) {   // n in 0..32, n <= nb
    hi = __funnelshift_lc(lo, hi, n);
    lo = n >= 32 ? 0u : lo << n;
    nb -= n;
  }
  __device__ __forceinline__ void skip(int n) { consume32(n); ensure(); }   // needs nb >= 33, keeps it
  __device__ __forceinline__ static int lookup(const int16_t* lut, uint32_t window) {
    int e;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(lut) + ((window >> (31 - DEC_LUT_BITS)) & ((2u << DEC_LUT_BITS) - 2u));
    asm("ld.shared.s16 %0, [%1];" : "=r"(e) : "r"(addr));
    return e;
  }
  // a code longer than the table index, computed (VLC.cpp:283-317); needs nb >= 33
  __device__ __forceinline__ int long_vlc(bool& range_err) {
    const uint32_t w = hi;
    const uint32_t f = w & 0xAAAAAAAAu;
    if (f == 0) {   // more than 16 magnitude bits: outside the reference's 32-bit VLC domain
      range_err = true;
      consume32(32);
      return 0;
    }
    const int k = __clz(f) >> 1;            // k >= 1 (k == 0 is in the table)
    const uint32_t t = w >> (32 - 2 * k);   // the 2k leading bits, right aligned
    const uint32_t mag = ((1u << k) | compress16(t)) - 1u;
    const uint32_t neg = (w >> (30 - 2 * k)) & 1u;
    consume32(2 * k + 2);
    return neg ? -(int)mag : (int)mag;
  }
  // one signed interleaved exp-Golomb value; needs nb >= 33 and keeps it
  __device__ __forceinline__ int get_vlc(const int16_t* lut, bool& range_err) {
    const int e = lookup(lut, hi);
    int v;
    if (e != 0) { consume(e & 15); v = e >> 4; }
    else v = long_vlc(range_err);
    ensure();
    return v;
  }
  // two values with one refill: table codes are at most DEC_LUT_BITS = 12 bits, 33 - 2 * 12 >= 1
  __device__ __forceinline__ void get_vlc2(const int16_t* lut, bool& range_err, int& v0, int& v1) {
    const int e0 = lookup(lut, hi);
    if (e0 != 0) { consume(e0 & 15); v0 = e0 >> 4; }
    else { v0 = long_vlc(range_err); ensure(); }
    const int e1 = lookup(lut, hi);
    if (e1 != 0) { consume(e1 & 15); v1 = e1 >> 4; }
    else { ensure(); v1 = long_vlc(range_err); }
    ensure();
  }
  // the same as sign-magnitude words t = 2 * |v| + (v < 0) for the narrow block, straight from the table; `big` collects the
  // magnitudes of the long codes (only those can overflow the narrow word)
  __device__ __forceinline__ static unsigned lookup_sm(const uint16_t* lut, uint32_t window) {
    unsigned e;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(lut) + ((window >> (31 - DEC_LUT_BITS)) & ((2u << DEC_LUT_BITS) - 2u));
    asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(addr));
    return e;
  }
  __device__ __forceinline__ static unsigned smag_of(int v, unsigned& big) {
    const unsigned a = (unsigned)abs(v);
    big |= a;
    return 2u * min(a, (unsigned)VC2_NARROW_MAX_MAG) + ((unsigned)v >> 31);
  }
  __device__ __forceinline__ void get_smag2(const uint16_t* lut, bool& range_err, unsigned& big, unsigned& t0, unsigned& t1) {
    const unsigned e0 = lookup_sm(lut, hi);
    if (e0 != 0u) { consume((int)(e0 & 15u)); t0 = e0 >> 4; }
    else { const int v = long_vlc(range_err); ensure(); t0 = smag_of(v, big); }
    const unsigned e1 = lookup_sm(lut, hi);
    if (e1 != 0u) { consume((int)(e1 & 15u)); t1 = e1 >> 4; }
    else { ensure(); const int v = long_vlc(range_err); t1 = smag_of(v, big); }
    ensure();
  }
  __device__ __forceinline__ uint32_t get_bits(int n) {   // n in 1..32
    const uint32_t w = n == 32 ? hi : (hi >> (32 - n));
    skip(n);
    return w;
  }
};

// ------------------------------------------------------------------------------------------
// HQ / LD slice decoder: ONE THREAD PER SLICE, a warp = one group of 32 consecutive slices.
// Every lane walks its own slice bytes; the decoded (and inverse quantised) coefficients of the
// warp leave as contiguous 512-byte runs of the group-interleaved layout.
// ------------------------------------------------------------------------------------------
template <bool LD, bool DEQ>
__global__ void __launch_bounds__(128) slice_unpack_kernel(const UnpackParams p) {
  __shared__ int16_t s_dec[1 << DEC_LUT_BITS];
  stage_table(reinterpret_cast<uint32_t*>(s_dec), reinterpret_cast<const uint32_t*>(d_dec_lut), (1 << DEC_LUT_BITS) / 2);
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  const int nc4 = g.comp_start[3] >> 2;
  int4* dst = reinterpret_cast<int4*>(p.coef + (long long)pic * g.coef_pic_stride) + (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  const uint32_t* so = p.slice_off + (long long)pic * p.slice_off_pic_stride;
  const uint8_t* bytes = p.in + (long long)pic * p.in_pic_stride + so[s];
  const int size = (int)(so[s + 1] - so[s]);
  const long long sidx = (long long)pic * nslices + s;
  unsigned flags = 0;
  bool range_err = false, badq = false;
  BitReader br;
  int qi;

  if (!LD) {
    // prefix | qindex | len | data | len | data | len | data   (Slices.cpp:544-605)
    bool bad = size < g.prefix + 4;
    qi = bad ? 0 : bytes[g.prefix];
    int pos = g.prefix + 1;
    for (int c = 0; c < 3; ++c) {
      int len = 0;
      if (!bad) {
        len = bytes[pos] * g.scalar;
        if (pos + 1 + len + (2 - c) > size) { bad = true; len = 0; }
      }
      const int start = pos + 1;
      pos = start + len;
      br.init(bytes + (bad ? 0 : start), 8 * len);
      int k = 0, b = 0, bend = g.band_start[c][1];
      BandP bp = band_params(qi, g.qmatrix[0], badq);
      int4* cdst = dst + (size_t)(g.comp_start[c] >> 2) * 32;
      const int n = g.band_start[c][g.nbands];
      const int np = n >> 2;
      int piece = 0;
      while (piece < np) {
        while (k == bend) {   // coefficient k exists (piece < np), so does its band
          ++b;
          bend = g.band_start[c][b + 1];
          bp = band_params(qi, g.qmatrix[b], badq);
        }
        // warp uniform: the whole pieces inside the current band are one inner loop (see walk_component)
        const int run = min((bend - k) >> 2, np - piece);
        if (run > 0) {
          k += 4 * run;
          piece += run;
          for (int i = 0; i < run; ++i) {
            int v[4];
            br.get_vlc2(s_dec, range_err, v[0], v[1]);
            br.get_vlc2(s_dec, range_err, v[2], v[3]);
            if (DEQ) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = scale_band(v[e], bp);
            }
            *cdst = make_int4(v[0], v[1], v[2], v[3]);
            cdst += 32;
          }
        } else {
          int v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            while (k == bend) {
              ++b;
              bend = g.band_start[c][b + 1];
              bp = band_params(qi, g.qmatrix[b], badq);
            }
            const int x = br.get_vlc(s_dec, range_err);
            v[e] = DEQ ? scale_band(x, bp) : x;
            ++k;
          }
          *cdst = make_int4(v[0], v[1], v[2], v[3]);
          cdst += 32;
          ++piece;
        }
      }
    }
    if (bad) flags |= VC2_FLAG_STREAM;
  } else {
    // qindex (7 bits) | luma length | luma (bounded) | U/V interleaved (bounded)   (Slices.cpp:253-296)
    br.init(bytes, 8 * size);
    qi = (int)br.get_bits(7);
    int lb = 0;   // utils::intlog2(8*bytes-7)
    { const int v = 8 * size - 7; while ((1 << lb) < v) ++lb; }
    const int ybits = (int)br.get_bits(lb);
    const int uvbits = 8 * size - 7 - lb - ybits;
    if (uvbits < 0) flags |= VC2_FLAG_STREAM;
    for (int blk = 0; blk < 2; ++blk) {
      // the two blocks start and end on arbitrary bits: restart the reader at the byte holding the first bit
      const int bstart = blk == 0 ? 7 + lb : 7 + lb + ybits;
      const int rem = blk == 0 ? ybits : max(uvbits, 0);
      br.init(bytes + min(bstart >> 3, size), (bstart & 7) + rem);
      br.skip(bstart & 7);
      const int c = blk;   // band geometry: Y, or chroma (U and V are alike)
      int k = 0, b = 0, bend = g.band_start[c][1];
      BandP bp = band_params(qi, g.qmatrix[0], badq);
      const int n = g.band_start[c][g.nbands];
      int4* d0 = dst + (size_t)(g.comp_start[c] >> 2) * 32;
      int4* d1 = dst + (size_t)(g.comp_start[2] >> 2) * 32;
      for (int piece = 0; piece < (n >> 2); ++piece) {
        int v[4], w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          while (k == bend) {
            ++b;
            bend = g.band_start[c][b + 1];
            bp = band_params(qi, g.qmatrix[b], badq);
          }
          const bool deq = DEQ && b != 0;   // the LL band is reconstructed by the DC prediction kernel
          const int x = br.get_vlc(s_dec, range_err);
          v[e] = deq ? scale_band(x, bp) : x;
          if (blk == 1) {   // u0 v0 u1 v1 ... (Slices.cpp:287-294)
            const int y = br.get_vlc(s_dec, range_err);
            w[e] = deq ? scale_band(y, bp) : y;
          }
          ++k;
        }
        d0[(size_t)piece * 32] = make_int4(v[0], v[1], v[2], v[3]);
        if (blk == 1) d1[(size_t)piece * 32] = make_int4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  p.qidx[sidx] = qi;
  if (badq && DEQ) flags |= VC2_FLAG_QUANT_INDEX;
  if (range_err) flags |= VC2_FLAG_VLC_RANGE;
  if (flags) atomicOr(&p.err_flags[sidx], flags);
}

// ------------------------------------------------------------------------------------------
// HQ slice decoder into the NARROW coefficient block: the parsed (still quantised) coefficients are stored as 16-bit
// sign-magnitude words, the inverse lifting kernels scale them on the way in.  A magnitude that does not fit 15 bits
// raises narrow_ovf[picture]: the caller decodes that picture again through the 32-bit path.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) slice_unpack_narrow_kernel(const UnpackParams p) {
  __shared__ uint16_t s_dec[1 << DEC_LUT_BITS];
  stage_table(reinterpret_cast<uint32_t*>(s_dec), reinterpret_cast<const uint32_t*>(d_dec_lut_sm), (1 << DEC_LUT_BITS) / 2);
  __syncthreads();
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  const int nc4 = g.comp_start[3] >> 2;
  uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.coef) + (long long)pic * g.coef_pic_stride) + (size_t)(s >> 5) * nc4 * 32 + (s & 31);
  const uint32_t* so = p.slice_off + (long long)pic * p.slice_off_pic_stride;
  const uint8_t* bytes = p.in + (long long)pic * p.in_pic_stride + so[s];
  const int size = (int)(so[s + 1] - so[s]);
  const long long sidx = (long long)pic * nslices + s;
  unsigned flags = 0, big = 0;
  bool range_err = false;
  BitReader br;
  // prefix | qindex | len | data | len | data | len | data   (Slices.cpp:544-605)
  bool bad = size < g.prefix + 4;
  const int qi = bad ? 0 : bytes[g.prefix];
  int pos = g.prefix + 1;
  for (int c = 0; c < 3; ++c) {
    int len = 0;
    if (!bad) {
      len = bytes[pos] * g.scalar;
      if (pos + 1 + len + (2 - c) > size) { bad = true; len = 0; }
    }
    const int start = pos + 1;
    pos = start + len;
    br.init(bytes + (bad ? 0 : start), 8 * len);
    uint2* cdst = dst + (size_t)(g.comp_start[c] >> 2) * 32;
    const int np = g.band_start[c][g.nbands] >> 2;
    for (int i = 0; i < np; ++i) {
      unsigned t[4];   // the table holds the narrow words themselves: no sign / magnitude conversion for the codes it covers
      br.get_smag2(s_dec, range_err, big, t[0], t[1]);
      br.get_smag2(s_dec, range_err, big, t[2], t[3]);
      *cdst = make_uint2(t[0] | (t[1] << 16), t[2] | (t[3] << 16));
      cdst += 32;
    }
  }
  if (bad) flags |= VC2_FLAG_STREAM;
  if (big > (unsigned)VC2_NARROW_MAX_MAG) atomicOr(p.narrow_ovf + pic, 1u);
  {   // does the whole picture use one index?  (HQ_ConstQ streams: the inverse lifting kernels then scale with per-band constants)
    const uint8_t* first = p.in + (long long)pic * p.in_pic_stride + so[0];
    const int q0 = (int)(so[1] - so[0]) >= g.prefix + 4 ? first[g.prefix] : 0;
    if (qi != q0) atomicOr(&p.band_scale[pic].diff, (unsigned)(qi ^ q0) | 0x100u);
  }
  // the reference scales every band of the slice: an index beyond the quantiser table is an error (Quantisation.cpp:60-63)
  for (int b = 0; b < g.nbands; ++b) if (max(qi - g.qmatrix[b], 0) > 119) flags |= VC2_FLAG_QUANT_INDEX;
  p.qidx[sidx] = qi;
  if (range_err) flags |= VC2_FLAG_VLC_RANGE;
  if (flags) atomicOr(&p.err_flags[sidx], flags);
}

// per picture: the scale factors of every band for the index of slice 0 - what the inverse lifting kernels use when the
// parser found one index for the whole picture (BandScale::diff == 0)
__global__ void narrow_scale_kernel(const UnpackParams p, const uint2* __restrict__ tab) {
  const int pic = blockIdx.x, b = threadIdx.x;
  const int nslices = p.g.slices_x * p.g.slices_y;
  if (b >= p.g.nbands) return;
  const int q = min(max(p.qidx[(long long)pic * nslices] - p.g.qmatrix[b], 0), 127);
  p.band_scale[pic].fo[b] = tab[q];
}

// ------------------------------------------------------------------------------------------
// Slice index of an HQ payload: slice s+1 starts where the three length-prefixed components of slice s
// end (Slices.cpp:544-605), a chain of dependent byte loads.  One CTA per picture: warps 1..7 stream the
// payload through a shared-memory ring while thread 0 walks the chain inside it (three dependent
// shared-memory loads per slice instead of three DRAM round trips).
// A payload that runs off its end leaves the remaining slices empty; the parser flags those.
// ------------------------------------------------------------------------------------------
#ifndef VC2_INDEX_SEG_KB
#define VC2_INDEX_SEG_KB 16
#endif
constexpr int INDEX_SEG = VC2_INDEX_SEG_KB * 1024;    // bytes per ring segment (a slice longer than a segment takes the plain walk in global memory)
constexpr int INDEX_RING = 4 * INDEX_SEG;             // four segments: two being walked, one being filled, one spare
constexpr unsigned INDEX_MASK = INDEX_RING - 1;

__global__ void __launch_bounds__(256) hq_index_kernel(const IndexParams p) {
  // The ring sits at a shared-memory address that is a multiple of its size (the allocation is twice as large, the
  // window starts behind the system's reserved bytes), so that ring address = ring_base | (position & mask) is ONE
  // logic instruction on the walker's dependent chain.  Two stop flags follow the ring.
  extern __shared__ uint4 s_raw[];
  const unsigned raw_base = (unsigned)__cvta_generic_to_shared(s_raw);
  const unsigned ring_base = (raw_base + INDEX_MASK) & ~INDEX_MASK;
  uint4* s_ring = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(s_raw) + (ring_base - raw_base));
  int* s_stop = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_ring) + INDEX_RING);
  const int pic = blockIdx.x, ns = p.nslices;
  const uint8_t* in = p.in + (long long)pic * p.in_pic_stride;
  const unsigned len = p.len_dev ? min(p.len_dev[pic], (unsigned)min(p.in_pic_stride, 0xFFFFFFFFll)) : p.len[pic];
  uint32_t* off = p.slice_off + (long long)pic * (ns + 1);
  const unsigned maxslice = (unsigned)(p.prefix + 4 + 3 * 255 * p.scalar);
  const long long readable = p.in_pic_stride & ~15ll;   // the picture's buffer, whole 16-byte pieces
  auto load_seg = [&](unsigned k, int first, int nthreads) {   // segment k of the payload -> ring slot k % 4
    const long long b0 = (long long)k * INDEX_SEG;
    if (b0 >= readable) return;
    const int n16 = (int)(min((long long)INDEX_SEG, readable - b0) >> 4);
    const uint4* src = reinterpret_cast<const uint4*>(in + b0);
    uint4* dst = s_ring + (size_t)(k & 3) * (INDEX_SEG / 16);
    for (int i = first; i < n16; i += nthreads) dst[i] = __ldg(src + i);
  };
  // the same as an asynchronous copy (LDGSTS): issued one segment earlier than it is needed, so that the barrier
  // that ends an iteration never waits for a DRAM round trip that has only just started
  auto load_seg_async = [&](unsigned k, int first, int nthreads) {
    const long long b0 = (long long)k * INDEX_SEG;
    if (b0 < readable) {
      const int n16 = (int)(min((long long)INDEX_SEG, readable - b0) >> 4);
      const uint4* src = reinterpret_cast<const uint4*>(in + b0);
      const unsigned dst = ring_base + (k & 3u) * INDEX_SEG;
      for (int i = first; i < n16; i += nthreads)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one group per segment and thread, also when it is empty
  };
  if (maxslice > (unsigned)INDEX_SEG) {
    // slices may be longer than a ring segment (huge scalar): plain walk in global memory
    if (threadIdx.x == 0) {
      unsigned pos = 0;
      int s = 0;
      for (; s < ns; ++s) {
        unsigned q = pos + p.prefix + 1;
        bool bad = false;
        for (int c = 0; c < 3 && !bad; ++c) {
          if (q >= len) bad = true; else q += 1u + (unsigned)in[q] * (unsigned)p.scalar;
        }
        if (bad || q > len) break;
        off[s] = pos;
        pos = q;
      }
      for (int t = s; t <= ns; ++t) off[t] = pos;
    }
    return;
  }
  // ring slots while segment k is walked: k, k + 1 valid | k + 2 landing | k + 3 being issued (the slot of k - 1)
  load_seg(0, threadIdx.x, blockDim.x);
  load_seg(1, threadIdx.x, blockDim.x);
  if (threadIdx.x >= 32) load_seg_async(2, threadIdx.x - 32, blockDim.x - 32);   // only the threads that later wait for it
  __syncthreads();
  unsigned pos = 0;   // walker state (thread 0)
  int s = 0;
  for (unsigned k = 0;; ++k) {
    if (threadIdx.x >= 32) {
      load_seg_async(k + 3, threadIdx.x - 32, blockDim.x - 32);     // warps 1..7 fill the ring ahead of the walker
      asm volatile("cp.async.wait_group 1;" ::: "memory");          // segment k + 2 has landed
    } else if (threadIdx.x == 0) {
      // every slice that STARTS in segment k; its bytes end before segment k + 2 (maxslice <= INDEX_SEG)
      const uint8_t* w = reinterpret_cast<const uint8_t*>(s_ring);
      const unsigned seg_end = (k + 1) * INDEX_SEG;
      const unsigned hdr = p.prefix + 1, sc = p.scalar;
      bool bad = false;
      // fast path: four slices per loop-carried branch.  The walk is one chain of dependent shared-memory loads
      // (three per slice); a branch per slice would add its resolution to every link, and so would any address
      // arithmetic: with the ring aligned to its size a link is  load -> multiply-add -> mask|base -> load.
      // a = stream position of the next length byte.  The positions may run past the two valid ring segments -
      // the reads are masked into the ring and a group is only committed when its fourth slice starts inside
      // segment k, in which case every byte it looked at was valid.
      {
        auto hop = [&](unsigned a, unsigned add) {
          unsigned b;
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"((a & INDEX_MASK) | ring_base));
          unsigned r;   // a + add is ready before the byte arrives: keep it out of the chain
          asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(sc), "r"(a + add));
          return r;
        };
        unsigned a = pos + hdr;
        while (s + 4 <= ns) {
          const unsigned a1 = hop(hop(hop(a, 1u), 1u), 1u + hdr);    // first length byte of slice s + 1
          const unsigned a2 = hop(hop(hop(a1, 1u), 1u), 1u + hdr);
          const unsigned a3 = hop(hop(hop(a2, 1u), 1u), 1u + hdr);
          const unsigned a4 = hop(hop(hop(a3, 1u), 1u), 1u + hdr);
          if (a3 - hdr >= seg_end || a4 - hdr > len) break;
          off[s] = a - hdr; off[s + 1] = a1 - hdr; off[s + 2] = a2 - hdr; off[s + 3] = a3 - hdr;
          a = a4;
          s += 4;
        }
        pos = a - hdr;
      }
      while (s < ns && pos < seg_end) {
        unsigned q = pos + hdr;
        q += 1u + w[q & INDEX_MASK] * sc;
        q += 1u + w[q & INDEX_MASK] * sc;
        q += 1u + w[q & INDEX_MASK] * sc;
        if (q > len) { bad = true; break; }   // also catches a length byte at or beyond len: q only grows
        off[s] = pos;
        pos = q;
        ++s;
      }
      if (bad) {
        for (int t = s; t <= ns; ++t) off[t] = pos;   // empty slices: the parser raises VC2_FLAG_STREAM
        s = ns + 1;
      } else if (s == ns) {
        off[ns] = pos;
        s = ns + 1;
      }
      s_stop[k & 1] = s > ns;
    }
    __syncthreads();
    if (s_stop[k & 1]) break;
  }
}

// ------------------------------------------------------------------------------------------
// stand-alone quantise / dequantise of an in-place ordered plane (Library surface)
// ------------------------------------------------------------------------------------------
__global__ void quant_kernel(const QuantParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= p.pw || y >= p.ph) return;
  const int d = p.depth;
  const int t = (y | x) & ((1 << d) - 1);
  int band = 0;
  if (t) {
    const int l = __ffs(t) - 1, L = d - l;
    const int hx = (x >> l) & 1, hy = (y >> l) & 1;
    band = 3 * (L - 1) + (hx ? (hy ? 3 : 1) : 2);
  }
  const long long i = (long long)y * p.pw + x;
  if (band == 0 && p.skip_ll) { p.dst[i] = p.src[i]; return; }
  const int sy = y / (p.ph / p.slices_y), sx = x / (p.pw / p.slices_x);
  const int q = max(p.qidx[sy * p.slices_x + sx] - p.qmatrix[band], 0);
  const QParam qp = qparam(q);
  const int v = p.src[i];
  p.dst[i] = p.inverse ? scale_one(v, qp.qf, qp.qo) : quant_one(v, qp.qm, qp.ql);
}

// ------------------------------------------------------------------------------------------
// LD LL-band reconstruction: rec[y][x] = scale(c, q'(slice)) + predictDC(rec, y, x), raster
// recurrence over the whole band (Quantisation.cpp:191-208, 287-306).  One CTA per plane runs
// the anti-diagonal wavefront: element (y, x) depends on (y-1, x-1), (y-1, x), (y, x-1).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_dc_body(LdDcParams p) {
  const int H = p.H, Wd = p.W;
  auto at = [&](int y, int x) -> int32_t* {
    if (!p.interleaved) return p.base + (((long long)y * p.pitch + x) << p.depth);
    const int sy = y / p.bh, sx = x / p.bw;
    return p.base + coef_index(sy * p.slices_x + sx, p.k0 + (y - sy * p.bh) * p.bw + (x - sx * p.bw), p.nc4);
  };
  for (int diag = 0; diag < H + Wd - 1; ++diag) {
    const int ylo = max(0, diag - (Wd - 1)), yhi = min(H - 1, diag);
    for (int y = ylo + (int)threadIdx.x; y <= yhi; y += blockDim.x) {
      const int x = diag - y;
      int32_t* e = at(y, x);
      const int yb = ((y + 1) * p.slices_y - 1) / H, xb = ((x + 1) * p.slices_x - 1) / Wd;
      const int q = max(p.qidx[yb * p.slices_x + xb] - p.qm0, 0);
      const QParam qp = qparam(q);
      int pred;
      if (y > 0 && x > 0) {
        const int sum = *at(y - 1, x - 1) + *at(y - 1, x) + *at(y, x - 1);
        pred = sum >= 0 ? (sum + 1) / 3 : (sum - 1) / 3;
      } else if (y > 0) pred = *at(y - 1, x);
      else if (x > 0) pred = *at(y, x - 1);
      else pred = 0;
      *e = scale_one(*e, qp.qf, qp.qo) + pred;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) ld_dc_kernel(const LdDcParams p) { ld_dc_body(p); }
// ------------------------------------------------------------------------------------------
// Library surface: code-list bits of every slice of an in-place ordered quantised plane (one thread per slice;
// the slice's bands in coding order, raster inside the slice's part of each band - split_into_subbands of the slice)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int signed_vlc_bits(int v) {   // SignedVLC(v).numOfBits(), VLC.cpp:21-52, 78-85
  if (v == 0) return 1;
  const uint32_t m = (uint32_t)abs(v) + 1u;
  return 2 * (31 - __clz(m)) + 2;
}
__global__ void slice_bits_kernel(const SliceBitsParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.slices_y * p.slices_x) return;
  const int sh = p.ph / p.slices_y, sw = p.pw / p.slices_x;
  const int y0 = (s / p.slices_x) * sh, x0 = (s % p.slices_x) * sw;
  int gross = 0, count = 0;
  for (int band = 0; band < 3 * p.depth + 1; ++band) {
    const int level = band == 0 ? 0 : (band - 1) / 3 + 1;
    const int stride = band == 0 ? (1 << p.depth) : (1 << (p.depth + 1 - level));
    const int kind = band == 0 ? -1 : (band - 1) % 3;              // 0 HL, 1 LH, 2 HH
    const int oy = (kind == 1 || kind == 2) ? stride / 2 : 0, ox = (kind == 0 || kind == 2) ? stride / 2 : 0;
    for (int y = y0 + oy; y < y0 + sh; y += stride)
      for (int x = x0 + ox; x < x0 + sw; x += stride) {
        const long long i = (long long)y * p.pw + x;
        int nb = signed_vlc_bits(p.q[i]);
        gross += nb;
        if (nb > 1) count = gross;
        if (p.q2) {
          nb = signed_vlc_bits(p.q2[i]);
          gross += nb;
          if (nb > 1) count = gross;
        }
      }
  }
  p.bits[s] = count;
}

// forward counterpart of ld_dc_kernel: anti-diagonal wavefront, the prediction comes from the locally decoded band
__global__ void __launch_bounds__(1024) ld_dc_quant_kernel(const LdDcQuantParams p) {
  const int H = p.H, Wd = p.W;
  for (int diag = 0; diag < H + Wd - 1; ++diag) {
    const int ylo = max(0, diag - (Wd - 1)), yhi = min(H - 1, diag);
    for (int y = ylo + (int)threadIdx.x; y <= yhi; y += blockDim.x) {
      const int x = diag - y;
      const int yb = ((y + 1) * p.slices_y - 1) / H, xb = ((x + 1) * p.slices_x - 1) / Wd;
      const int q = max(p.qidx[yb * p.slices_x + xb] - p.qm0, 0);
      const QParam qp = qparam(q);
      const int32_t* r = p.restored;
      int pred;
      if (y > 0 && x > 0) {
        const int sum = r[(y - 1) * Wd + x - 1] + r[(y - 1) * Wd + x] + r[y * Wd + x - 1];
        pred = sum >= 0 ? (sum + 1) / 3 : (sum - 1) / 3;
      } else if (y > 0) pred = r[(y - 1) * Wd + x];
      else if (x > 0) pred = r[y * Wd + x - 1];
      else pred = 0;
      const long long i = ((long long)y * p.pitch + x) << p.depth;
      const int qv = quant_one(p.src[i] - pred, qp.qm, qp.ql);
      p.dst[i] = qv;
      p.restored[y * Wd + x] = scale_one(qv, qp.qf, qp.qo) + pred;
    }
    __syncthreads();
  }
}

// one CTA per (picture, plane): the wavefronts of a batch run side by side
__global__ void __launch_bounds__(1024) ld_dc_batch_kernel(const LdDcBatch b) {
  LdDcParams p = b.c[blockIdx.y];
  p.base += (long long)blockIdx.x * p.base_pic_stride;
  p.qidx += (long long)blockIdx.x * p.qidx_pic_stride;
  ld_dc_body(p);
}

}  // namespace

// VC2_SEARCH_SMEM=1: rate control with the slices in shared memory (hq_search_kernel).  Bit exact, MEASURED SLOWER and therefore
// off: 11.5 against 7.95 ms per 256 C2 pictures (profiles/r2_v6_search_smem_ab.txt) - 64 KB per 128 slices leaves 12 warps per
// SM where the streaming kernel has 32, and the probe arithmetic is a dependent chain per coefficient that needs the warps.
// VC2_SEARCH_WARP=1: rate control with a warp per slice and the slice in registers (hq_search_warp_kernel).  Bit exact, MEASURED
// SLOWER and therefore off: 18.2 against 7.95 ms per 256 C2 pictures (profiles/r2_v6_search_smem_ab.txt).  With 256 coefficients a
// lane holds eight, and everything that is per slice rather than per coefficient - the band table, the reductions, the byte
// rounding with its division - costs a whole warp instruction for ONE slice instead of 32: 6400 warp instructions per slice
// against 3100 (ncu, gpurun_out/r2_g36_search.ncu-rep), 15 % of them the arithmetic the kernel exists for.
// VC2_SEARCH_SUB=1: rate control with eight lanes per slice and the slice in registers (hq_search_sub_kernel).  Bit exact, MEASURED
// SLOWER and therefore off: 11.25 against 7.95 ms per 256 C2 pictures.  Same instruction count as the thread-per-slice kernel (1.54 G
// against 1.64 G warp instructions per 64 pictures, ncu) at 128 registers and 16 warps per SM: IPC 1.9 against 2.75.
static bool search_in_subwarps() {
  static const bool on = getenv("VC2_SEARCH_SUB") && atoi(getenv("VC2_SEARCH_SUB")) != 0;
  return on;
}
static bool search_in_registers() {
  static const bool on = getenv("VC2_SEARCH_WARP") && atoi(getenv("VC2_SEARCH_WARP")) != 0;
  return on;
}
static bool search_in_smem() {
  static const bool on = getenv("VC2_SEARCH_SMEM") && atoi(getenv("VC2_SEARCH_SMEM")) != 0;
  return on;
}

cudaError_t pack_launch(cudaStream_t s, const PackParams& p, int npictures) {
  const int nslices = p.g.slices_x * p.g.slices_y;
  if (p.narrow) {
    if (p.search || p.const_q < 0 || !p.emit || p.mode != VC2_HQ_VBR) return cudaErrorInvalidValue;
    if (p.fuse && p.tiles != (nslices + 127) / 128) return cudaErrorInvalidValue;
    hq_pack_narrow_kernel<<<dim3((nslices + 127) / 128, npictures), 128, 0, s>>>(p);
    return cudaGetLastError();
  }
  if (p.search && !p.emit && search_in_subwarps()) {
    // rate control alone, eight lanes per slice with the slice in registers: up to eight pieces per lane, and every slot of eight
    // consecutive pieces inside one component
    const int pieces = p.g.comp_start[3] >> 2;
    const bool whole = (p.g.comp_start[1] >> 2) % 8 == 0 && (p.g.comp_start[2] >> 2) % 8 == 0;
    if (whole && pieces <= 64) {
      const dim3 grid((nslices + 31) / 32, npictures);
      const int r = (pieces + 7) / 8;
      if (r <= 2) hq_search_sub_kernel<2><<<grid, 256, 0, s>>>(p);
      else if (r <= 4) hq_search_sub_kernel<4><<<grid, 256, 0, s>>>(p);
      else hq_search_sub_kernel<8><<<grid, 256, 0, s>>>(p);
      return cudaGetLastError();
    }
  }
  if (p.search && !p.emit && search_in_registers()) {
    // rate control alone, a warp per slice with the slice in registers: R pieces per lane
    const int pieces = (p.g.comp_start[3] >> 2);
    const int r = (pieces + 31) / 32;
    const dim3 grid((nslices + 7) / 8, npictures);
    if (r <= 1) { hq_search_warp_kernel<1><<<grid, 256, 0, s>>>(p); return cudaGetLastError(); }
    if (r <= 2) { hq_search_warp_kernel<2><<<grid, 256, 0, s>>>(p); return cudaGetLastError(); }
    if (r <= 4) { hq_search_warp_kernel<4><<<grid, 256, 0, s>>>(p); return cudaGetLastError(); }
    if (r <= 8) { hq_search_warp_kernel<8><<<grid, 256, 0, s>>>(p); return cudaGetLastError(); }
    // larger slices: the thread-per-slice kernel below
  }
  if (p.search && !p.emit && search_in_smem()) {
    // rate control alone: the slices of a CTA in shared memory when they fit (16-bit magnitudes; 128, 64 or 32 slices per CTA)
    const size_t per_slice = (size_t)p.g.comp_start[3] * 2;
    int spc = 0;
    for (int t = 128; t >= 32 && !spc; t >>= 1)
      if (per_slice * t <= (size_t)(t == 32 ? 160 : 72) * 1024) spc = t;
    if (spc) {
      const size_t smem = per_slice * spc;
      cudaError_t e = cudaFuncSetAttribute(hq_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      hq_search_kernel<<<dim3((nslices + spc - 1) / spc, npictures), spc, smem, s>>>(p);
      return cudaGetLastError();
    }
  }
  hq_pack_kernel<<<dim3((nslices + 127) / 128, npictures), 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t ld_encode_launch(cudaStream_t s, const LdEncParams& p, int npictures) {
  const int nslices = p.g.slices_x * p.g.slices_y;
  ld_ac_bits_kernel<<<dim3((nslices + 31) / 32, 4, npictures), 128, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  ld_rate_kernel<<<npictures, 256, 0, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  ld_pack_kernel<<<dim3((nslices + 127) / 128, npictures), 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t ld_pack_launch(cudaStream_t s, const LdEncParams& p, int npictures) {
  if (!p.prequantised) return cudaErrorInvalidValue;
  const int nslices = p.g.slices_x * p.g.slices_y;
  ld_pack_kernel<<<dim3((nslices + 127) / 128, npictures), 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t slice_bits_launch(cudaStream_t s, const SliceBitsParams& p) {
  const int n = p.slices_y * p.slices_x;
  slice_bits_kernel<<<(n + 63) / 64, 64, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t ld_dc_quant_launch(cudaStream_t s, const LdDcQuantParams& p) {
  ld_dc_quant_kernel<<<1, 1024, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t assemble_launch(cudaStream_t s, const AssembleParams& p, int npictures) {
  slice_scan_kernel<<<npictures, 1024, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  assemble_kernel<<<dim3((p.nslices + 7) / 8, npictures), 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t unpack_launch(cudaStream_t s, const UnpackParams& p, int npictures, const uint2* scale_tab) {
  const int nslices = p.g.slices_x * p.g.slices_y;
  const dim3 grid((nslices + 127) / 128, npictures);
  if (p.narrow) {
    if (p.ld) return cudaErrorInvalidValue;
    slice_unpack_narrow_kernel<<<grid, 128, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    narrow_scale_kernel<<<npictures, 32, 0, s>>>(p, scale_tab);
    return cudaGetLastError();
  }
  if (p.ld) {
    if (p.dequantise) slice_unpack_kernel<true, true><<<grid, 128, 0, s>>>(p);
    else slice_unpack_kernel<true, false><<<grid, 128, 0, s>>>(p);
  } else {
    if (p.dequantise) slice_unpack_kernel<false, true><<<grid, 128, 0, s>>>(p);
    else slice_unpack_kernel<false, false><<<grid, 128, 0, s>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t index_launch(cudaStream_t s, const IndexParams& p, int npictures) {
  cudaError_t e = cudaFuncSetAttribute(hq_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * INDEX_RING + 16);   // per device
  if (e != cudaSuccess) return e;
  if (npictures < 1 || (!p.len_dev && npictures > VC2_INDEX_MAX_PICTURES)) return cudaErrorInvalidValue;
  hq_index_kernel<<<npictures, 256, 2 * INDEX_RING + 16, s>>>(p);
  return cudaGetLastError();
}

cudaError_t quant_launch(cudaStream_t s, const QuantParams& p) {
  const dim3 block(32, 8), grid((p.pw + 31) / 32, (p.ph + 7) / 8);
  quant_kernel<<<grid, block, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t ld_dc_batch_launch(cudaStream_t s, const LdDcBatch& b, int npictures) {
  ld_dc_batch_kernel<<<dim3(npictures, b.nplanes), 1024, 0, s>>>(b);
  return cudaGetLastError();
}

cudaError_t ld_dc_launch(cudaStream_t s, const LdDcParams& p) {
  ld_dc_kernel<<<1, 1024, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace vc2
