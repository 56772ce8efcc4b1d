// Device helpers shared by the lifting kernels (dwt.cu: streaming register-ring kernels, dwt_tile.cu: shared-memory
// tile kernels): the 1-D lifting step in registers with neighbours by warp shuffle, raw sample conversion, and the
// addressing of the group-interleaved coefficient block.  See dwt.cuh for the reference line citations.
#pragma once
#include "dwt.cuh"
#include "slices.cuh"

namespace vc2 {
namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
__host__ __device__ constexpr int pmod(int a, int m) { return ((a % m) + m) % m; }

// ---- horizontal lifting step in registers ------------------------------------------------------
// A lane holds pairs PPL*lane .. PPL*lane+PPL-1 of a row segment: e[a] / o[a] = even / odd sample of
// pair PPL*lane+a.  Neighbouring pairs come from other lanes by shuffle.  When the segment touches
// the left/right edge of the lattice (hedge), the source-parity sequence is first extended beyond
// [plo, phi] with its edge value = the reference's tap clamping.
template <int PPL>
__device__ __forceinline__ int pick(const int (&x)[PPL], int i) {
  int v = x[0];
#pragma unroll
  for (int a = 1; a < PPL; ++a) v = (i == a) ? x[a] : v;
  return v;
}

// One lifting update, x[t] += SIGN * ((ADD + sum_k CLk*x[t-(2k+1)] + CRk*x[t+(2k+1)]) >> SH)  (forward; the inverse
// flips the sign).  A subtraction is folded into the sum: -(s >> n) == (2^n - 1 - s) >> n for an arithmetic shift, so
// every update is "multiply-add chain, shift, add" (the shift-and-add is one LEA.HI.SX32).
// c + (sa ? -a : a) + (sb ? -b : b) as one three-input add: the asm block keeps the compiler from re-associating the
// rounding constant away from the two unit-weight taps
template <bool SA, bool SB>
__device__ __forceinline__ int add3(int c, int a, int b) {
  int r;
  if (SA && SB) asm("{ .reg .s32 t; sub.s32 t, %1, %2; sub.s32 %0, t, %3; }" : "=r"(r) : "r"(c), "r"(a), "r"(b));
  else if (SA) asm("{ .reg .s32 t; sub.s32 t, %1, %2; add.s32 %0, t, %3; }" : "=r"(r) : "r"(c), "r"(a), "r"(b));
  else if (SB) asm("{ .reg .s32 t; add.s32 t, %1, %2; sub.s32 %0, t, %3; }" : "=r"(r) : "r"(c), "r"(a), "r"(b));
  else asm("{ .reg .s32 t; add.s32 t, %1, %2; add.s32 %0, t, %3; }" : "=r"(r) : "r"(c), "r"(a), "r"(b));
  return r;
}

template <int K, int S, int DIR, class Tap>
__device__ __forceinline__ int lift_update(int t, Tap tap) {   // tap(k, right) = source value at distance 2k+1 to the left / right
  using ST = Step<K, S>;
  constexpr bool NEG = ST::SIGN * DIR < 0;
  constexpr int ADD = NEG ? ((1 << ST::SH) - 1 - ST::ADD) : ST::ADD;
  // taps of weight +-1 join the rounding constant (one three-input add per pair), the others are multiply-adds
  int u = ADD;
#pragma unroll
  for (int k = 0; k < ST::N; ++k) {
    const int cl = NEG ? -ST::cl(k) : ST::cl(k), cr = NEG ? -ST::cr(k) : ST::cr(k);
    const bool ul = cl == 1 || cl == -1, ur = cr == 1 || cr == -1;
    if (ul && ur) {
      if (cl < 0 && cr < 0) u = add3<true, true>(u, tap(k, false), tap(k, true));
      else if (cl < 0) u = add3<true, false>(u, tap(k, false), tap(k, true));
      else if (cr < 0) u = add3<false, true>(u, tap(k, false), tap(k, true));
      else u = add3<false, false>(u, tap(k, false), tap(k, true));
    } else {
      if (cl == 1) u += tap(k, false); else if (cl == -1) u -= tap(k, false);
      if (cr == 1) u += tap(k, true); else if (cr == -1) u -= tap(k, true);
    }
  }
#pragma unroll
  for (int k = 0; k < ST::N; ++k) {
    const int cl = NEG ? -ST::cl(k) : ST::cl(k), cr = NEG ? -ST::cr(k) : ST::cr(k);
    if (cl == cr && cl != 0 && cl != 1 && cl != -1) u = (tap(k, false) + tap(k, true)) * cl + u;
    else {
      if (cl != 0 && cl != 1 && cl != -1) u = tap(k, false) * cl + u;
      if (cr != 0 && cr != 1 && cr != -1) u = tap(k, true) * cr + u;
    }
  }
  return t + (u >> ST::SH);
}

// WIDTH: lanes per row segment (32, or 16 when a warp works on two segments at once); `lane` is the lane inside its segment
template <int K, int S, int DIR, int PPL, int WIDTH = 32>
__device__ __forceinline__ void hstep(int (&e)[PPL], int (&o)[PPL], int lane, bool hedge, int plo, int phi) {
  using ST = Step<K, S>;
  constexpr int P = ST::P, N = ST::N;
  int (&src)[PPL] = P ? e : o;
  int (&tgt)[PPL] = P ? o : e;
  if (hedge) {
    const int vlo = __shfl_sync(FULL, pick<PPL>(src, plo % PPL), plo / PPL, WIDTH);
    const int vhi = __shfl_sync(FULL, pick<PPL>(src, phi % PPL), phi / PPL, WIDTH);
#pragma unroll
    for (int a = 0; a < PPL; ++a) {
      const int p = PPL * lane + a;
      if (p < plo) src[a] = vlo; else if (p > phi) src[a] = vhi;
    }
  }
  constexpr int QMIN = -(P ? N - 1 : N), QMAX = PPL - 1 + (P ? N : N - 1);
  int nb[QMAX - QMIN + 1];   // source values at pair offsets QMIN..QMAX from pair PPL*lane
#pragma unroll
  for (int q = QMIN; q <= QMAX; ++q) {
    const int r = pmod(q, PPL), dl = (q - r) / PPL;
    nb[q - QMIN] = dl == 0 ? src[r] : __shfl_sync(FULL, src[r], lane + dl, WIDTH);
  }
#pragma unroll
  for (int a = 0; a < PPL; ++a)
    tgt[a] = lift_update<K, S, DIR>(tgt[a], [&](int k, bool right) { return right ? nb[a + (P ? k + 1 : k) - QMIN] : nb[a - (P ? k : k + 1) - QMIN]; });
}

template <int K, int DIR, int PPL, int WIDTH = 32>
__device__ __forceinline__ void hsteps(int (&e)[PPL], int (&o)[PPL], int lane, bool hedge, int plo, int phi) {
  constexpr int N = Wavelet<K>::NSTEPS;
  if constexpr (DIR > 0) {
    hstep<K, 0, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
    hstep<K, 1, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
    if constexpr (N == 4) {
      hstep<K, 2, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
      hstep<K, 3, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
    }
  } else {
    if constexpr (N == 4) {
      hstep<K, 3, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
      hstep<K, 2, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
    }
    hstep<K, 1, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
    hstep<K, 0, DIR, PPL, WIDTH>(e, o, lane, hedge, plo, phi);
  }
}

// The same for NR independent row segments at once (the tile kernels lift several rows / columns per call so that the
// shuffles of one segment overlap the arithmetic of another): e[r] / o[r] = pairs of segment r.
template <int K, int S, int DIR, int PPL, int WIDTH, int NR>
__device__ __forceinline__ void hstep_rows(int (&e)[NR][PPL], int (&o)[NR][PPL], int lane, bool hedge, int plo, int phi) {
  using ST = Step<K, S>;
  constexpr int P = ST::P, N = ST::N;
  int (&src)[NR][PPL] = P ? e : o;
  int (&tgt)[NR][PPL] = P ? o : e;
  if (hedge) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int vlo = __shfl_sync(FULL, pick<PPL>(src[r], plo % PPL), plo / PPL, WIDTH);
      const int vhi = __shfl_sync(FULL, pick<PPL>(src[r], phi % PPL), phi / PPL, WIDTH);
#pragma unroll
      for (int a = 0; a < PPL; ++a) {
        const int p = PPL * lane + a;
        if (p < plo) src[r][a] = vlo; else if (p > phi) src[r][a] = vhi;
      }
    }
  }
  constexpr int QMIN = -(P ? N - 1 : N), QMAX = PPL - 1 + (P ? N : N - 1);
  int nb[NR][QMAX - QMIN + 1];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int q = QMIN; q <= QMAX; ++q) {
      const int m = pmod(q, PPL), dl = (q - m) / PPL;
      nb[r][q - QMIN] = dl == 0 ? src[r][m] : __shfl_sync(FULL, src[r][m], lane + dl, WIDTH);
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int a = 0; a < PPL; ++a)
      tgt[r][a] = lift_update<K, S, DIR>(tgt[r][a], [&](int k, bool right) { return right ? nb[r][a + (P ? k + 1 : k) - QMIN] : nb[r][a - (P ? k : k + 1) - QMIN]; });
  }
}

template <int K, int DIR, int PPL, int WIDTH, int NR>
__device__ __forceinline__ void hsteps_rows(int (&e)[NR][PPL], int (&o)[NR][PPL], int lane, bool hedge, int plo, int phi) {
  constexpr int N = Wavelet<K>::NSTEPS;
  if constexpr (DIR > 0) {
    hstep_rows<K, 0, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
    hstep_rows<K, 1, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
    if constexpr (N == 4) {
      hstep_rows<K, 2, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
      hstep_rows<K, 3, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
    }
  } else {
    if constexpr (N == 4) {
      hstep_rows<K, 3, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
      hstep_rows<K, 2, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
    }
    hstep_rows<K, 1, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
    hstep_rows<K, 0, DIR, PPL, WIDTH, NR>(e, o, lane, hedge, plo, phi);
  }
}

// ---- global memory access of one lane's V columns of one row -------------------------------------------
struct BandAddr {   // group-interleaved addressing of band samples (see vc2_common.cuh)
  int bh, bw, lgbh, lgbw, nx, nc4;
  __device__ __forceinline__ long long at(int base, int by, int bx) const {
    const int sy = lgbh >= 0 ? (by >> lgbh) : (by / bh);
    const int sx = lgbw >= 0 ? (bx >> lgbw) : (bx / bw);
    return coef_index(sy * nx + sx, base + (by - sy * bh) * bw + (bx - sx * bw), nc4);
  }
};

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// prefetch distance in row pairs: StripCtx::pd (DwtParams::pd, host tuned)

// forward: touch the cache line(s) of this lane's samples of picture row `row`
template <int KIND, int V>
__device__ __forceinline__ void prefetch_pix(const DwtComp& C, int pic, int row, int gx) {
  if (row > C.pix_h - 1 || gx < 0 || gx + V - 1 >= C.pix_w) return;
  const int esz = KIND == SAMPLE_I32 ? 4 : KIND == SAMPLE_U16BE ? 2 : 1;
  const uint8_t* base = (const uint8_t*)C.pix + (KIND == SAMPLE_I32 ? (long long)pic * C.pix_pic_stride * 4 : (long long)pic * C.pix_pic_stride);
  prefetch_l1(base + ((long long)row * C.pix_pitch + gx) * esz);
}

__device__ __forceinline__ int sample_u16be(unsigned w, int sshift, int soffset) {
  return (int)(__byte_perm(w, 0, 0x4401) >> sshift) - soffset;
}

// forward: load V consecutive samples of picture row `row` starting at column gx (may be negative / past the edge)
template <int KIND, int V>
__device__ __forceinline__ void load_pix(const DwtComp& C, int pic, int row, int gx, int (&x)[V]) {
  const int sy = min(row, C.pix_h - 1);                 // waveletPad: replicate the last row (WaveletTransform.cpp:88)
  const bool inside = gx >= 0 && gx + V - 1 < C.pix_w;   // all V samples are real picture samples
  if (KIND == SAMPLE_I32) {
    const int* rp = (const int*)C.pix + (long long)pic * C.pix_pic_stride + (long long)sy * C.pix_pitch;
    if (inside && ((reinterpret_cast<uintptr_t>(rp + gx) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < V; j += 4) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(rp + gx + j));
        x[j] = q.x; x[j + 1] = q.y; x[j + 2] = q.z; x[j + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) x[j] = rp[min(max(gx + j, 0), C.pix_w - 1)];   // :89 replicate the last column
    }
  } else if (KIND == SAMPLE_U16BE) {
    const uint16_t* rp = (const uint16_t*)((const uint8_t*)C.pix + (long long)pic * C.pix_pic_stride) + (long long)sy * C.pix_pitch;
    if (inside && ((reinterpret_cast<uintptr_t>(rp + gx) & (2 * V - 1)) == 0)) {
      unsigned w[V / 2];
      if (V == 8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp + gx));
        w[0] = q.x; w[1] = q.y; w[V / 2 - 2] = q.z; w[V / 2 - 1] = q.w;
      } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(rp + gx));
        w[0] = q.x; w[1] = q.y;
      }
#pragma unroll
      for (int j = 0; j < V / 2; ++j) {
        x[2 * j] = sample_u16be(w[j] & 0xFFFFu, C.sshift, C.soffset);
        x[2 * j + 1] = (int)(__byte_perm(w[j], 0, 0x4423) >> C.sshift) - C.soffset;
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) x[j] = sample_u16be(rp[min(max(gx + j, 0), C.pix_w - 1)], C.sshift, C.soffset);
    }
  } else {
    const uint8_t* rp = (const uint8_t*)C.pix + (long long)pic * C.pix_pic_stride + (long long)sy * C.pix_pitch;
#pragma unroll
    for (int j = 0; j < V; ++j) x[j] = (int)((unsigned)rp[min(max(gx + j, 0), C.pix_w - 1)] >> C.sshift) - C.soffset;
  }
}

// inverse: V consecutive reconstructed samples of picture row `row` starting at column gx (inside the picture rows; columns
// beyond pix_w are dropped): accuracy-bit rounding (WaveletTransform.cpp:338), clip (Picture.cpp:284-292) and the raw sample
// format (Arrays.cpp:396-414: offset binary, MSB justified, big endian)
template <int K, int KIND, int V>
__device__ __forceinline__ void store_pix(const DwtComp& C, int pic, int row, int gx, int (&v)[V]) {
  constexpr int SHIFT = Wavelet<K>::SHIFT;
  const int rnd = SHIFT ? (1 << (SHIFT - 1)) : 0;
  const int lo = C.clip_min + C.soffset, hi = C.clip_max + C.soffset;   // clip(v) + offset == min(max(v + offset, lo), hi): add-and-max is one instruction
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (SHIFT) v[k] = (v[k] + rnd) >> SHIFT;
    if (KIND != SAMPLE_I32) v[k] = (int)((unsigned)min(__viaddmax_s32(v[k], C.soffset, lo), hi) << C.sshift);
  }
  const long long rowoff = (long long)row * C.pix_pitch;
  const bool whole = gx + V - 1 < C.pix_w;
  if (KIND == SAMPLE_I32) {
    int* dst = (int*)C.pix + (long long)pic * C.pix_pic_stride + rowoff + gx;
    if (whole && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < V; j += 4) *reinterpret_cast<int4*>(dst + j) = make_int4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) if (gx + k < C.pix_w) dst[k] = v[k];
    }
  } else if (KIND == SAMPLE_U16BE) {
    uint16_t* dst = (uint16_t*)((uint8_t*)C.pix + (long long)pic * C.pix_pic_stride) + rowoff + gx;
    unsigned w[V / 2];
#pragma unroll
    for (int j = 0; j < V / 2; ++j) w[j] = __byte_perm((unsigned)v[2 * j], (unsigned)v[2 * j + 1], 0x4501);
    if (whole && ((reinterpret_cast<uintptr_t>(dst) & (2 * V - 1)) == 0)) {
      if (V == 8) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[V / 2 - 2], w[V / 2 - 1]);
      else *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) if (gx + k < C.pix_w) dst[k] = (uint16_t)((k & 1) ? (w[k / 2] >> 16) : (w[k / 2] & 0xFFFFu));
    }
  } else {
    uint8_t* dst = (uint8_t*)C.pix + (long long)pic * C.pix_pic_stride + rowoff + gx;
#pragma unroll
    for (int k = 0; k < V; ++k) if (gx + k < C.pix_w) dst[k] = (uint8_t)v[k];
  }
}

// PPL consecutive samples of band row `by` starting at band column bx0 (multiple of PPL): vector access when
// the run is one aligned piece of the interleaved layout, else element by element
template <int PPL, bool STORE>
__device__ __forceinline__ void band_access(int32_t* coef, const BandAddr& ba, int base, int by, int bx0, int bxmax, int (&x)[PPL]) {
  if (bx0 < 0 || bx0 > bxmax) {
    if (!STORE) {
#pragma unroll
      for (int j = 0; j < PPL; ++j) x[j] = 0;
    }
    return;
  }
  const bool vec = PPL == 4 && (ba.bw & 3) == 0 && (base & 3) == 0 && bx0 + 3 <= bxmax;
  if (vec) {
    int4* p = reinterpret_cast<int4*>(coef + ba.at(base, by, bx0));
    if (STORE) *p = make_int4(x[0], x[1], x[PPL > 2 ? 2 : 0], x[PPL > 3 ? 3 : 0]);
    else { const int4 q = __ldg(p); x[0] = q.x; x[1] = q.y; x[PPL > 2 ? 2 : 0] = q.z; x[PPL > 3 ? 3 : 0] = q.w; }
  } else {
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
      if (bx0 + j <= bxmax) {
        int32_t* p = coef + ba.at(base, by, bx0 + j);
        if (STORE) *p = x[j]; else x[j] = *p;
      } else if (!STORE) x[j] = 0;
    }
  }
}
template <int PPL>
__device__ __forceinline__ void band_prefetch(const int32_t* coef, const BandAddr& ba, int base, int by, int bx0, int bxmax) {
  if (bx0 < 0 || bx0 > bxmax) return;
  prefetch_l1(coef + ba.at(base, by, bx0));
}
// same for the compact LL plane between levels
template <int PPL, bool STORE>
__device__ __forceinline__ void ll_access(int32_t* ll, long long pitch, int by, int bx0, int bxmax, int (&x)[PPL]) {
  if (bx0 < 0 || bx0 > bxmax) {
    if (!STORE) {
#pragma unroll
      for (int j = 0; j < PPL; ++j) x[j] = 0;
    }
    return;
  }
  int32_t* p = ll + (long long)by * pitch + bx0;
  if (PPL == 4 && bx0 + 3 <= bxmax && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    int4* q = reinterpret_cast<int4*>(p);
    if (STORE) *q = make_int4(x[0], x[1], x[PPL > 2 ? 2 : 0], x[PPL > 3 ? 3 : 0]);
    else { const int4 t = __ldg(q); x[0] = t.x; x[1] = t.y; x[PPL > 2 ? 2 : 0] = t.z; x[PPL > 3 ? 3 : 0] = t.w; }
  } else {
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
      if (bx0 + j <= bxmax) { if (STORE) p[j] = x[j]; else x[j] = p[j]; }
      else if (!STORE) x[j] = 0;
    }
  }
}

}  // namespace
}  // namespace vc2
