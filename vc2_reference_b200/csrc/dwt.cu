// Forward / inverse lifting level kernels and the in-place <-> slice-major layout kernels.
// See dwt.cuh for the reference line citations and the tiling scheme.
#include "dwt.cuh"
#include "slices.cuh"

namespace vc2 {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int TH = 64;      // useful tile rows (lattice samples)
constexpr int RW = 128;     // region columns: one warp row segment, 4 samples per lane
constexpr int RP = RW / 2;  // pairs per region row = shared-memory row pitch of E and O
constexpr int NT = 256;     // threads per CTA (8 warps)

template <int K>
struct Tile {
  static constexpr int R = Wavelet<K>::R, HX = Wavelet<K>::HX;
  static constexpr int RH = TH + 2 * R;
  static constexpr int TWU = RW - 2 * HX;   // useful columns per tile
  static constexpr int BYTES = 2 * RH * RP * (int)sizeof(int);
};

// reach of one lifting step along the lifted axis, and the reach still to come after / before it
template <int K, int S> __host__ __device__ constexpr int step_reach() { return Wavelet<K>::R == 0 ? 0 : 2 * Step<K, S>::N - 1; }
template <int K, int S> __host__ __device__ constexpr int reach_after() {   // forward order: steps S+1 .. NSTEPS-1
  if constexpr (S + 1 >= Wavelet<K>::NSTEPS) return 0;
  else return step_reach<K, S + 1>() + reach_after<K, S + 1>();
}
template <int K, int S> __host__ __device__ constexpr int reach_before() {  // inverse order: steps S-1 .. 0
  if constexpr (S == 0) return 0;
  else return step_reach<K, S - 1>() + reach_before<K, S - 1>();
}

// ---- horizontal lifting step in registers ------------------------------------------------------
// A lane holds pairs 2*lane and 2*lane+1 of a 64-pair row segment: e[a] / o[a] = even / odd sample
// of pair 2*lane+a.  Neighbouring pairs come from other lanes by shuffle.  When the segment
// touches the left/right edge of the lattice (hedge), the source-parity sequence is first extended
// beyond [plo, phi] with its edge value = the reference's tap clamping.
template <int K, int S, int DIR>
__device__ __forceinline__ void hstep(int (&e)[2], int (&o)[2], int lane, bool hedge, int plo, int phi) {
  using ST = Step<K, S>;
  constexpr int P = ST::P, N = ST::N;
  int (&src)[2] = P ? e : o;
  int (&tgt)[2] = P ? o : e;
  if (hedge) {
    const int vlo = __shfl_sync(FULL, (plo & 1) ? src[1] : src[0], plo >> 1);
    const int vhi = __shfl_sync(FULL, (phi & 1) ? src[1] : src[0], phi >> 1);
    const int p0 = 2 * lane;
    if (p0 < plo) src[0] = vlo; else if (p0 > phi) src[0] = vhi;
    if (p0 + 1 < plo) src[1] = vlo; else if (p0 + 1 > phi) src[1] = vhi;
  }
  constexpr int QMIN = -(P ? N - 1 : N), QMAX = 1 + (P ? N : N - 1);
  int nb[QMAX - QMIN + 1];   // source values at pair offsets QMIN..QMAX from pair 2*lane
#pragma unroll
  for (int q = QMIN; q <= QMAX; ++q) {
    const int r = q & 1, dl = (q - r) / 2;
    nb[q - QMIN] = dl == 0 ? src[r] : __shfl_sync(FULL, src[r], lane + dl);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    unsigned sum = (unsigned)ST::ADD;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (ST::cl(k) != 0) sum += (unsigned)ST::cl(k) * (unsigned)nb[a - (P ? k : k + 1) - QMIN];
      if (ST::cr(k) != 0) sum += (unsigned)ST::cr(k) * (unsigned)nb[a + (P ? k + 1 : k) - QMIN];
    }
    const int delta = ((int)sum) >> ST::SH;
    tgt[a] = (ST::SIGN * DIR > 0) ? (int)((unsigned)tgt[a] + (unsigned)delta) : (int)((unsigned)tgt[a] - (unsigned)delta);
  }
}

template <int K, int DIR>
__device__ __forceinline__ void hsteps(int (&e)[2], int (&o)[2], int lane, bool hedge, int plo, int phi) {
  constexpr int N = Wavelet<K>::NSTEPS;
  if constexpr (DIR > 0) {
    hstep<K, 0, DIR>(e, o, lane, hedge, plo, phi);
    hstep<K, 1, DIR>(e, o, lane, hedge, plo, phi);
    if constexpr (N == 4) {
      hstep<K, 2, DIR>(e, o, lane, hedge, plo, phi);
      hstep<K, 3, DIR>(e, o, lane, hedge, plo, phi);
    }
  } else {
    if constexpr (N == 4) {
      hstep<K, 3, DIR>(e, o, lane, hedge, plo, phi);
      hstep<K, 2, DIR>(e, o, lane, hedge, plo, phi);
    }
    hstep<K, 1, DIR>(e, o, lane, hedge, plo, phi);
    hstep<K, 0, DIR>(e, o, lane, hedge, plo, phi);
  }
}

// ---- vertical lifting step on the shared-memory tile ---------------------------------------------
// x points at the target sample X[r][j]; taps sit at fixed row offsets (row pitch RP)
template <int K, int S, int DIR>
__device__ __forceinline__ int vlift(const int* x) {
  using ST = Step<K, S>;
  unsigned sum = (unsigned)ST::ADD;
#pragma unroll
  for (int k = 0; k < ST::N; ++k) {
    if (ST::cl(k) != 0) sum += (unsigned)ST::cl(k) * (unsigned)x[-(2 * k + 1) * RP];
    if (ST::cr(k) != 0) sum += (unsigned)ST::cr(k) * (unsigned)x[(2 * k + 1) * RP];
  }
  const int delta = ((int)sum) >> ST::SH;
  return (ST::SIGN * DIR > 0) ? (int)((unsigned)x[0] + (unsigned)delta) : (int)((unsigned)x[0] - (unsigned)delta);
}

// copy the first / last valid row of parity q into the out-of-lattice halo rows of that parity
// (rows [0, rlo) and (rhi, RH)); only CTAs at the top / bottom edge of the lattice have any
__device__ __forceinline__ void extend_rows(int* E, int* O, int q, int rlo, int rhi, int RH) {
  const int col = threadIdx.x & 127, half = threadIdx.x >> 7;
  int* X = (col < RP ? E : O) + (col & (RP - 1));
  const int first = rlo + q, last = rhi - (1 - q);
  for (int r = q + 2 * half; r < rlo; r += 4) X[r * RP] = X[first * RP];
  for (int r = last + 2 + 2 * half; r < RH; r += 4) X[r * RP] = X[last * RP];
}

struct BandAddr {   // group-interleaved addressing of one band sample (see vc2_common.cuh)
  int bh, bw, lgbh, lgbw, nx, nc4;
  int sx, xin;       // slice column and column inside the slice's part of the band: fixed per thread
  __device__ __forceinline__ void set_col(int bx) {
    sx = lgbw >= 0 ? (bx >> lgbw) : (bx / bw);
    xin = bx - sx * bw;
  }
  __device__ __forceinline__ long long at(int base, int by) const {
    const int sy = lgbh >= 0 ? (by >> lgbh) : (by / bh);
    return coef_index(sy * nx + sx, base + (by - sy * bh) * bw + xin, nc4);
  }
};

__device__ __forceinline__ int sample_u16be(unsigned w, int sshift, int soffset) {
  return (int)((((w >> 8) | (w << 8)) & 0xFFFFu) >> sshift) - soffset;
}

// ------------------------------------------------------------------------------------------
// forward level:  pix (dense plane)  ->  LL (compact plane or band 0), HL, LH, HH (slice-major)
// ------------------------------------------------------------------------------------------
template <int K, int S>
__device__ __forceinline__ void fwd_vstep(int* E, int* O, const DwtComp& C, int pic, int ys, int rlo, int rhi, int xs, int phi,
                                          bool vedge) {
  using T = Tile<K>;
  using ST = Step<K, S>;
  constexpr int R = T::R, RH = T::RH, P = ST::P, NS = Wavelet<K>::NSTEPS;
  constexpr bool LAST = (S == NS - 1);
  constexpr bool FINAL_FOR_PARITY = (S >= NS - 2);   // steps alternate parity: the last two finish one parity each
  if (vedge) {
    extend_rows(E, O, 1 - P, rlo, rhi, RH);
    __syncthreads();
  }
  constexpr int RA = reach_after<K, S>();
  const int tlo = max(R - RA, rlo), thi = min(R + TH - 1 + RA, rhi);
  const int col = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int j = col & (RP - 1);
  const bool isO = col >= RP;
  if (j >= T::HX / 2 && j < T::HX / 2 + T::TWU / 2 && j <= phi) {
    int* X = (isO ? O : E) + j;
    const int bx = (xs >> 1) + j;
    BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2, 0, 0};
    ba.set_col(bx);
    int32_t* coef = C.coef + (long long)pic * C.coef_pic_stride;
    int t0 = tlo + ((tlo & 1) != P ? 1 : 0);
    for (int r = t0 + 2 * half; r <= thi; r += 4) {
      const int v = vlift<K, S, +1>(X + r * RP);
      if (!LAST) X[r * RP] = v;
      if (FINAL_FOR_PARITY && r >= R && r < R + TH) {
        const int by = (ys + r) >> 1;
        if (P == 0 && !isO && C.ll) {
          C.ll[(long long)pic * C.ll_pic_stride + (long long)by * C.ll_pitch + bx] = v;
        } else {
          const int base = P ? (isO ? C.base_hh : C.base_lh) : (isO ? C.base_hl : C.base_ll);
          coef[ba.at(base, by)] = v;
        }
      }
    }
  }
  if (!LAST) __syncthreads();
}

template <int K, int KIND>
__global__ void __launch_bounds__(NT) dwt_fwd_kernel(const DwtParams p) {
  using T = Tile<K>;
  constexpr int R = T::R, RH = T::RH, HX = T::HX, SHIFT = Wavelet<K>::SHIFT, NS = Wavelet<K>::NSTEPS;
  extern __shared__ int smem[];
  int* E = smem;
  int* O = smem + RH * RP;

  const int comp = blockIdx.z % p.ncomp, pic = blockIdx.z / p.ncomp;
  const DwtComp& C = p.c[comp];
  const int x0 = blockIdx.x * T::TWU, y0 = blockIdx.y * TH;
  if (x0 >= C.lat_w || y0 >= C.lat_h) return;
  const int xs = x0 - HX, ys = y0 - R;
  const int plo = max(0, -xs / 2), phi = min(RP - 1, (C.lat_w - 2 - xs) / 2);
  const bool hedge = xs < 0 || xs + RW > C.lat_w;
  const int rlo = max(0, -ys), rhi = min(RH - 1, C.lat_h - 1 - ys);
  const bool vedge = ys < 0 || ys + RH > C.lat_h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- phase 1: load a row segment, accuracy shift, horizontal lifting in registers, park in the tile
  {
    const int gx = xs + 4 * lane;
    const bool inside = gx >= 0 && gx + 3 < C.pix_w;   // all four samples are real picture samples
    for (int r = rlo + warp; r <= rhi; r += NT / 32) {
      const int sy = min(ys + r, C.pix_h - 1);          // waveletPad: replicate the last row (WaveletTransform.cpp:88)
      int v[4];
      if (KIND == SAMPLE_I32) {
        const int* row = (const int*)C.pix + (long long)pic * C.pix_pic_stride + (long long)sy * C.pix_pitch;
        if (inside && ((reinterpret_cast<uintptr_t>(row + gx) & 15) == 0)) {
          const int4 q = *reinterpret_cast<const int4*>(row + gx);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = row[min(max(gx + k, 0), C.pix_w - 1)];   // :89 replicate the last column
        }
      } else if (KIND == SAMPLE_U16BE) {
        const uint16_t* row = (const uint16_t*)((const uint8_t*)C.pix + (long long)pic * C.pix_pic_stride) + (long long)sy * C.pix_pitch;
        if (inside && ((reinterpret_cast<uintptr_t>(row + gx) & 7) == 0)) {
          const uint2 q = *reinterpret_cast<const uint2*>(row + gx);
          v[0] = sample_u16be(q.x & 0xFFFFu, C.sshift, C.soffset);
          v[1] = sample_u16be(q.x >> 16, C.sshift, C.soffset);
          v[2] = sample_u16be(q.y & 0xFFFFu, C.sshift, C.soffset);
          v[3] = sample_u16be(q.y >> 16, C.sshift, C.soffset);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = sample_u16be(row[min(max(gx + k, 0), C.pix_w - 1)], C.sshift, C.soffset);
        }
      } else {
        const uint8_t* row = (const uint8_t*)C.pix + (long long)pic * C.pix_pic_stride + (long long)sy * C.pix_pitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (int)((unsigned)row[min(max(gx + k, 0), C.pix_w - 1)] >> C.sshift) - C.soffset;
      }
      int e[2] = {(int)((unsigned)v[0] << SHIFT), (int)((unsigned)v[2] << SHIFT)};
      int o[2] = {(int)((unsigned)v[1] << SHIFT), (int)((unsigned)v[3] << SHIFT)};
      hsteps<K, +1>(e, o, lane, hedge, plo, phi);
      *reinterpret_cast<int2*>(E + r * RP + 2 * lane) = make_int2(e[0], e[1]);
      *reinterpret_cast<int2*>(O + r * RP + 2 * lane) = make_int2(o[0], o[1]);
    }
  }
  __syncthreads();

  // ---- phase 2: vertical lifting on the tile, results straight to global memory
  fwd_vstep<K, 0>(E, O, C, pic, ys, rlo, rhi, xs, phi, vedge);
  fwd_vstep<K, 1>(E, O, C, pic, ys, rlo, rhi, xs, phi, vedge);
  if constexpr (NS == 4) {
    fwd_vstep<K, 2>(E, O, C, pic, ys, rlo, rhi, xs, phi, vedge);
    fwd_vstep<K, 3>(E, O, C, pic, ys, rlo, rhi, xs, phi, vedge);
  }
}

// ------------------------------------------------------------------------------------------
// inverse level:  LL, HL, LH, HH  ->  pix (dense plane; cropped / clipped / packed at level 0)
// ------------------------------------------------------------------------------------------
template <int K, int S>
__device__ __forceinline__ void inv_vstep(int* E, int* O, int rlo, int rhi, int plo, int phi, bool vedge) {
  using T = Tile<K>;
  using ST = Step<K, S>;
  constexpr int R = T::R, RH = T::RH, P = ST::P;
  if (vedge) {
    extend_rows(E, O, 1 - P, rlo, rhi, RH);
    __syncthreads();
  }
  constexpr int RA = reach_before<K, S>();
  const int tlo = max(R - RA, rlo), thi = min(R + TH - 1 + RA, rhi);
  const int col = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int j = col & (RP - 1);
  if (j >= plo && j <= phi) {
    int* X = (col >= RP ? O : E) + j;
    const int t0 = tlo + ((tlo & 1) != P ? 1 : 0);
    for (int r = t0 + 2 * half; r <= thi; r += 4) X[r * RP] = vlift<K, S, -1>(X + r * RP);
  }
  __syncthreads();
}

template <int K, int KIND>
__global__ void __launch_bounds__(NT) dwt_inv_kernel(const DwtParams p) {
  using T = Tile<K>;
  constexpr int R = T::R, RH = T::RH, HX = T::HX, SHIFT = Wavelet<K>::SHIFT, NS = Wavelet<K>::NSTEPS;
  extern __shared__ int smem[];
  int* E = smem;
  int* O = smem + RH * RP;

  const int comp = blockIdx.z % p.ncomp, pic = blockIdx.z / p.ncomp;
  const DwtComp& C = p.c[comp];
  const int x0 = blockIdx.x * T::TWU, y0 = blockIdx.y * TH;
  if (x0 >= C.lat_w || y0 >= C.lat_h) return;
  if (x0 >= C.pix_w || y0 >= C.pix_h) return;   // nothing of this tile survives the crop (WaveletTransform.cpp:340)
  const int xs = x0 - HX, ys = y0 - R;
  const int plo = max(0, -xs / 2), phi = min(RP - 1, (C.lat_w - 2 - xs) / 2);
  const bool hedge = xs < 0 || xs + RW > C.lat_w;
  const int rlo = max(0, -ys), rhi = min(RH - 1, C.lat_h - 1 - ys);
  const bool vedge = ys < 0 || ys + RH > C.lat_h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- phase 1: gather the four subbands into the tile
  {
    const int col = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int j = col & (RP - 1);
    const bool isO = col >= RP;
    if (j >= plo && j <= phi) {
      int* X = (isO ? O : E) + j;
      const int bx = (xs >> 1) + j;
      BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2, 0, 0};
      ba.set_col(bx);
      const int32_t* coef = C.coef + (long long)pic * C.coef_pic_stride;
      for (int r = rlo + half; r <= rhi; r += 2) {
        const int gy = ys + r, by = gy >> 1;
        int v;
        if (!(gy & 1) && !isO && C.ll) v = C.ll[(long long)pic * C.ll_pic_stride + (long long)by * C.ll_pitch + bx];
        else {
          const int base = (gy & 1) ? (isO ? C.base_hh : C.base_lh) : (isO ? C.base_hl : C.base_ll);
          v = coef[ba.at(base, by)];
        }
        X[r * RP] = v;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: vertical inverse lifting (steps in reverse order, sign flipped)
  if constexpr (NS == 4) {
    inv_vstep<K, 3>(E, O, rlo, rhi, plo, phi, vedge);
    inv_vstep<K, 2>(E, O, rlo, rhi, plo, phi, vedge);
  }
  inv_vstep<K, 1>(E, O, rlo, rhi, plo, phi, vedge);
  inv_vstep<K, 0>(E, O, rlo, rhi, plo, phi, vedge);

  // ---- phase 3: horizontal inverse lifting in registers, rounding, crop, clip, pack, store
  {
    const int gx = xs + 4 * lane;
    const bool mine = gx >= x0 && gx < x0 + T::TWU && gx < C.pix_w;   // lane groups are wholly useful or wholly halo
    const int rnd = SHIFT ? (1 << (SHIFT - 1)) : 0;
    const int rend = min(R + TH - 1, min(rhi, C.pix_h - 1 - ys));
    for (int r = R + warp; r <= rend; r += NT / 32) {
      const int2 ev = *reinterpret_cast<const int2*>(E + r * RP + 2 * lane);
      const int2 ov = *reinterpret_cast<const int2*>(O + r * RP + 2 * lane);
      int e[2] = {ev.x, ev.y}, o[2] = {ov.x, ov.y};
      hsteps<K, -1>(e, o, lane, hedge, plo, phi);
      if (!mine) continue;
      int v[4] = {e[0], o[0], e[1], o[1]};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (SHIFT) v[k] = (v[k] + rnd) >> SHIFT;
        if (KIND != SAMPLE_I32) v[k] = (int)((unsigned)(min(max(v[k], C.clip_min), C.clip_max) + C.soffset) << C.sshift);
      }
      const long long rowoff = (long long)(ys + r) * C.pix_pitch;
      const bool whole = gx + 3 < C.pix_w;
      if (KIND == SAMPLE_I32) {
        int* dst = (int*)C.pix + (long long)pic * C.pix_pic_stride + rowoff + gx;
        if (whole && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<int4*>(dst) = make_int4(v[0], v[1], v[2], v[3]);
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k) if (gx + k < C.pix_w) dst[k] = v[k];
        }
      } else if (KIND == SAMPLE_U16BE) {
        // offset binary, MSB justified, big endian (Arrays.cpp:396-414)
        uint16_t* dst = (uint16_t*)((uint8_t*)C.pix + (long long)pic * C.pix_pic_stride) + rowoff + gx;
        unsigned h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __byte_perm((unsigned)v[k], 0, 0x4401);
        if (whole && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) *reinterpret_cast<uint2*>(dst) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k) if (gx + k < C.pix_w) dst[k] = (uint16_t)h[k];
        }
      } else {
        uint8_t* dst = (uint8_t*)C.pix + (long long)pic * C.pix_pic_stride + rowoff + gx;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (gx + k < C.pix_w) dst[k] = (uint8_t)v[k];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// layout kernels: reference in-place interleaved plane <-> slice-major coefficient block
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long slice_major_index(const SliceGeom& g, int c, int y, int x) {
  const int d = g.depth;
  const int t = (y | x) & ((1 << d) - 1);
  int b, by, bx;
  if (t == 0) { b = 0; by = y >> d; bx = x >> d; }
  else {
    const int l = __ffs(t) - 1;          // 0 = finest level
    const int L = d - l;                 // VC-2 level number
    const int hx = (x >> l) & 1, hy = (y >> l) & 1;
    b = 3 * (L - 1) + (hx ? (hy ? 3 : 1) : 2);
    by = y >> (l + 1); bx = x >> (l + 1);
  }
  const int bh = g.part_h[c][b], bw = g.part_w[c][b];
  const int sy = by / bh, sx = bx / bw;
  return coef_index(sy * g.slices_x + sx, g.comp_start[c] + g.band_start[c][b] + (by - sy * bh) * bw + (bx - sx * bw), g.comp_start[3] >> 2);
}

__global__ void layout_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, const SliceGeom g, int c, bool to_slice_major) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const PlaneGeom& pg = g.plane[c];
  if (x >= pg.pw || y >= pg.ph) return;
  const long long a = (long long)y * pg.pw + x, b = slice_major_index(g, c, y, x);
  if (to_slice_major) dst[b] = src[a];
  else dst[a] = src[b];
}

template <int K, int KIND>
cudaError_t launch_fwd(cudaStream_t s, const DwtParams& p, dim3 grid) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(dwt_fwd_kernel<K, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<K>::BYTES);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dwt_fwd_kernel<K, KIND><<<grid, NT, Tile<K>::BYTES, s>>>(p);
  return cudaGetLastError();
}
template <int K, int KIND>
cudaError_t launch_inv(cudaStream_t s, const DwtParams& p, dim3 grid) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(dwt_inv_kernel<K, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<K>::BYTES);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dwt_inv_kernel<K, KIND><<<grid, NT, Tile<K>::BYTES, s>>>(p);
  return cudaGetLastError();
}

template <int K, int KIND>
cudaError_t launch_level(cudaStream_t s, bool inverse, const DwtParams& p, int npictures) {
  int mw = 0, mh = 0;
  for (int c = 0; c < p.ncomp; ++c) {
    mw = p.c[c].lat_w > mw ? p.c[c].lat_w : mw;
    mh = p.c[c].lat_h > mh ? p.c[c].lat_h : mh;
  }
  const dim3 grid((mw + Tile<K>::TWU - 1) / Tile<K>::TWU, (mh + TH - 1) / TH, npictures * p.ncomp);
  return inverse ? launch_inv<K, KIND>(s, p, grid) : launch_fwd<K, KIND>(s, p, grid);
}

template <int KIND>
cudaError_t dispatch(cudaStream_t s, bool inverse, int kernel, const DwtParams& p, int npictures) {
#define VC2_CASE(K) case K: return launch_level<K, KIND>(s, inverse, p, npictures);
  switch (kernel) {
    VC2_CASE(VC2_DD97) VC2_CASE(VC2_LEGALL) VC2_CASE(VC2_DD137) VC2_CASE(VC2_HAAR0)
    VC2_CASE(VC2_HAAR1) VC2_CASE(VC2_FIDELITY) VC2_CASE(VC2_DAUB97)
    default: return cudaErrorInvalidValue;
  }
#undef VC2_CASE
}

}  // namespace

cudaError_t dwt_level_launch(cudaStream_t s, bool inverse, int kernel, int sample_kind, const DwtParams& p, int npictures) {
  switch (sample_kind) {
    case SAMPLE_I32: return dispatch<SAMPLE_I32>(s, inverse, kernel, p, npictures);
    case SAMPLE_U16BE: return dispatch<SAMPLE_U16BE>(s, inverse, kernel, p, npictures);
    case SAMPLE_U8: return dispatch<SAMPLE_U8>(s, inverse, kernel, p, npictures);
    default: return cudaErrorInvalidValue;
  }
}

// src/dst: one in-place plane (ph x pw) and one picture's slice-major block
cudaError_t layout_launch(cudaStream_t s, bool to_slice_major, const int32_t* src, int32_t* dst, const SliceGeom& g, int c) {
  const PlaneGeom& pg = g.plane[c];
  const dim3 block(32, 8), grid((pg.pw + 31) / 32, (pg.ph + 7) / 8);
  layout_kernel<<<grid, block, 0, s>>>(src, dst, g, c, to_slice_major);
  return cudaGetLastError();
}

}  // namespace vc2
