// Forward / inverse lifting level kernels and the in-place <-> planar layout kernels.
// See dwt.cuh for the reference line citations and the tiling scheme.
#include "dwt.cuh"

namespace vc2 {

namespace {

constexpr int TH = 64;    // tile rows (lattice samples)
constexpr int TW = 128;   // tile columns (lattice samples)
constexpr int BX = 32;    // blockDim.x : lanes run along a row (unit-stride shared memory)
constexpr int BY = 8;     // blockDim.y

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// One lifting step on one target sample.
//   tgt    : the sample being updated (parity Step::P)
//   oth    : other-parity sequence, element ii at oth[ii * stride]
//   idx    : pair index of the target; [lo, hi] = valid pair indices (reference tap clamping)
template <int K, int S, int DIR>
__device__ __forceinline__ void lift(int* tgt, const int* oth, int stride, int idx, int lo, int hi) {
  using ST = Step<K, S>;
  unsigned sum = (unsigned)ST::ADD;
#define VC2_TAPL(k) ((unsigned)oth[clampi(idx - (ST::P ? (k) : (k) + 1), lo, hi) * stride])
#define VC2_TAPR(k) ((unsigned)oth[clampi(idx + (ST::P ? (k) + 1 : (k)), lo, hi) * stride])
  if (ST::CL0 != 0) sum += (unsigned)ST::CL0 * VC2_TAPL(0);
  if (ST::CR0 != 0) sum += (unsigned)ST::CR0 * VC2_TAPR(0);
  if (ST::N > 1) {
    if (ST::CL1 != 0) sum += (unsigned)ST::CL1 * VC2_TAPL(1);
    if (ST::CR1 != 0) sum += (unsigned)ST::CR1 * VC2_TAPR(1);
  }
  if (ST::N > 2) {
    if (ST::CL2 != 0) sum += (unsigned)ST::CL2 * VC2_TAPL(2);
    if (ST::CR2 != 0) sum += (unsigned)ST::CR2 * VC2_TAPR(2);
    if (ST::CL3 != 0) sum += (unsigned)ST::CL3 * VC2_TAPL(3);
    if (ST::CR3 != 0) sum += (unsigned)ST::CR3 * VC2_TAPR(3);
  }
#undef VC2_TAPL
#undef VC2_TAPR
  const int delta = ((int)sum) >> ST::SH;
  if (ST::SIGN * DIR > 0) *tgt = (int)((unsigned)*tgt + (unsigned)delta);
  else *tgt = (int)((unsigned)*tgt - (unsigned)delta);
}

// Shared-memory tile: E = even columns, O = odd columns; RH rows, RP pairs per row.
template <int K>
struct Tile {
  static constexpr int R = Wavelet<K>::R;
  static constexpr int RH = TH + 2 * R;
  static constexpr int RP = (TW + 2 * R) / 2;
  static constexpr int BYTES = 2 * RH * RP * (int)sizeof(int);
};

// horizontal pass of one step over rows [r0, r1), pairs [jlo, jhi]
template <int K, int S, int DIR>
__device__ __forceinline__ void hpass(int* E, int* O, int r0, int r1, int jlo, int jhi) {
  constexpr int RP = Tile<K>::RP;
  for (int r = r0 + threadIdx.y; r < r1; r += BY) {
    int* tg = (Step<K, S>::P ? O : E) + r * RP;
    const int* ot = (Step<K, S>::P ? E : O) + r * RP;
    for (int j = jlo + threadIdx.x; j <= jhi; j += BX) lift<K, S, DIR>(tg + j, ot, 1, j, jlo, jhi);
  }
}

// vertical pass of one step over row pairs [ilo, ihi], pair columns [j0, j1) of both E and O
template <int K, int S, int DIR>
__device__ __forceinline__ void vpass(int* E, int* O, int ilo, int ihi, int j0, int j1) {
  constexpr int RP = Tile<K>::RP;
  constexpr int P = Step<K, S>::P;
  const int nj = j1 - j0;
  for (int i = ilo + threadIdx.y; i <= ihi; i += BY) {
    for (int c = threadIdx.x; c < 2 * nj; c += BX) {
      int* X = (c < nj) ? E : O;
      const int j = j0 + ((c < nj) ? c : c - nj);
      lift<K, S, DIR>(X + (2 * i + P) * RP + j, X + (1 - P) * RP + j, 2 * RP, i, ilo, ihi);
    }
  }
}

template <int K, int DIR, bool HORIZ>
__device__ __forceinline__ void all_steps(int* E, int* O, int a0, int a1, int b0, int b1) {
  // forward: steps 0..N-1 ; inverse: N-1..0 with the sign flipped
  constexpr int N = Wavelet<K>::NSTEPS;
#define VC2_RUN(S)                                                    \
  {                                                                   \
    if (HORIZ) hpass<K, S, DIR>(E, O, a0, a1, b0, b1);                \
    else vpass<K, S, DIR>(E, O, a0, a1, b0, b1);                      \
    __syncthreads();                                                  \
  }
  if (DIR > 0) {
    VC2_RUN(0) VC2_RUN(1)
    if (N == 4) { VC2_RUN((N == 4 ? 2 : 0)) VC2_RUN((N == 4 ? 3 : 1)) }
  } else {
    if (N == 4) { VC2_RUN((N == 4 ? 3 : 1)) VC2_RUN((N == 4 ? 2 : 0)) }
    VC2_RUN(1) VC2_RUN(0)
  }
#undef VC2_RUN
}

__device__ __forceinline__ int load_sample_u16be(const uint16_t* p, int sshift, int soffset) {
  const unsigned w = *p;
  const unsigned v = ((w >> 8) | (w << 8)) & 0xFFFFu;
  return (int)(v >> sshift) - soffset;
}

// ------------------------------------------------------------------------------------------
// forward level:  pix (dense plane)  ->  LL, HL, LH, HH
// ------------------------------------------------------------------------------------------
template <int K, int KIND>
__global__ void __launch_bounds__(BX* BY) dwt_fwd_kernel(const DwtParams p) {
  using T = Tile<K>;
  constexpr int R = T::R, RH = T::RH, RP = T::RP, SHIFT = Wavelet<K>::SHIFT;
  extern __shared__ int smem[];
  int* E = smem;
  int* O = smem + RH * RP;

  const int comp = blockIdx.z % p.ncomp;
  const int pic = blockIdx.z / p.ncomp;
  const DwtComp& C = p.c[comp];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  if (x0 >= C.lat_w || y0 >= C.lat_h) return;
  const int gx0 = x0 - R, gy0 = y0 - R;

  // valid local ranges (pairs / rows / row pairs) = region intersected with the lattice
  const int jlo = max(0, -gx0 / 2), jhi = min(RP - 1, (C.lat_w - 2 - gx0) / 2);
  const int rlo = max(0, -gy0), rhi = min(RH - 1, C.lat_h - 1 - gy0);   // rows (inclusive)
  const int ilo = rlo / 2, ihi = (rhi - 1) / 2;                         // row pairs (inclusive)

  // ---- load (edge replicate = waveletPad, WaveletTransform.cpp:79-94) and accuracy shift
  {
    const bool vec = ((C.pix_pitch & 1) == 0) && ((C.pix_pic_stride & 1) == 0);
    for (int r = rlo + threadIdx.y; r <= rhi; r += BY) {
      const int sy = min(gy0 + r, C.pix_h - 1);
      for (int j = jlo + threadIdx.x; j <= jhi; j += BX) {
        const int gx = gx0 + 2 * j;
        int e, o;
        if (KIND == SAMPLE_I32) {
          const int* src = (const int*)C.pix + (long long)pic * C.pix_pic_stride + (long long)sy * C.pix_pitch;
          if (vec && gx + 1 < C.pix_w) {
            const int2 v = *reinterpret_cast<const int2*>(src + gx);
            e = v.x; o = v.y;
          } else {
            e = src[min(gx, C.pix_w - 1)];
            o = src[min(gx + 1, C.pix_w - 1)];
          }
        } else if (KIND == SAMPLE_U16BE) {
          const uint16_t* src = (const uint16_t*)((const uint8_t*)C.pix + (long long)pic * C.pix_pic_stride) +
                                (long long)sy * C.pix_pitch;
          if (vec && gx + 1 < C.pix_w) {
            const unsigned w = *reinterpret_cast<const unsigned*>(src + gx);   // bytes: e_hi e_lo o_hi o_lo
            const unsigned ev = __byte_perm(w, 0, 0x4401), ov = __byte_perm(w, 0, 0x4423);
            e = (int)(ev >> C.sshift) - C.soffset;
            o = (int)(ov >> C.sshift) - C.soffset;
          } else {
            e = load_sample_u16be(src + min(gx, C.pix_w - 1), C.sshift, C.soffset);
            o = load_sample_u16be(src + min(gx + 1, C.pix_w - 1), C.sshift, C.soffset);
          }
        } else {
          const uint8_t* src = (const uint8_t*)C.pix + (long long)pic * C.pix_pic_stride + (long long)sy * C.pix_pitch;
          e = (int)((unsigned)src[min(gx, C.pix_w - 1)] >> C.sshift) - C.soffset;
          o = (int)((unsigned)src[min(gx + 1, C.pix_w - 1)] >> C.sshift) - C.soffset;
        }
        E[r * RP + j] = (int)((unsigned)e << SHIFT);
        O[r * RP + j] = (int)((unsigned)o << SHIFT);
      }
    }
  }
  __syncthreads();

  // ---- horizontal lifting on every region row, then vertical lifting on the tile's own columns
  all_steps<K, +1, true>(E, O, rlo, rhi + 1, jlo, jhi);
  const int tj0 = R / 2, tj1 = min(R / 2 + TW / 2, jhi + 1);
  all_steps<K, +1, false>(E, O, ilo, ihi, tj0, tj1);

  // ---- store the four subbands (dense rows)
  {
    int32_t* ll = C.ll + (long long)pic * C.ll_pic_stride;
    int32_t* hl = C.hl + (long long)pic * C.band_pic_stride;
    int32_t* lh = C.lh + (long long)pic * C.band_pic_stride;
    int32_t* hh = C.hh + (long long)pic * C.band_pic_stride;
    const int ny = min(TH, C.lat_h - y0), nj = min(TW, C.lat_w - x0) / 2;
    for (int yy = threadIdx.y; yy < ny; yy += BY) {
      const int r = R + yy, by = (y0 + yy) >> 1;
      const bool odd = yy & 1;
      int32_t* de = odd ? (lh + (long long)by * C.band_pitch) : (ll + (long long)by * C.ll_pitch);
      int32_t* dO = (odd ? hh : hl) + (long long)by * C.band_pitch;
      for (int jj = threadIdx.x; jj < nj; jj += BX) {
        const int bx = (x0 >> 1) + jj;
        de[bx] = E[r * RP + R / 2 + jj];
        dO[bx] = O[r * RP + R / 2 + jj];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// inverse level:  LL, HL, LH, HH  ->  pix (dense plane; cropped / clipped / packed at level 0)
// ------------------------------------------------------------------------------------------
template <int K, int KIND>
__global__ void __launch_bounds__(BX* BY) dwt_inv_kernel(const DwtParams p) {
  using T = Tile<K>;
  constexpr int R = T::R, RH = T::RH, RP = T::RP, SHIFT = Wavelet<K>::SHIFT;
  extern __shared__ int smem[];
  int* E = smem;
  int* O = smem + RH * RP;

  const int comp = blockIdx.z % p.ncomp;
  const int pic = blockIdx.z / p.ncomp;
  const DwtComp& C = p.c[comp];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  if (x0 >= C.lat_w || y0 >= C.lat_h) return;
  // nothing of this tile survives the crop (WaveletTransform.cpp:340)
  if (x0 >= C.pix_w || y0 >= C.pix_h) return;
  const int gx0 = x0 - R, gy0 = y0 - R;
  const int jlo = max(0, -gx0 / 2), jhi = min(RP - 1, (C.lat_w - 2 - gx0) / 2);
  const int rlo = max(0, -gy0), rhi = min(RH - 1, C.lat_h - 1 - gy0);
  const int ilo = rlo / 2, ihi = (rhi - 1) / 2;

  {
    const int32_t* ll = C.ll + (long long)pic * C.ll_pic_stride;
    const int32_t* hl = C.hl + (long long)pic * C.band_pic_stride;
    const int32_t* lh = C.lh + (long long)pic * C.band_pic_stride;
    const int32_t* hh = C.hh + (long long)pic * C.band_pic_stride;
    for (int r = rlo + threadIdx.y; r <= rhi; r += BY) {
      const int gy = gy0 + r, by = gy >> 1;
      const bool odd = gy & 1;
      const int32_t* se = odd ? (lh + (long long)by * C.band_pitch) : (ll + (long long)by * C.ll_pitch);
      const int32_t* so = (odd ? hh : hl) + (long long)by * C.band_pitch;
      for (int j = jlo + threadIdx.x; j <= jhi; j += BX) {
        const int bx = (gx0 >> 1) + j;   // gx0 is even; arithmetic shift is exact
        E[r * RP + j] = se[bx];
        O[r * RP + j] = so[bx];
      }
    }
  }
  __syncthreads();

  // vertical inverse on every region column, then horizontal inverse on the tile's own rows
  all_steps<K, -1, false>(E, O, ilo, ihi, jlo, jhi + 1);
  const int tr0 = R, tr1 = min(R + TH, rhi + 1);
  all_steps<K, -1, true>(E, O, tr0, tr1, jlo, jhi);

  {
    const int ny = min(TH, C.pix_h - y0);
    const int nx = min(TW, C.pix_w - x0);   // may be odd after the crop
    const int rnd = SHIFT ? (1 << (SHIFT - 1)) : 0;
    for (int yy = threadIdx.y; yy < ny; yy += BY) {
      const int r = R + yy;
      const long long row = (long long)(y0 + yy) * C.pix_pitch;
      for (int jj = threadIdx.x; 2 * jj < nx; jj += BX) {
        int e = E[r * RP + R / 2 + jj], o = O[r * RP + R / 2 + jj];
        if (SHIFT) {
          e = (e + rnd) >> SHIFT;
          o = (o + rnd) >> SHIFT;
        }
        const int x = x0 + 2 * jj;
        const bool has_o = (2 * jj + 1) < nx;
        if (KIND == SAMPLE_I32) {
          int* dst = (int*)C.pix + (long long)pic * C.pix_pic_stride + row;
          dst[x] = e;
          if (has_o) dst[x + 1] = o;
        } else {
          // clip (Picture.cpp:284-292), offset binary, MSB justify, big endian (Arrays.cpp:396-414)
          e = min(max(e, C.clip_min), C.clip_max);
          o = min(max(o, C.clip_min), C.clip_max);
          const unsigned ev = (unsigned)(e + C.soffset) << C.sshift;
          const unsigned ov = (unsigned)(o + C.soffset) << C.sshift;
          if (KIND == SAMPLE_U16BE) {
            uint16_t* dst = (uint16_t*)((uint8_t*)C.pix + (long long)pic * C.pix_pic_stride) + row;
            dst[x] = (uint16_t)(((ev >> 8) & 0xFF) | ((ev & 0xFF) << 8));
            if (has_o) dst[x + 1] = (uint16_t)(((ov >> 8) & 0xFF) | ((ov & 0xFF) << 8));
          } else {
            uint8_t* dst = (uint8_t*)C.pix + (long long)pic * C.pix_pic_stride + row;
            dst[x] = (uint8_t)ev;
            if (has_o) dst[x + 1] = (uint8_t)ov;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// layout kernels: reference in-place interleaved order <-> planar subbands
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long planar_index(const PlaneGeom& g, int y, int x) {
  const int d = g.depth;
  const int t = (y | x) & ((1 << d) - 1);
  if (t == 0) return (long long)(y >> d) * (g.pw >> d) + (x >> d);
  const int l = __ffs(t) - 1;          // 0 = finest level
  const int L = d - l;                 // VC-2 level number
  const int hx = (x >> l) & 1, hy = (y >> l) & 1;
  const int type = hx ? (hy ? 3 : 1) : 2;
  const long long n0 = (long long)(g.ph >> d) * (g.pw >> d);
  const long long off = (n0 << (2 * (L - 1))) * type;
  return off + (long long)(y >> (l + 1)) * (g.pw >> (l + 1)) + (x >> (l + 1));
}

__global__ void inplace_to_planar_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, PlaneGeom g,
                                         long long src_pic_stride, long long dst_pic_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= g.pw || y >= g.ph) return;
  dst[blockIdx.z * dst_pic_stride + planar_index(g, y, x)] = src[blockIdx.z * src_pic_stride + (long long)y * g.pw + x];
}

__global__ void planar_to_inplace_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, PlaneGeom g,
                                         long long src_pic_stride, long long dst_pic_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= g.pw || y >= g.ph) return;
  dst[blockIdx.z * dst_pic_stride + (long long)y * g.pw + x] = src[blockIdx.z * src_pic_stride + planar_index(g, y, x)];
}

template <int K, int KIND>
cudaError_t launch_fwd(cudaStream_t s, const DwtParams& p, dim3 grid) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(dwt_fwd_kernel<K, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<K>::BYTES);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dwt_fwd_kernel<K, KIND><<<grid, dim3(BX, BY), Tile<K>::BYTES, s>>>(p);
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
  dwt_inv_kernel<K, KIND><<<grid, dim3(BX, BY), Tile<K>::BYTES, s>>>(p);
  return cudaGetLastError();
}

template <int KIND>
cudaError_t dispatch(cudaStream_t s, bool inverse, int kernel, const DwtParams& p, dim3 grid) {
#define VC2_CASE(K) case K: return inverse ? launch_inv<K, KIND>(s, p, grid) : launch_fwd<K, KIND>(s, p, grid);
  switch (kernel) {
    VC2_CASE(VC2_DD97) VC2_CASE(VC2_LEGALL) VC2_CASE(VC2_DD137) VC2_CASE(VC2_HAAR0)
    VC2_CASE(VC2_HAAR1) VC2_CASE(VC2_FIDELITY) VC2_CASE(VC2_DAUB97)
    default: return cudaErrorInvalidValue;
  }
#undef VC2_CASE
}

}  // namespace

// grid covering the largest component lattice, z = pictures * components
static dim3 level_grid(const DwtParams& p, int npictures) {
  int mw = 0, mh = 0;
  for (int c = 0; c < p.ncomp; ++c) {
    mw = p.c[c].lat_w > mw ? p.c[c].lat_w : mw;
    mh = p.c[c].lat_h > mh ? p.c[c].lat_h : mh;
  }
  return dim3((mw + TW - 1) / TW, (mh + TH - 1) / TH, npictures * p.ncomp);
}

cudaError_t dwt_level_launch(cudaStream_t s, bool inverse, int kernel, int sample_kind, const DwtParams& p, int npictures) {
  const dim3 grid = level_grid(p, npictures);
  switch (sample_kind) {
    case SAMPLE_I32: return dispatch<SAMPLE_I32>(s, inverse, kernel, p, grid);
    case SAMPLE_U16BE: return dispatch<SAMPLE_U16BE>(s, inverse, kernel, p, grid);
    case SAMPLE_U8: return dispatch<SAMPLE_U8>(s, inverse, kernel, p, grid);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t layout_launch(cudaStream_t s, bool to_planar, const int32_t* src, int32_t* dst, const PlaneGeom& g,
                          long long src_pic_stride, long long dst_pic_stride, int npictures) {
  const dim3 block(32, 8), grid((g.pw + 31) / 32, (g.ph + 7) / 8, npictures);
  if (to_planar) inplace_to_planar_kernel<<<grid, block, 0, s>>>(src, dst, g, src_pic_stride, dst_pic_stride);
  else planar_to_inplace_kernel<<<grid, block, 0, s>>>(src, dst, g, src_pic_stride, dst_pic_stride);
  return cudaGetLastError();
}

}  // namespace vc2
