// Shared-memory TILE kernels for one lifting level (forward and inverse), sm_100a.  See dwt.cuh for the reference
// line citations (WaveletTransform.cpp:262-342, 478-1265).
//
// One CTA owns a tile of TH x TW lattice samples (halo included) in shared memory and runs three phases over it:
//
//   forward   P1  rows:    a warp (or half warp) takes one lattice row, eight samples per lane straight from the picture
//                          (raw big-endian words are converted on the way in), lifts it horizontally in registers -
//                          neighbours by warp shuffle - and stores the even and the odd columns as separate halves of the
//                          tile row
//             P2  columns: the SAME 1-D lifting routine, with the lanes now spread over the ROWS: a lane takes eight
//                          consecutive rows of four adjacent columns (eight 16-byte shared-memory loads), the
//                          neighbouring rows come from the neighbouring lanes by shuffle, the result goes back in place
//             P3  rows:    every useful row leaves as whole 16-byte pieces of the group-interleaved coefficient block
//                          (HL / LH / HH, and LL at the last level) or of the compact LL plane, one band row per warp:
//                          the requests of a warp are whole 128-byte runs
//   inverse   P1 fetches band rows into the tile, P2 is the vertical inverse lifting, P3 lifts the rows horizontally,
//             rounds, clips, converts and stores picture rows.
//
// The tile is the transposer between "lanes across columns" (all global memory traffic, coalesced) and "lanes across
// rows" (vertical lifting without a register ring: nothing rotates, nothing is warmed up, and the reference's edge rule
// is the same source-sequence extension in both directions).  16-byte chunks of a tile row are XOR-swizzled with bits
// 3..5 of the row number so that the eight lanes of a shared-memory phase, which sit eight rows apart in P2, hit eight
// different bank groups.
#include "dwt_lift.cuh"

#ifndef VC2_DWT_PART
#error "compile with -DVC2_DWT_PART=1 (forward) or =2 (inverse)"
#endif

namespace vc2 {

namespace {

// TWL lanes per tile row (8 samples each), TVL lanes per tile column group (8 rows each), NW warps per CTA
template <int K, int TWL_, int TVL_, int NW_, int MINB_>
struct Tile {
  static constexpr int TWL = TWL_, TVL = TVL_, NW = NW_, MINB = MINB_;   // MINB: resident CTAs per SM the registers are held to
  static constexpr int HL = (Wavelet<K>::R + 7) / 8;      // halo, in lanes, on each side (both directions)
  static constexpr int TW = 8 * TWL, TH = 8 * TVL;
  static constexpr int XU = TW - 16 * HL, YU = TH - 16 * HL;   // useful samples
  static constexpr int CHUNKS = 2 * TWL;                  // 16-byte chunks per tile row: TWL of even columns, then TWL of odd ones
  static constexpr int RPW = 32 / TWL;                    // rows per warp and iteration in the row phases
  static constexpr int TPW = 32 / TVL;                    // column tasks per warp and iteration in the column phase
  static constexpr int SMEM = TH * TW * 4;
  static_assert(TH % (NW * RPW) == 0 && YU % (NW * RPW) == 0, "row phases run without a tail");
};

struct TileList {     // flat list of the tiles of one picture: component c owns tiles [start[c], start[c + 1])
  int tx[3], start[4];
};

__device__ __forceinline__ int swz(int row) { return (row >> 3) & 7; }

// per-lane constants of the vectorised band access (band part widths are multiples of four, power-of-two part heights)
struct BandFast {
  int32_t* coefpic;
  int sx, kx4, lgbh, bhm1, bw4, nx, nc4;
  int o_ll, o_hl, o_lh, o_hh;
  int32_t* llp;
  int ll_pitch;
  __device__ __forceinline__ int idx(int by) const {
    const int sy = by >> lgbh, ry = by & bhm1;
    const int s = sy * nx + sx;
    return ((((s >> 5) * nc4 + ry * bw4 + kx4) << 5) + (s & 31)) << 2;
  }
};
// false when this level / tile has to use the element-wise access
__device__ __forceinline__ bool band_fast_setup(const DwtComp& C, int pic, int bx0, bool inside, BandFast& F) {
  if (!inside) return false;
  if (C.lgbh < 0 || C.lgbw < 0 || (C.bw & 3) || ((C.base_ll | C.base_hl | C.base_lh | C.base_hh) & 3)) return false;
  F.coefpic = C.coef + (long long)pic * C.coef_pic_stride;
  F.sx = bx0 >> C.lgbw;
  F.kx4 = (bx0 & (C.bw - 1)) >> 2;
  F.lgbh = C.lgbh; F.bhm1 = C.bh - 1; F.bw4 = C.bw >> 2; F.nx = C.nx; F.nc4 = C.NC >> 2;
  F.o_ll = (C.base_ll >> 2) * 128; F.o_hl = (C.base_hl >> 2) * 128; F.o_lh = (C.base_lh >> 2) * 128; F.o_hh = (C.base_hh >> 2) * 128;
  F.llp = nullptr; F.ll_pitch = C.ll_pitch;
  if (C.ll) {
    F.llp = C.ll + (long long)pic * C.ll_pic_stride + bx0;
    if ((C.ll_pitch & 3) || (reinterpret_cast<uintptr_t>(C.ll + (long long)pic * C.ll_pic_stride) & 15)) return false;
  }
  if (reinterpret_cast<uintptr_t>(F.coefpic) & 15) return false;
  return true;
}

struct TileCtx {
  int pic, comp;
  int x0, y0, xs, ys;          // first useful sample; first sample of the tile (halo included)
  int plo, phi, vplo, vphi;    // valid pair range of a tile row / of a tile column, in pairs from the tile origin
  bool hedge, vedge;
};

template <class T>
__device__ __forceinline__ bool tile_setup(const DwtParams& p, const TileList& tl, TileCtx& S) {
  const int per_pic = tl.start[p.ncomp];
  const int t = blockIdx.x;
  S.pic = t / per_pic;
  int r = t - S.pic * per_pic;
  S.comp = (p.ncomp > 1 && r >= tl.start[1]) + (p.ncomp > 2 && r >= tl.start[2]);
  r -= tl.start[S.comp];
  const DwtComp& C = p.c[S.comp];
  const int ty = r / tl.tx[S.comp], tx = r - ty * tl.tx[S.comp];
  S.x0 = tx * T::XU; S.y0 = ty * T::YU;
  S.xs = S.x0 - 8 * T::HL; S.ys = S.y0 - 8 * T::HL;
  S.plo = max(0, -S.xs / 2);
  S.phi = min(T::TW / 2 - 1, (C.lat_w - 2 - S.xs) / 2);
  S.hedge = S.xs < 0 || S.xs + T::TW > C.lat_w;
  S.vplo = max(0, -S.ys / 2);
  S.vphi = min(T::TH / 2 - 1, (C.lat_h - 2 - S.ys) / 2);
  S.vedge = S.ys < 0 || S.ys + T::TH > C.lat_h;
  return true;
}

// P2 of both directions: vertical lifting of the column chunks [first, first + count) of each half of the tile rows
template <int K, int DIR, class T>
__device__ __forceinline__ void tile_columns(int4* mid, const TileCtx& S, int first, int count) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane & (T::TVL - 1), sub = lane / T::TVL;
  const int ntasks = 2 * count;
  const bool keep = j >= T::HL && j < T::TVL - T::HL;       // the rows of the halo lanes are not needed any more
#pragma unroll 1
  for (int task = warp * T::TPW + sub; task < ntasks; task += T::NW * T::TPW) {
    const int chunk = task < count ? first + task : T::TWL + first + (task - count);
    int4* col = mid + (8 * j) * T::CHUNKS + (chunk ^ (j & 7));   // rows 8j .. 8j+7 share swz(row) = j & 7
    int e[4][4], o[4][4];   // [column][pair]: e = even rows 8j + 2a, o = odd rows 8j + 2a + 1
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int4 q0 = col[(2 * a) * T::CHUNKS], q1 = col[(2 * a + 1) * T::CHUNKS];
      e[0][a] = q0.x; e[1][a] = q0.y; e[2][a] = q0.z; e[3][a] = q0.w;
      o[0][a] = q1.x; o[1][a] = q1.y; o[2][a] = q1.z; o[3][a] = q1.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) hsteps<K, DIR, 4, T::TVL>(e[c], o[c], j, S.vedge, S.vplo, S.vphi);
    if (keep) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        col[(2 * a) * T::CHUNKS] = make_int4(e[0][a], e[1][a], e[2][a], e[3][a]);
        col[(2 * a + 1) * T::CHUNKS] = make_int4(o[0][a], o[1][a], o[2][a], o[3][a]);
      }
    }
  }
}

#if VC2_DWT_PART == 1
// ------------------------------------------------------------------------------------------
// forward level:  pix (dense plane)  ->  LL (compact plane or band 0), HL, LH, HH (interleaved)
// ------------------------------------------------------------------------------------------
template <int K, int KIND, class T>
__global__ void __launch_bounds__(32 * T::NW, T::MINB) dwt_tile_fwd_kernel(const DwtParams p, const TileList tl) {
  extern __shared__ int4 mid[];
  TileCtx S;
  tile_setup<T>(p, tl, S);
  const DwtComp& C = p.c[S.comp];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = lane & (T::TWL - 1), sub = lane / T::TWL;
  constexpr int SHIFT = Wavelet<K>::SHIFT;

  // P1: load, convert, shift in the accuracy bit (WaveletTransform.cpp:270), horizontal lifting
#pragma unroll 2
  for (int r = warp * T::RPW + sub; r < T::TH; r += T::NW * T::RPW) {
    const int y = min(max(S.ys + r, 0), C.lat_h - 1);   // rows outside the lattice are never used: any row will do
    int x[8];
    load_pix<KIND, 8>(C, S.pic, y, S.xs + 8 * hl, x);
    int e[4], o[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      e[a] = (int)((unsigned)x[2 * a] << SHIFT);
      o[a] = (int)((unsigned)x[2 * a + 1] << SHIFT);
    }
    hsteps<K, +1, 4, T::TWL>(e, o, hl, S.hedge, S.plo, S.phi);
    int4* row = mid + r * T::CHUNKS;
    row[hl ^ swz(r)] = make_int4(e[0], e[1], e[2], e[3]);
    row[(T::TWL + hl) ^ swz(r)] = make_int4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();

  // P2: vertical lifting of the useful columns
  tile_columns<K, +1, T>(mid, S, T::HL, T::TWL - 2 * T::HL);
  __syncthreads();

  // P3: the useful rows leave as band rows
  const int gx = S.xs + 8 * hl;
  const bool mine = hl >= T::HL && hl < T::TWL - T::HL && gx < C.lat_w;
  const int bx0 = (S.xs >> 1) + 4 * hl, bxmax = C.lat_w / 2 - 1;
  BandFast F;
  const bool fast = band_fast_setup(C, S.pic, bx0, mine && bx0 + 3 <= bxmax, F);
  const BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2};
  int32_t* coef = C.coef + (long long)S.pic * C.coef_pic_stride;
#pragma unroll 1
  for (int r = 8 * T::HL + warp * T::RPW + sub; r < T::TH - 8 * T::HL; r += T::NW * T::RPW) {
    const int y = S.ys + r;
    if (y >= C.lat_h || !mine) continue;
    const int4* row = mid + r * T::CHUNKS;
    const int4 lo4 = row[hl ^ swz(r)], hi4 = row[(T::TWL + hl) ^ swz(r)];
    const int by = y >> 1;
    if (fast) {
      const int i = F.idx(by);
      if (y & 1) {
        *reinterpret_cast<int4*>(F.coefpic + i + F.o_lh) = lo4;
        *reinterpret_cast<int4*>(F.coefpic + i + F.o_hh) = hi4;
      } else {
        if (F.llp) *reinterpret_cast<int4*>(F.llp + (long long)by * F.ll_pitch) = lo4;
        else *reinterpret_cast<int4*>(F.coefpic + i + F.o_ll) = lo4;
        *reinterpret_cast<int4*>(F.coefpic + i + F.o_hl) = hi4;
      }
    } else {
      int lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, hi[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
      if (y & 1) {
        band_access<4, true>(coef, ba, C.base_lh, by, bx0, bxmax, lo);
        band_access<4, true>(coef, ba, C.base_hh, by, bx0, bxmax, hi);
      } else {
        if (C.ll) ll_access<4, true>(C.ll + (long long)S.pic * C.ll_pic_stride, C.ll_pitch, by, bx0, bxmax, lo);
        else band_access<4, true>(coef, ba, C.base_ll, by, bx0, bxmax, lo);
        band_access<4, true>(coef, ba, C.base_hl, by, bx0, bxmax, hi);
      }
    }
  }
}
#else
// ------------------------------------------------------------------------------------------
// inverse level:  LL, HL, LH, HH  ->  pix (dense plane; cropped / clipped / packed at level 0)
// ------------------------------------------------------------------------------------------
template <int K, int KIND, class T>
__global__ void __launch_bounds__(32 * T::NW, T::MINB) dwt_tile_inv_kernel(const DwtParams p, const TileList tl) {
  extern __shared__ int4 mid[];
  TileCtx S;
  tile_setup<T>(p, tl, S);
  const DwtComp& C = p.c[S.comp];
  if (S.x0 >= C.pix_w || S.y0 >= C.pix_h) return;   // nothing of this tile survives the crop (WaveletTransform.cpp:340)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = lane & (T::TWL - 1), sub = lane / T::TWL;
  const int gx = S.xs + 8 * hl;
  const int bx0 = (S.xs >> 1) + 4 * hl, bxmax = C.lat_w / 2 - 1;

  // P1: band rows into the tile
  {
    BandFast F;
    const bool fast = band_fast_setup(C, S.pic, bx0, bx0 >= 0 && bx0 + 3 <= bxmax, F);
    const BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2};
    int32_t* coef = C.coef + (long long)S.pic * C.coef_pic_stride;
#pragma unroll 2
    for (int r = warp * T::RPW + sub; r < T::TH; r += T::NW * T::RPW) {
      const int y = S.ys + r;
      if (y < 0 || y >= C.lat_h) continue;     // never used (the column phase extends the sequence inside the lattice)
      const int by = y >> 1;
      int4 lo4, hi4;
      if (fast) {
        const int i = F.idx(by);
        if (y & 1) {
          lo4 = __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_lh));
          hi4 = __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_hh));
        } else {
          lo4 = F.llp ? __ldg(reinterpret_cast<const int4*>(F.llp + (long long)by * F.ll_pitch)) : __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_ll));
          hi4 = __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_hl));
        }
      } else {
        int lo[4], hi[4];
        if (y & 1) {
          band_access<4, false>(coef, ba, C.base_lh, by, bx0, bxmax, lo);
          band_access<4, false>(coef, ba, C.base_hh, by, bx0, bxmax, hi);
        } else {
          if (C.ll) ll_access<4, false>(C.ll + (long long)S.pic * C.ll_pic_stride, C.ll_pitch, by, bx0, bxmax, lo);
          else band_access<4, false>(coef, ba, C.base_ll, by, bx0, bxmax, lo);
          band_access<4, false>(coef, ba, C.base_hl, by, bx0, bxmax, hi);
        }
        lo4 = make_int4(lo[0], lo[1], lo[2], lo[3]);
        hi4 = make_int4(hi[0], hi[1], hi[2], hi[3]);
      }
      int4* row = mid + r * T::CHUNKS;
      row[hl ^ swz(r)] = lo4;
      row[(T::TWL + hl) ^ swz(r)] = hi4;
    }
  }
  __syncthreads();

  // P2: vertical inverse lifting of every column (the row phase reaches into the halo columns)
  tile_columns<K, -1, T>(mid, S, 0, T::TWL);
  __syncthreads();

  // P3: horizontal inverse lifting, rounding, clip, sample format
  const bool mine = hl >= T::HL && hl < T::TWL - T::HL && gx < C.pix_w;
#pragma unroll 1
  for (int r = 8 * T::HL + warp * T::RPW + sub; r < T::TH - 8 * T::HL; r += T::NW * T::RPW) {
    const int y = S.ys + r;
    const int4* row = mid + r * T::CHUNKS;
    const int4 lo4 = row[hl ^ swz(r)], hi4 = row[(T::TWL + hl) ^ swz(r)];
    int e[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, o[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
    hsteps<K, -1, 4, T::TWL>(e, o, hl, S.hedge, S.plo, S.phi);
    if (!mine || y >= C.pix_h) continue;
    int v[8];
#pragma unroll
    for (int a = 0; a < 4; ++a) { v[2 * a] = e[a]; v[2 * a + 1] = o[a]; }
    store_pix<K, KIND, 8>(C, S.pic, y, gx, v);
  }
}
#endif

template <int K, int KIND, class T>
cudaError_t launch_tile(cudaStream_t s, const DwtParams& p, int npictures) {
  TileList tl;
  int n = 0;
  for (int c = 0; c < 3; ++c) {
    tl.start[c] = n;
    tl.tx[c] = 1;
    if (c < p.ncomp) {
      tl.tx[c] = (p.c[c].lat_w + T::XU - 1) / T::XU;
      n += tl.tx[c] * ((p.c[c].lat_h + T::YU - 1) / T::YU);
    }
  }
  tl.start[3] = n;
  for (int c = p.ncomp; c < 3; ++c) tl.start[c] = n;
#if VC2_DWT_PART == 1
  auto kern = dwt_tile_fwd_kernel<K, KIND, T>;
#else
  auto kern = dwt_tile_inv_kernel<K, KIND, T>;
#endif
  static bool configured[16] = {};   // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 16) configured[dev] = true;
  }
  kern<<<(unsigned)((long long)n * npictures), 32 * T::NW, T::SMEM, s>>>(p, tl);
  return cudaGetLastError();
}

template <int K, int KIND>
cudaError_t pick_tile(cudaStream_t s, const DwtParams& p, int npictures, int cfg) {
  if (cfg == 2) return launch_tile<K, KIND, Tile<K, 16, 16, 8, 3>>(s, p, npictures);   // 128 x 128 tile, 64 KB, three CTAs per SM
  return launch_tile<K, KIND, Tile<K, 32, 16, 16, 1>>(s, p, npictures);               // 128 rows x 256 columns, 128 KB, one CTA per SM
}

template <int KIND>
cudaError_t dispatch_tile(cudaStream_t s, int kernel, const DwtParams& p, int npictures, int cfg) {
#define VC2_CASE(K) case K: return pick_tile<K, KIND>(s, p, npictures, cfg);
  switch (kernel) {
    VC2_CASE(VC2_DD97) VC2_CASE(VC2_LEGALL) VC2_CASE(VC2_DD137) VC2_CASE(VC2_HAAR0)
    VC2_CASE(VC2_HAAR1) VC2_CASE(VC2_FIDELITY) VC2_CASE(VC2_DAUB97)
    default: return cudaErrorInvalidValue;
  }
#undef VC2_CASE
}

}  // namespace

// cfg: 1 = 128 x 256 tile (one CTA of 16 warps per SM), 2 = 128 x 128 tile (three CTAs of 8 warps per SM)
#if VC2_DWT_PART == 1
cudaError_t dwt_tile_fwd_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int cfg) {
#else
cudaError_t dwt_tile_inv_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int cfg) {
#endif
  switch (sample_kind) {
    case SAMPLE_I32: return dispatch_tile<SAMPLE_I32>(s, kernel, p, npictures, cfg);
    case SAMPLE_U16BE: return dispatch_tile<SAMPLE_U16BE>(s, kernel, p, npictures, cfg);
    case SAMPLE_U8: return dispatch_tile<SAMPLE_U8>(s, kernel, p, npictures, cfg);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace vc2
