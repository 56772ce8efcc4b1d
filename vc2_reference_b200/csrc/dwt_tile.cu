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
#include <cuda.h>   // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no link against libcuda)

#ifndef VC2_DWT_PART
#error "compile with -DVC2_DWT_PART=1 (forward) or =2 (inverse)"
#endif

namespace vc2 {

namespace {

// TWL lanes per tile row (8 samples each), TVL lanes per tile column group (8 rows each), NW warps per CTA
template <int K, int TWL_, int TVL_, int NW_, int MINB_, int TPC_>
struct Tile {
  static constexpr int TWL = TWL_, TVL = TVL_, NW = NW_, MINB = MINB_;   // MINB: resident CTAs per SM the registers are held to
  static constexpr int HL = (Wavelet<K>::R + 7) / 8;      // halo, in lanes, on each side (both directions)
  static constexpr int TW = 8 * TWL, TH = 8 * TVL;
  static constexpr int XU = TW - 16 * HL, YU = TH - 16 * HL;   // useful samples
  static constexpr int CHUNKS = 2 * TWL;                  // 16-byte chunks per tile row: TWL of even columns, then TWL of odd ones
  static constexpr int RPW = 32 / TWL;                    // rows per warp and iteration in the row phases
  static constexpr int TPW = 32 / TVL;                    // column tasks per warp and iteration in the column phase
  static constexpr int NG = NW * RPW;                     // row groups: a group is TWL lanes working on one row at a time
  static constexpr int SMEM = TH * TW * 4;
  static constexpr int TPC = TPC_;                        // horizontally adjacent tiles per CTA (amortises the set-up)
  static_assert(TH == 8 * NG, "a row group owns one swizzle block of eight tile rows");
};

// TMA staging of the forward level-0 input (16-bit samples, no padding, 16-byte aligned rows): one thread asks the
// copy engine for the tile's 128 x 128 samples (cp.async.bulk.tensor, zero fill outside the picture), the CTA waits
// on an mbarrier, every lane takes its eight rows out of shared memory.  The raw tile (32 KB) lives in the upper
// half of the lifting tile (64 KB): all lanes hold their rows in registers before the first lifted row is written.
struct TileMaps { CUtensorMap m[3]; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

struct TileList {     // grid.x = CTA in its tile row, grid.y = tile row (both: the largest component), grid.z = picture * ncomp + component
  int cx[3], tx[3], ty[3];   // CTAs per tile row, tiles per tile row, tile rows
};

__device__ __forceinline__ int swz(int row) { return (row >> 3) & 7; }

// per-lane constants of the vectorised band access (band part widths are multiples of four, power-of-two part heights):
// the piece (four coefficients) of band row `by` at this lane's band columns sits at element idx(by) behind the band's start
struct BandFast {
  uint8_t *p_ll, *p_hl, *p_lh, *p_hh;   // picture block + band start (p_ll: compact LL plane at this lane's columns when plane_ll)
  bool plane_ll;
  int ll_pitch;
  int esh;                              // log2 of the element size of the block: 2 (int32) or 1 (narrow, 16-bit)
  int sx, kx4, lgbh, bhm1, bw4, nx, nc4;
  __device__ __forceinline__ int slice(int by) const { return (by >> lgbh) * nx + sx; }
  __device__ __forceinline__ int idx(int by) const {
    const int ry = by & bhm1, s = slice(by);
    return ((((s >> 5) * nc4 + ry * bw4 + kx4) << 5) + (s & 31)) << 2;
  }
  template <class V> __device__ __forceinline__ V* at(uint8_t* band, int i) const { return reinterpret_cast<V*>(band + ((unsigned)i << esh)); }
  __device__ __forceinline__ int4* ll_row(int by) const { return reinterpret_cast<int4*>(p_ll + (((long long)by * ll_pitch) << 2)); }
};
// false when this level / lane has to use the element-wise access
__device__ __forceinline__ bool band_fast_setup(const DwtComp& C, int pic, int narrow, int bx0, bool inside, BandFast& F) {
  if (!inside) return false;
  if (C.lgbh < 0 || C.lgbw < 0 || (C.bw & 3) || ((C.base_ll | C.base_hl | C.base_lh | C.base_hh) & 3)) return false;
  F.esh = narrow ? 1 : 2;
  uint8_t* coefpic = reinterpret_cast<uint8_t*>(C.coef) + (((long long)pic * C.coef_pic_stride) << F.esh);
  F.sx = bx0 >> C.lgbw;
  F.kx4 = (bx0 & (C.bw - 1)) >> 2;
  F.lgbh = C.lgbh; F.bhm1 = C.bh - 1; F.bw4 = C.bw >> 2; F.nx = C.nx; F.nc4 = C.NC >> 2;
  F.p_ll = coefpic + (((C.base_ll >> 2) * 128) << F.esh); F.p_hl = coefpic + (((C.base_hl >> 2) * 128) << F.esh);
  F.p_lh = coefpic + (((C.base_lh >> 2) * 128) << F.esh); F.p_hh = coefpic + (((C.base_hh >> 2) * 128) << F.esh);
  F.plane_ll = C.ll != nullptr; F.ll_pitch = C.ll_pitch;
  if (C.ll) {
    F.p_ll = reinterpret_cast<uint8_t*>(C.ll + (long long)pic * C.ll_pic_stride + bx0);
    if ((C.ll_pitch & 3) || (reinterpret_cast<uintptr_t>(C.ll + (long long)pic * C.ll_pic_stride) & 15)) return false;
  }
  if (reinterpret_cast<uintptr_t>(coefpic) & 15) return false;
  return true;
}

// ---- narrow coefficient block: quantised, 16-bit sign-magnitude (see dwt.cuh) ----------------------------------------
// |quant(v, q)| (Quantisation.cpp:69-76) of four coefficients of band b as sign-magnitude words, two per register
__device__ __forceinline__ uint2 quant4_smag(const DwtComp& C, int b, const int4 v, bool& ovf) {
  const unsigned a0 = (unsigned)abs(v.x), a1 = (unsigned)abs(v.y), a2 = (unsigned)abs(v.z), a3 = (unsigned)abs(v.w);
  unsigned m0, m1, m2, m3;
  if ((a0 | a1 | a2 | a3) < (unsigned)VC2_NARROW_FAST_MAX) {   // one full-rate multiply: exact below the bound (checked by the host)
    const unsigned mul = C.qmul[b];
    const int sh = C.qsh[b];
    m0 = (a0 * mul) >> sh; m1 = (a1 * mul) >> sh; m2 = (a2 * mul) >> sh; m3 = (a3 * mul) >> sh;
  } else {
    const unsigned mm = C.qm31[b];
    const int sh = C.ql31[b];
    m0 = __umulhi(a0 << 2, mm) >> sh; m1 = __umulhi(a1 << 2, mm) >> sh; m2 = __umulhi(a2 << 2, mm) >> sh; m3 = __umulhi(a3 << 2, mm) >> sh;
    if ((m0 | m1 | m2 | m3) > (unsigned)VC2_NARROW_MAX_MAG) {
      ovf = true;
      m0 = min(m0, (unsigned)VC2_NARROW_MAX_MAG); m1 = min(m1, (unsigned)VC2_NARROW_MAX_MAG);
      m2 = min(m2, (unsigned)VC2_NARROW_MAX_MAG); m3 = min(m3, (unsigned)VC2_NARROW_MAX_MAG);
    }
  }
  const unsigned t0 = 2u * m0 + ((unsigned)v.x >> 31), t1 = 2u * m1 + ((unsigned)v.y >> 31);
  const unsigned t2 = 2u * m2 + ((unsigned)v.z >> 31), t3 = 2u * m3 + ((unsigned)v.w >> 31);
  return make_uint2(t0 | (t1 << 16), t2 | (t3 << 16));
}
// scale(q, index) (Quantisation.cpp:86-95) of four sign-magnitude words; fo = (quant_factor, quant_offset + 2)
__device__ __forceinline__ int scale1_smag(unsigned t, const uint2 fo) {
  const unsigned mag = t >> 1;
  const unsigned m = mag ? (mag * fo.x + fo.y) >> 2 : 0u;
  return (t & 1u) ? -(int)m : (int)m;
}
__device__ __forceinline__ int4 scale4_smag(const uint2 w, const uint2 fo) {
  return make_int4(scale1_smag(w.x & 0xFFFFu, fo), scale1_smag(w.x >> 16, fo), scale1_smag(w.y & 0xFFFFu, fo), scale1_smag(w.y >> 16, fo));
}
// The same through a table in shared memory, for pictures with one index (HQ_ConstQ streams): lut[t] = scale1_smag(t, fo) for
// the VC2_SCALE_LUT smallest sign-magnitude words of a band - two ALU instructions and a shared-memory load per coefficient
// instead of ten ALU instructions.  MEASURED SLOWER and therefore off (VC2_SCALE_LUT = 0): level 0 inverse 1.079 ms per 32 C3
// pictures without, 1.336 with 128 entries, 1.419 with 512 (bit exact all three; profiles/r2_v4_scale_lut_ab.txt) - the
// data-dependent shared-memory loads and the branch around the fallback cost more than the arithmetic they replace.
#ifndef VC2_SCALE_LUT
#define VC2_SCALE_LUT 0
#endif
__device__ __forceinline__ int4 scale4_lut(const uint2 w, const int* __restrict__ lut, const uint2 fo) {
  if (VC2_SCALE_LUT == 0) return scale4_smag(w, fo);
  if (((w.x | w.y) & ~(unsigned)(((VC2_SCALE_LUT - 1) << 16) | (VC2_SCALE_LUT - 1))) == 0u)
    return make_int4(lut[w.x & 0xFFFFu], lut[w.x >> 16], lut[w.y & 0xFFFFu], lut[w.y >> 16]);
  return scale4_smag(w, fo);
}
__device__ __forceinline__ uint2 scale_params(const DwtParams& p, const DwtComp& C, int pic, int slice, int b) {
  const int q = min(max(__ldg(p.qidx + (long long)pic * p.nslices + slice) - C.qmat[b], 0), 127);
  return __ldg(p.scale_tab + q);
}
// element-wise access to the narrow block (small band parts at the deep levels, picture edges)
// uni: the whole picture has one index, fo_uni = its scale factors for this band
template <bool STORE>
__device__ __forceinline__ void band_access16(const DwtParams& p, const DwtComp& C, int pic, const BandAddr& ba, int b, int base, int by, int bx0, int bxmax,
                                              int (&x)[4], bool& ovf, bool uni = false, uint2 fo_uni = make_uint2(0, 0)) {
  uint16_t* coef = reinterpret_cast<uint16_t*>(C.coef) + (long long)pic * C.coef_pic_stride;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int bx = bx0 + j;
    if (bx < 0 || bx > bxmax) { if (!STORE) x[j] = 0; continue; }
    uint16_t* e = coef + ba.at(base, by, bx);
    if (STORE) {
      const unsigned a = (unsigned)abs(x[j]);
      unsigned m = a < (unsigned)VC2_NARROW_FAST_MAX ? (a * C.qmul[b]) >> C.qsh[b] : __umulhi(a << 2, C.qm31[b]) >> C.ql31[b];
      if (m > (unsigned)VC2_NARROW_MAX_MAG) { ovf = true; m = VC2_NARROW_MAX_MAG; }
      *e = (uint16_t)(2u * m + ((unsigned)x[j] >> 31));
    } else {
      if (uni) x[j] = scale1_smag(*e, fo_uni);
      else {
        const int sy = ba.lgbh >= 0 ? (by >> ba.lgbh) : (by / ba.bh), sx = ba.lgbw >= 0 ? (bx >> ba.lgbw) : (bx / ba.bw);
        x[j] = scale1_smag(*e, scale_params(p, C, pic, sy * ba.nx + sx, b));
      }
    }
  }
}

struct TileCtx {
  int pic, comp;
  int x0, y0, xs, ys;          // first useful sample; first sample of the tile (halo included)
  int plo, phi, vplo, vphi;    // valid pair range of a tile row / of a tile column, in pairs from the tile origin
  bool hedge, vedge;
  int tx, tx_end;              // this CTA's tiles of the tile row: [tx, tx_end)
};

// the CTA's place; false when the component is smaller than the grid (chroma)
template <class T>
__device__ __forceinline__ bool tile_setup(const DwtParams& p, const TileList& tl, TileCtx& S) {
  const int z = blockIdx.z;
  S.pic = p.ncomp == 3 ? z / 3 : z;
  S.comp = p.ncomp == 3 ? z - 3 * S.pic : 0;
  const int cx = blockIdx.x, ty = blockIdx.y;
  const int ncx = S.comp == 0 ? tl.cx[0] : tl.cx[1], nty = S.comp == 0 ? tl.ty[0] : tl.ty[1];   // the chroma planes are alike
  if (cx >= ncx || ty >= nty) return false;
  const DwtComp& C = p.c[S.comp];
  S.tx = cx * T::TPC;
  S.tx_end = min(S.tx + T::TPC, S.comp == 0 ? tl.tx[0] : tl.tx[1]);
  S.y0 = ty * T::YU;
  S.ys = S.y0 - 8 * T::HL;
  S.vplo = max(0, -S.ys / 2);
  S.vphi = min(T::TH / 2 - 1, (C.lat_h - 2 - S.ys) / 2);
  S.vedge = S.ys < 0 || S.ys + T::TH > C.lat_h;
  return true;
}
template <class T>
__device__ __forceinline__ void tile_column_setup(const DwtComp& C, TileCtx& S, int tx) {
  S.x0 = tx * T::XU;
  S.xs = S.x0 - 8 * T::HL;
  S.plo = max(0, -S.xs / 2);
  S.phi = min(T::TW / 2 - 1, (C.lat_w - 2 - S.xs) / 2);
  S.hedge = S.xs < 0 || S.xs + T::TW > C.lat_w;
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
    hsteps_rows<K, DIR, 4, T::TVL, 4>(e, o, j, S.vedge, S.vplo, S.vphi);
    if (keep) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        col[(2 * a) * T::CHUNKS] = make_int4(e[0][a], e[1][a], e[2][a], e[3][a]);
        col[(2 * a + 1) * T::CHUNKS] = make_int4(o[0][a], o[1][a], o[2][a], o[3][a]);
      }
    }
  }
}

__device__ __forceinline__ int mad_lo(int a, int b, int c) {   // a * b + c as ONE multiply-add (not re-associated by the compiler)
  int r;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

#if VC2_DWT_PART == 1
// per-lane constants of the picture row loads
struct RowLoad {
  const uint8_t* lane_base;   // first byte of this lane's eight samples in picture row 0
  int pitchB;                 // bytes per picture row (a plane stays below 2 GB)
  int last_row;               // waveletPad: rows beyond it replicate it (WaveletTransform.cpp:88)
  int mul, nsub;              // v = (word >> sshift) * mul + nsub: sample offset and accuracy shift in one multiply-add
  int sshift;
  bool vec, skip;             // whole 16-byte pieces inside the picture | nothing of this lane is inside the lattice
};

// the raw words of this lane's eight samples of lattice row y (whole 16-byte pieces inside the picture)
template <int KIND> struct RawRow { uint4 a, b; };
template <int KIND>
__device__ __forceinline__ RawRow<KIND> tile_load_raw(const RowLoad& L, int y) {
  RawRow<KIND> r;
  r.a = make_uint4(0, 0, 0, 0); r.b = r.a;
  if (L.vec) {
    const uint8_t* p = L.lane_base + (unsigned)(min(max(y, 0), L.last_row) * L.pitchB);
    r.a = __ldg(reinterpret_cast<const uint4*>(p));
    if (KIND == SAMPLE_I32) r.b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
  return r;
}
// converted and shifted (Arrays.cpp:351-376, WaveletTransform.cpp:270), as four pairs
template <int KIND>
__device__ __forceinline__ void tile_convert(const RowLoad& L, const RawRow<KIND>& r, int (&e)[4], int (&o)[4]) {
  if (KIND == SAMPLE_I32) {
    e[0] = (int)r.a.x * L.mul; o[0] = (int)r.a.y * L.mul; e[1] = (int)r.a.z * L.mul; o[1] = (int)r.a.w * L.mul;
    e[2] = (int)r.b.x * L.mul; o[2] = (int)r.b.y * L.mul; e[3] = (int)r.b.z * L.mul; o[3] = (int)r.b.w * L.mul;
  } else {
    const unsigned w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      e[a] = mad_lo((int)(__byte_perm(w[a], 0, 0x4401) >> L.sshift), L.mul, L.nsub);
      o[a] = mad_lo((int)(__byte_perm(w[a], 0, 0x4423) >> L.sshift), L.mul, L.nsub);
    }
  }
}
// the same for a row that is not made of whole aligned pieces (picture edge, odd widths, 8-bit samples)
template <int KIND>
__device__ __forceinline__ void tile_load_row_slow(const DwtComp& C, int pic, const RowLoad& L, int y, int gx, int (&e)[4], int (&o)[4]) {
  if (!L.skip) {
    int x[8];
    load_pix<KIND, 8>(C, pic, min(max(y, 0), L.last_row), gx, x);
#pragma unroll
    for (int a = 0; a < 4; ++a) { e[a] = x[2 * a] * L.mul; o[a] = x[2 * a + 1] * L.mul; }
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) { e[a] = 0; o[a] = 0; }
  }
}

// ------------------------------------------------------------------------------------------
// forward level:  pix (dense plane)  ->  LL (compact plane or band 0), HL, LH, HH (interleaved)
// ------------------------------------------------------------------------------------------
template <int K, int KIND, class T>
__global__ void __launch_bounds__(32 * T::NW, T::MINB) dwt_tile_fwd_kernel(const DwtParams p, const TileList tl, const __grid_constant__ TileMaps maps, const int use_tma) {
  extern __shared__ __align__(128) int4 mid[];
  __shared__ unsigned long long tma_bar;
  TileCtx S;
  if (!tile_setup<T>(p, tl, S)) return;
  const DwtComp& C = p.c[S.comp];   // (a copy in shared memory instead of these indexed constant loads was measured: 10 % slower)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = lane & (T::TWL - 1), g = warp * T::RPW + lane / T::TWL;   // lane inside its row, row group
  constexpr int SHIFT = Wavelet<K>::SHIFT;
  constexpr int esz = KIND == SAMPLE_I32 ? 4 : KIND == SAMPLE_U16BE ? 2 : 1;
  const uint8_t* picbase = (const uint8_t*)C.pix + (long long)S.pic * C.pix_pic_stride * (KIND == SAMPLE_I32 ? 4 : 1);
  const BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2};
  int32_t* coef = C.coef + (long long)S.pic * C.coef_pic_stride;
  const int bxmax = C.lat_w / 2 - 1;
  const int r0 = 8 * g;                                   // P1: this group's eight tile rows
  int4* rows = mid + r0 * T::CHUNKS;
  const int ce = hl ^ swz(r0), co = (T::TWL + hl) ^ swz(r0);

#pragma unroll 1
  for (int tx = S.tx; tx < S.tx_end; ++tx) {
    tile_column_setup<T>(C, S, tx);
    const int gx = S.xs + 8 * hl;
    // P1: load, convert, shift in the accuracy bit, horizontal lifting
    {
      RowLoad L;
      L.pitchB = C.pix_pitch * esz;
      L.lane_base = picbase + (long long)gx * esz;
      L.last_row = min(C.lat_h, C.pix_h) - 1;
      L.sshift = C.sshift;
      L.mul = 1 << SHIFT;
      L.nsub = KIND == SAMPLE_I32 ? 0 : -(C.soffset << SHIFT);
      L.vec = KIND != SAMPLE_U8 && gx >= 0 && gx + 7 < C.pix_w && ((L.pitchB | reinterpret_cast<uintptr_t>(L.lane_base)) & 15) == 0;
      L.skip = gx + 7 < 0 || gx >= C.lat_w;
      auto lift_store = [&](int i, int (&e)[2][4], int (&o)[2][4]) {
        hsteps_rows<K, +1, 4, T::TWL, 2>(e, o, hl, S.hedge, S.plo, S.phi);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          rows[(i + k) * T::CHUNKS + ce] = make_int4(e[k][0], e[k][1], e[k][2], e[k][3]);
          rows[(i + k) * T::CHUNKS + co] = make_int4(o[k][0], o[k][1], o[k][2], o[k][3]);
        }
      };
      if (KIND == SAMPLE_U16BE && use_tma) {
        // the tile's raw samples through the copy engine (coordinates may lie outside the picture: zero fill, never used)
        uint8_t* rawt = reinterpret_cast<uint8_t*>(mid) + T::SMEM / 2;
        const unsigned phase = (unsigned)(tx - S.tx) & 1u;
        if (threadIdx.x == 0) {
          if (tx == S.tx) mbar_init(&tma_bar, 1);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile's earlier generic-proxy accesses are ordered before the copy
          mbar_expect_tx(&tma_bar, T::TH * T::TW * 2);
          tma_load_3d(rawt, &maps.m[S.comp], &tma_bar, S.xs, S.ys, S.pic);
        }
        __syncthreads();            // the barrier is initialised (first tile) before anyone waits on it
        mbar_wait(&tma_bar, phase);
        RawRow<KIND> raw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          raw[i].a = *reinterpret_cast<const uint4*>(rawt + (r0 + i) * (T::TW * 2) + hl * 16);
          raw[i].b = raw[i].a;
        }
        __syncthreads();            // every lane holds its rows: the lifted rows may overwrite the raw tile
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          int e[2][4], o[2][4];
          tile_convert<KIND>(L, raw[i], e[0], o[0]);
          tile_convert<KIND>(L, raw[i + 1], e[1], o[1]);
          lift_store(i, e, o);
        }
      } else if (__all_sync(FULL, L.vec || L.skip)) {
        // every load of the row group is in flight before the first one is used: one memory round trip per tile
        constexpr int NB = KIND == SAMPLE_I32 ? 4 : 8;   // rows per batch (32 registers of raw words)
#pragma unroll 1
        for (int b = 0; b < 8; b += NB) {
          RawRow<KIND> raw[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) raw[i] = tile_load_raw<KIND>(L, S.ys + r0 + b + i);
#pragma unroll
          for (int i = 0; i < NB; i += 2) {
            int e[2][4], o[2][4];
            tile_convert<KIND>(L, raw[i], e[0], o[0]);
            tile_convert<KIND>(L, raw[i + 1], e[1], o[1]);
            lift_store(b + i, e, o);
          }
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < 8; i += 2) {
          int e[2][4], o[2][4];
          tile_load_row_slow<KIND>(C, S.pic, L, S.ys + r0 + i, gx, e[0], o[0]);
          tile_load_row_slow<KIND>(C, S.pic, L, S.ys + r0 + i + 1, gx, e[1], o[1]);
          lift_store(i, e, o);
        }
      }
    }
    __syncthreads();

    // P2: vertical lifting of the useful columns
    tile_columns<K, +1, T>(mid, S, T::HL, T::TWL - 2 * T::HL);
    __syncthreads();

    // P3: the useful rows leave as band rows, two lattice rows (one band row of LL | HL and of LH | HH) at a time
    const bool mine = hl >= T::HL && hl < T::TWL - T::HL && gx < C.lat_w;
    if (mine) {
      const int bx0 = (S.xs >> 1) + 4 * hl;
      BandFast F;
      const bool fast = band_fast_setup(C, S.pic, p.narrow, bx0, bx0 + 3 <= bxmax, F);
      bool ovf = false;
#pragma unroll 1
      for (int m = 4 * T::HL + g; m < T::TH / 2 - 4 * T::HL; m += T::NG) {
        const int y = S.ys + 2 * m;
        if (y >= C.lat_h) break;
        const int4* row = mid + 2 * m * T::CHUNKS;
        const int c0 = hl ^ swz(2 * m), c1 = (T::TWL + hl) ^ swz(2 * m);
        const int4 ll4 = row[c0], hl4 = row[c1], lh4 = row[T::CHUNKS + c0], hh4 = row[T::CHUNKS + c1];
        const int by = y >> 1;
        if (fast) {
          const int i = F.idx(by);
          if (p.narrow) {
            if (F.plane_ll) *F.ll_row(by) = ll4;
            else *F.at<uint2>(F.p_ll, i) = quant4_smag(C, 0, ll4, ovf);
            *F.at<uint2>(F.p_hl, i) = quant4_smag(C, 1, hl4, ovf);
            *F.at<uint2>(F.p_lh, i) = quant4_smag(C, 2, lh4, ovf);
            *F.at<uint2>(F.p_hh, i) = quant4_smag(C, 3, hh4, ovf);
          } else {
            if (F.plane_ll) *F.ll_row(by) = ll4;
            else *F.at<int4>(F.p_ll, i) = ll4;
            *F.at<int4>(F.p_hl, i) = hl4;
            *F.at<int4>(F.p_lh, i) = lh4;
            *F.at<int4>(F.p_hh, i) = hh4;
          }
        } else {
          int v[4] = {ll4.x, ll4.y, ll4.z, ll4.w};
          if (C.ll) ll_access<4, true>(C.ll + (long long)S.pic * C.ll_pic_stride, C.ll_pitch, by, bx0, bxmax, v);
          else if (p.narrow) band_access16<true>(p, C, S.pic, ba, 0, C.base_ll, by, bx0, bxmax, v, ovf);
          else band_access<4, true>(coef, ba, C.base_ll, by, bx0, bxmax, v);
          v[0] = hl4.x; v[1] = hl4.y; v[2] = hl4.z; v[3] = hl4.w;
          if (p.narrow) band_access16<true>(p, C, S.pic, ba, 1, C.base_hl, by, bx0, bxmax, v, ovf);
          else band_access<4, true>(coef, ba, C.base_hl, by, bx0, bxmax, v);
          v[0] = lh4.x; v[1] = lh4.y; v[2] = lh4.z; v[3] = lh4.w;
          if (p.narrow) band_access16<true>(p, C, S.pic, ba, 2, C.base_lh, by, bx0, bxmax, v, ovf);
          else band_access<4, true>(coef, ba, C.base_lh, by, bx0, bxmax, v);
          v[0] = hh4.x; v[1] = hh4.y; v[2] = hh4.z; v[3] = hh4.w;
          if (p.narrow) band_access16<true>(p, C, S.pic, ba, 3, C.base_hh, by, bx0, bxmax, v, ovf);
          else band_access<4, true>(coef, ba, C.base_hh, by, bx0, bxmax, v);
        }
      }
      if (ovf) atomicOr(p.narrow_ovf + S.pic, 1u);
    }
    __syncthreads();   // the tile is free for the CTA's next column
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
  if (!tile_setup<T>(p, tl, S)) return;
  const DwtComp& C = p.c[S.comp];   // (a copy in shared memory instead of these indexed constant loads was measured: 10 % slower)
  if (S.y0 >= C.pix_h) return;   // nothing of this tile row survives the crop (WaveletTransform.cpp:340)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = lane & (T::TWL - 1), g = warp * T::RPW + lane / T::TWL;
  const int bxmax = C.lat_w / 2 - 1;
  const BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2};
  int32_t* coef = C.coef + (long long)S.pic * C.coef_pic_stride;
  // narrow block of a picture with one index: the scale() of the small words of the four bands as a table behind the tile
  int* const slut = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(mid) + T::SMEM);
  bool use_lut = false;
  if (p.narrow) {
    const BandScale* bs = p.band_scale + S.pic;
    use_lut = VC2_SCALE_LUT > 0 && __ldg(&bs->diff) == 0u;
    if (use_lut) {
      for (int i = threadIdx.x; i < 4 * VC2_SCALE_LUT; i += 32 * T::NW)
        slut[i] = scale1_smag((unsigned)(i % (VC2_SCALE_LUT > 0 ? VC2_SCALE_LUT : 1)), __ldg(&bs->fo[C.band[i / (VC2_SCALE_LUT > 0 ? VC2_SCALE_LUT : 1)]]));
      __syncthreads();
    }
  }

#pragma unroll 1
  for (int tx = S.tx; tx < S.tx_end; ++tx) {
    tile_column_setup<T>(C, S, tx);
    if (S.x0 >= C.pix_w) break;
    const int gx = S.xs + 8 * hl;
    const int bx0 = (S.xs >> 1) + 4 * hl;
    // P1: band rows into the tile, two lattice rows (one band row of LL | HL and of LH | HH) at a time
    {
      BandFast F;
      const bool fast = band_fast_setup(C, S.pic, p.narrow, bx0, bx0 >= 0 && bx0 + 3 <= bxmax, F);
      bool never = false;
      // one index for the whole picture (HQ_ConstQ streams): the scale factors are four constants, left behind by the parser
      bool uni = false;
      uint2 fu[4] = {make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0), make_uint2(0, 0)};
      if (p.narrow) {
        const BandScale* bs = p.band_scale + S.pic;
        uni = __ldg(&bs->diff) == 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) fu[b] = __ldg(&bs->fo[C.band[b]]);
      }
      auto put = [&](int m, const int4& ll4, const int4& hl4, const int4& lh4, const int4& hh4) {
        int4* row = mid + 2 * m * T::CHUNKS;
        const int c0 = hl ^ swz(2 * m), c1 = (T::TWL + hl) ^ swz(2 * m);
        row[c0] = ll4; row[c1] = hl4; row[T::CHUNKS + c0] = lh4; row[T::CHUNKS + c1] = hh4;
      };
      const bool lane_out = bx0 + 3 < 0 || bx0 > bxmax;   // nothing of this lane is inside the lattice: its columns are never used
      if (__all_sync(FULL, fast || lane_out)) {
        // whole pieces only.  Unrolled: all the loads of the lane are in flight together
        if (p.narrow) {
#ifndef VC2_INV_NB
#define VC2_INV_NB 2
#endif
          constexpr int NI = T::TH / 2 / T::NG, NB = VC2_INV_NB;   // band rows per lane, and per batch of loads
          static_assert(NI % NB == 0, "whole batches");
#pragma unroll 1
          for (int i0 = 0; i0 < NI; i0 += NB) {
            uint2 w[NB][4];
            int4 l4[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
              const int m = g + (i0 + i) * T::NG, y = S.ys + 2 * m;
              const bool on = fast && y >= 0 && y < C.lat_h;
              const int by = on ? y >> 1 : 0, k = on ? F.idx(by) : 0;
              l4[i] = make_int4(0, 0, 0, 0);
              w[i][0] = make_uint2(0, 0);
              if (on) {
                if (F.plane_ll) l4[i] = __ldg(F.ll_row(by)); else w[i][0] = __ldg(F.at<const uint2>(F.p_ll, k));
                w[i][1] = __ldg(F.at<const uint2>(F.p_hl, k)); w[i][2] = __ldg(F.at<const uint2>(F.p_lh, k)); w[i][3] = __ldg(F.at<const uint2>(F.p_hh, k));
              }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
              const int m = g + (i0 + i) * T::NG, y = S.ys + 2 * m;
              if (!fast || y < 0 || y >= C.lat_h) continue;
              if (uni) {
                put(m, F.plane_ll ? l4[i] : scale4_lut(w[i][0], slut, fu[0]), scale4_lut(w[i][1], slut + VC2_SCALE_LUT, fu[1]),
                    scale4_lut(w[i][2], slut + 2 * VC2_SCALE_LUT, fu[2]), scale4_lut(w[i][3], slut + 3 * VC2_SCALE_LUT, fu[3]));
                continue;
              }
              const int sl = F.slice(y >> 1);
              const uint2 f0 = scale_params(p, C, S.pic, sl, 0), f1 = scale_params(p, C, S.pic, sl, 1);
              const uint2 f2 = scale_params(p, C, S.pic, sl, 2), f3 = scale_params(p, C, S.pic, sl, 3);
              put(m, F.plane_ll ? l4[i] : scale4_smag(w[i][0], f0), scale4_smag(w[i][1], f1), scale4_smag(w[i][2], f2), scale4_smag(w[i][3], f3));
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < T::TH / 2 / T::NG; ++i) {
            const int m = g + i * T::NG, y = S.ys + 2 * m;
            if (!fast || y < 0 || y >= C.lat_h) continue;
            const int by = y >> 1, k = F.idx(by);
            put(m, F.plane_ll ? __ldg(F.ll_row(by)) : __ldg(F.at<const int4>(F.p_ll, k)), __ldg(F.at<const int4>(F.p_hl, k)),
                __ldg(F.at<const int4>(F.p_lh, k)), __ldg(F.at<const int4>(F.p_hh, k)));
          }
        }
      } else {
#pragma unroll 1
        for (int m = g; m < T::TH / 2; m += T::NG) {
          const int y = S.ys + 2 * m;
          if (y < 0 || y >= C.lat_h) continue;     // never used (the column phase extends the sequence inside the lattice)
          const int by = y >> 1;
          int4 ll4, hl4, lh4, hh4;
          int v[4];
          if (C.ll) ll_access<4, false>(C.ll + (long long)S.pic * C.ll_pic_stride, C.ll_pitch, by, bx0, bxmax, v);
          else if (p.narrow) band_access16<false>(p, C, S.pic, ba, 0, C.base_ll, by, bx0, bxmax, v, never, uni, fu[0]);
          else band_access<4, false>(coef, ba, C.base_ll, by, bx0, bxmax, v);
          ll4 = make_int4(v[0], v[1], v[2], v[3]);
          if (p.narrow) band_access16<false>(p, C, S.pic, ba, 1, C.base_hl, by, bx0, bxmax, v, never, uni, fu[1]);
          else band_access<4, false>(coef, ba, C.base_hl, by, bx0, bxmax, v);
          hl4 = make_int4(v[0], v[1], v[2], v[3]);
          if (p.narrow) band_access16<false>(p, C, S.pic, ba, 2, C.base_lh, by, bx0, bxmax, v, never, uni, fu[2]);
          else band_access<4, false>(coef, ba, C.base_lh, by, bx0, bxmax, v);
          lh4 = make_int4(v[0], v[1], v[2], v[3]);
          if (p.narrow) band_access16<false>(p, C, S.pic, ba, 3, C.base_hh, by, bx0, bxmax, v, never, uni, fu[3]);
          else band_access<4, false>(coef, ba, C.base_hh, by, bx0, bxmax, v);
          hh4 = make_int4(v[0], v[1], v[2], v[3]);
          put(m, ll4, hl4, lh4, hh4);
        }
      }
    }
    __syncthreads();

    // P2: vertical inverse lifting of every column (the row phase reaches into the halo columns)
    tile_columns<K, -1, T>(mid, S, 0, T::TWL);
    __syncthreads();

    // P3: horizontal inverse lifting, rounding, clip, sample format, two rows at a time
    const bool mine = hl >= T::HL && hl < T::TWL - T::HL && gx < C.pix_w;
#pragma unroll 1
    for (int m = 4 * T::HL + g; m < T::TH / 2 - 4 * T::HL; m += T::NG) {
      const int4* row = mid + 2 * m * T::CHUNKS;
      const int c0 = hl ^ swz(2 * m), c1 = (T::TWL + hl) ^ swz(2 * m);
      const int4 a0 = row[c0], a1 = row[c1], b0 = row[T::CHUNKS + c0], b1 = row[T::CHUNKS + c1];
      int e[2][4] = {{a0.x, a0.y, a0.z, a0.w}, {b0.x, b0.y, b0.z, b0.w}}, o[2][4] = {{a1.x, a1.y, a1.z, a1.w}, {b1.x, b1.y, b1.z, b1.w}};
      hsteps_rows<K, -1, 4, T::TWL, 2>(e, o, hl, S.hedge, S.plo, S.phi);
      const int y = S.ys + 2 * m;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (!mine || y + k >= C.pix_h) continue;
        int v[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) { v[2 * a] = e[k][a]; v[2 * a + 1] = o[k][a]; }
        store_pix<K, KIND, 8>(C, S.pic, y + k, gx, v);
      }
    }
    __syncthreads();   // the tile is free for the CTA's next column
  }
}
#endif

#if VC2_DWT_PART == 1
static bool tma_wanted() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VC2_DWT_TMA"); v = e ? atoi(e) != 0 : 0; }
  return v != 0;
}
// tensor map of one component plane of a batch: (columns, rows, pictures) of 16-bit words, box = one tile
static int make_tile_map(CUtensorMap* m, const void* base, int w, int h, int npictures, unsigned long long pitchB, unsigned long long picB, int box_w, int box_h) {
  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 0;
    enc = reinterpret_cast<Encode>(fn);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)npictures};
  const cuuint64_t strides[2] = {pitchB, picB};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 1 : 0;
}
#endif

template <int K, int KIND, class T>
cudaError_t launch_tile(cudaStream_t s, const DwtParams& p, int npictures) {
  TileList tl;
  int gx = 0, gy = 0;
  if (p.ncomp != 1 && p.ncomp != 3) return cudaErrorInvalidValue;
  if (p.ncomp == 3 && (p.c[1].lat_w != p.c[2].lat_w || p.c[1].lat_h != p.c[2].lat_h)) return cudaErrorInvalidValue;
  for (int c = 0; c < 3; ++c) {
    tl.tx[c] = tl.cx[c] = tl.ty[c] = 1;
    if (c < p.ncomp) {
      tl.tx[c] = (p.c[c].lat_w + T::XU - 1) / T::XU;
      tl.cx[c] = (tl.tx[c] + T::TPC - 1) / T::TPC;
      tl.ty[c] = (p.c[c].lat_h + T::YU - 1) / T::YU;
      gx = tl.cx[c] > gx ? tl.cx[c] : gx;
      gy = tl.ty[c] > gy ? tl.ty[c] : gy;
    }
  }
#if VC2_DWT_PART == 1
  auto kern = dwt_tile_fwd_kernel<K, KIND, T>;
  // TMA staging of the input: 16-bit samples, no padding, rows and pictures on 16-byte boundaries (VC2_DWT_TMA=1)
  alignas(64) TileMaps maps;   // (contents irrelevant when use_tma == 0)
  memset(&maps, 0, sizeof(maps));
  int use_tma = 0;
  if (KIND == SAMPLE_U16BE && tma_wanted()) {
    use_tma = 1;
    for (int c = 0; c < p.ncomp && use_tma; ++c) {
      const DwtComp& C = p.c[c];
      const unsigned long long pitchB = (unsigned long long)C.pix_pitch * 2;
      if (C.pix_w != C.lat_w || C.pix_h != C.lat_h || (pitchB & 15) || (C.pix_pic_stride & 15) || (reinterpret_cast<uintptr_t>(C.pix) & 15)) use_tma = 0;
      else use_tma = make_tile_map(&maps.m[c], C.pix, C.pix_w, C.pix_h, npictures, pitchB, (unsigned long long)C.pix_pic_stride, T::TW, T::TH);
    }
  }
#else
  auto kern = dwt_tile_inv_kernel<K, KIND, T>;
#endif
#if VC2_DWT_PART == 1
  const int smem_bytes = T::SMEM;
#else
  const int smem_bytes = T::SMEM + 4 * VC2_SCALE_LUT * (int)sizeof(int);   // + the scale() table of the narrow path
#endif
  static bool configured[16] = {};   // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
#ifdef VC2_TILE_CARVEOUT
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, VC2_TILE_CARVEOUT);
#endif
    if (dev >= 0 && dev < 16) configured[dev] = true;
  }
  if (gy > 65535 || (long long)npictures * p.ncomp > 65535) return cudaErrorInvalidValue;
#if VC2_DWT_PART == 1
  kern<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)(npictures * p.ncomp)), 32 * T::NW, smem_bytes, s>>>(p, tl, maps, use_tma);
#else
  kern<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)(npictures * p.ncomp)), 32 * T::NW, smem_bytes, s>>>(p, tl);
#endif
  return cudaGetLastError();
}

template <int K, int KIND>
cudaError_t pick_tile(cudaStream_t s, const DwtParams& p, int npictures, int cfg) {
#ifdef VC2_TILE_EXPERIMENT   // more shapes of the bench wavelet for A/B runs on the GPU box
  if constexpr (K == VC2_DD137 && KIND != SAMPLE_U8) {
    if (cfg == 2) return launch_tile<K, KIND, Tile<K, 16, 16, 8, 3, 4>>(s, p, npictures);
    if (cfg == 4) return launch_tile<K, KIND, Tile<K, 16, 16, 8, 2, 1>>(s, p, npictures);
    if (cfg == 5) return launch_tile<K, KIND, Tile<K, 32, 16, 16, 1, 1>>(s, p, npictures);
  }
#endif
#ifndef VC2_TILE_MINB
#define VC2_TILE_MINB 3
#endif
  return launch_tile<K, KIND, Tile<K, 16, 16, 8, VC2_TILE_MINB, 1>>(s, p, npictures);   // 128 x 128 tile, 64 KB, three CTAs of 8 warps per SM
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

// cfg: tile shape for A/B runs (VC2_TILE_EXPERIMENT builds); the product uses 128 x 128 tiles, three CTAs of 8 warps per SM
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
