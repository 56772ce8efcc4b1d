// Integer lifting DWT / IDWT for the seven VC-2 wavelet kernels (sm_100a).
//
// Reference semantics reproduced bit-exactly (paths relative to /root/reference):
//   forward level : src/Library/src/WaveletTransform.cpp:478-533 (DD97), 595-644 (LeGall),
//                   700-761 (DD137), 829-871 (Haar), 919-1001 (Fidelity), 1090-1175 (Daub97)
//   inverse level : :536-593, 647-698, 764-827, 874-917, 1004-1088, 1178-1265
//   level loop    : :262-281 (forward, with waveletPad :79-94), :321-342 (inverse, crop :340)
//
// One CTA owns a tile of one level's sample lattice plus a halo (the total reach of the kernel's
// lifting steps).  Horizontal lifting runs in REGISTERS: a warp holds one 128-sample row segment,
// four samples per lane, and fetches neighbours with warp shuffles.  Vertical lifting runs on a
// shared-memory tile split into even-column (E) and odd-column (O) arrays, one thread per column,
// taps at fixed row offsets.  The reference's edge rule - a tap that falls outside the array uses
// the nearest sample of the same parity - is implemented by EXTENDING the source-parity sequence
// into the halo (registers for rows, halo rows of the tile for columns) before each step, which
// is the same thing as clamping every tap index and leaves the inner loops free of index math.
#pragma once
#include "vc2_common.cuh"

namespace vc2 {

// ------------------------------------------------------------------------------------------
// Lifting step descriptors.  A step updates all samples of parity P from the 2*N nearest samples
// of the other parity:  x[t] += SIGN * ((ADD + sum_k CLk*x[t-(2k+1)] + CRk*x[t+(2k+1)]) >> SH)
// ------------------------------------------------------------------------------------------
template <int K, int S> struct Step;
#define VC2_STEP(K, S, P_, N_, ADD_, SH_, SIGN_, L0, L1, L2, L3, R0, R1, R2, R3)                  \
  template <> struct Step<K, S> {                                                                  \
    static constexpr int P = P_, N = N_, ADD = ADD_, SH = SH_, SIGN = SIGN_;                        \
    __host__ __device__ static constexpr int cl(int k) { return k == 0 ? L0 : k == 1 ? L1 : k == 2 ? L2 : L3; } \
    __host__ __device__ static constexpr int cr(int k) { return k == 0 ? R0 : k == 1 ? R1 : k == 2 ? R2 : R3; } \
  };

// LeGall 5/3  (WaveletTransform.cpp:609-625)
VC2_STEP(VC2_LEGALL, 0, 1, 1, 1, 1, -1, 1, 0, 0, 0, 1, 0, 0, 0)
VC2_STEP(VC2_LEGALL, 1, 0, 1, 2, 2, +1, 1, 0, 0, 0, 1, 0, 0, 0)
// Deslauriers-Dubuc 9/7  (:492-511)
VC2_STEP(VC2_DD97, 0, 1, 2, 8, 4, -1, 9, -1, 0, 0, 9, -1, 0, 0)
VC2_STEP(VC2_DD97, 1, 0, 1, 2, 2, +1, 1, 0, 0, 0, 1, 0, 0, 0)
// Deslauriers-Dubuc 13/7  (:714-736)
VC2_STEP(VC2_DD137, 0, 1, 2, 8, 4, -1, 9, -1, 0, 0, 9, -1, 0, 0)
VC2_STEP(VC2_DD137, 1, 0, 2, 16, 5, +1, 9, -1, 0, 0, 9, -1, 0, 0)
// Haar, with and without shift  (:843-855)
VC2_STEP(VC2_HAAR0, 0, 1, 1, 0, 0, -1, 1, 0, 0, 0, 0, 0, 0, 0)
VC2_STEP(VC2_HAAR0, 1, 0, 1, 1, 1, +1, 0, 0, 0, 0, 1, 0, 0, 0)
VC2_STEP(VC2_HAAR1, 0, 1, 1, 0, 0, -1, 1, 0, 0, 0, 0, 0, 0, 0)
VC2_STEP(VC2_HAAR1, 1, 0, 1, 1, 1, +1, 0, 0, 0, 0, 1, 0, 0, 0)
// Fidelity: update first, then predict  (:933-964)
VC2_STEP(VC2_FIDELITY, 0, 0, 4, 128, 8, +1, 161, -46, 21, -8, 161, -46, 21, -8)
VC2_STEP(VC2_FIDELITY, 1, 1, 4, 128, 8, -1, 81, -25, 10, -2, 81, -25, 10, -2)
// Daubechies 9/7 integer approximation  (:1104-1137)
VC2_STEP(VC2_DAUB97, 0, 1, 1, 2048, 12, -1, 6497, 0, 0, 0, 6497, 0, 0, 0)
VC2_STEP(VC2_DAUB97, 1, 0, 1, 2048, 12, -1, 217, 0, 0, 0, 217, 0, 0, 0)
VC2_STEP(VC2_DAUB97, 2, 1, 1, 2048, 12, +1, 3616, 0, 0, 0, 3616, 0, 0, 0)
VC2_STEP(VC2_DAUB97, 3, 0, 1, 2048, 12, +1, 1817, 0, 0, 0, 1817, 0, 0, 0)
#undef VC2_STEP

// number of lifting steps, accuracy shift (WaveletTransform.cpp:224-260), halo per side (R) and the
// horizontal halo rounded up to whole 4-sample lane groups (HX)
template <int K> struct Wavelet;
template <> struct Wavelet<VC2_DD97>     { static constexpr int NSTEPS = 2, SHIFT = 1, R = 4,  HX = 4; };
template <> struct Wavelet<VC2_LEGALL>   { static constexpr int NSTEPS = 2, SHIFT = 1, R = 2,  HX = 4; };
template <> struct Wavelet<VC2_DD137>    { static constexpr int NSTEPS = 2, SHIFT = 1, R = 6,  HX = 8; };
template <> struct Wavelet<VC2_HAAR0>    { static constexpr int NSTEPS = 2, SHIFT = 0, R = 0,  HX = 0; };
template <> struct Wavelet<VC2_HAAR1>    { static constexpr int NSTEPS = 2, SHIFT = 1, R = 0,  HX = 0; };
template <> struct Wavelet<VC2_FIDELITY> { static constexpr int NSTEPS = 2, SHIFT = 0, R = 14, HX = 16; };
template <> struct Wavelet<VC2_DAUB97>   { static constexpr int NSTEPS = 4, SHIFT = 1, R = 4,  HX = 4; };

// ------------------------------------------------------------------------------------------
// kernel parameter blocks
// ------------------------------------------------------------------------------------------
enum SampleKind { SAMPLE_I32 = 0, SAMPLE_U16BE = 1, SAMPLE_U8 = 2 };

struct DwtComp {
  // dense plane per picture: level input (forward) / level output (inverse)
  void* pix;                 // int32 plane, or raw sample bytes at level 0 on the fused path
  long long pix_pic_stride;  // elements (int32) or bytes (raw) between consecutive pictures
  int pix_h, pix_w;          // valid dims of that plane (level 0: unpadded picture; deeper: lattice)
  int pix_pitch;             // elements per row
  int lat_h, lat_w;          // lattice dims at this level = padded dims >> level
  // slice-major coefficient block (see vc2_common.cuh) holding this level's HL, LH, HH (and LL at the last level)
  int32_t* coef;
  long long coef_pic_stride;
  int bh, bw, lgbh, lgbw;    // this level's band part per slice (rows, cols); lg = log2 or -1 if not a power of two
  int nx, NC;                // slices per row, coefficients per slice
  int base_ll, base_hl, base_lh, base_hh;   // comp_start + band_start of the four bands inside a slice
  // compact LL plane (next level's input / previous level's output); NULL at the last level (LL = band 0 in coef)
  int32_t* ll;
  long long ll_pic_stride;
  int ll_pitch;
  // raw sample conversion (Arrays.cpp:351-376, 396-397): v = (word >> sshift) - soffset
  int sshift, soffset;
  int clip_min, clip_max;    // inverse level 0 on the fused path (Picture.cpp:284-292)
  // NARROW coefficient block (tile kernels only): the block holds QUANTISED coefficients as 16-bit sign-magnitude words
  // t = 2 * |q| + (q < 0) - the index of the packer's code table - in the same group-interleaved order.  The forward
  // kernels quantise on the way out (Quantisation.cpp:69-76, one index per band: HQ_ConstQ), the inverse kernels
  // scale on the way in (Quantisation.cpp:86-95, the index of the slice each piece belongs to).  [0..3] = LL, HL, LH, HH.
  uint32_t qmul[4];          // forward: |q| = (|v| * qmul) >> qsh for |v| < VC2_NARROW_FAST_MAX (host-verified) ...
  int qsh[4];
  uint32_t qm31[4];          // ... else the one-multiply-high division of vc2_quant_magic31: |q| = mulhi(|v| << 2, qm31) >> ql31
  int ql31[4];
  int qmat[4];               // inverse: quantisation matrix entries of the four bands (Quantisation.cpp:16-20)
  int band[4];               // ... and their band numbers
};


struct DwtParams {
  DwtComp c[3];
  int ncomp;
  int pd;   // prefetch distance of the streaming kernels, in row pairs
  int fast; // 1: interior rows of interior strips run the fast loop (dwt.cu); 0: general loop only (VC2_DWT_FAST=0)
  // narrow coefficient block (see DwtComp)
  int narrow;
  uint32_t* narrow_ovf;      // [picture] set when a quantised magnitude does not fit 15 bits: the caller re-runs the picture wide
  const int32_t* qidx;       // inverse: [picture][slice] quantisation index of every slice
  const BandScale* band_scale;   // inverse: [picture] one index for the whole picture? then its scale factors per band
  int nslices;
  const uint2* scale_tab;    // inverse: [128] (quant_factor, quant_offset + 2) in device memory
};

}  // namespace vc2
