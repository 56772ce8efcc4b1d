// Shared device/host definitions for the B200 VC-2 hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vc2_cabi.h"

#define VC2_MAX_DEPTH 6
#define VC2_MAX_BANDS (3 * VC2_MAX_DEPTH + 1)
// narrow coefficient block (dwt.cuh): 16-bit sign-magnitude words 2 * |q| + (q < 0)
#define VC2_NARROW_MAX_MAG 32767   /* largest |q| the block holds */
#define VC2_NARROW_FAST_MAX 8192   /* |v| below this: the quantiser's one-multiply form (verified by the host per band) */

namespace vc2 {

// ---------------------------------------------------------------------------------------------
// GROUP-INTERLEAVED SLICE LAYOUT used INSIDE the device pipeline.
//
// One picture's coefficients (all three padded component planes) are stored slice by slice.  The
// coefficients of ONE slice form a list of NC values in the reference's coding order: Y, C1, C2;
// inside a component the subbands in coding order (band 0 = LL, then VC-2 level L = 1..depth:
// HL, LH, HH = bands 3(L-1)+1..3), each subband's part of the slice in raster order
// (split_into_subbands, WaveletTransform.cpp:428-450; HQSliceIO, Slices.cpp:488-530):
//
//   k(c, b, y, x) = comp_start[c] + band_start[c][b] + y * part_w[c][b] + x          0 <= k < NC
//
// Slices are taken in raster order, s = sy * slices_x + sx, and 32 consecutive slices form a GROUP.
// Inside a group the 32 coefficient lists are interleaved in pieces of four coefficients (16 bytes):
//
//   index(s, k) = (((s / 32) * (NC / 4) + k / 4) * 32 + s % 32) * 4 + k % 4
//
// The slice coders run ONE THREAD PER SLICE, a warp = one group: lane r streams through list r, and
// every 128-bit access of the warp is one contiguous 512-byte run.  The DWT kernels address the same
// layout: four consecutive coefficients of a band row are one aligned 16-byte piece whenever
// part_w % 4 == 0.  The reference's in-place interleaved Array2D order exists only at the Library
// boundary (layout kernels in dwt.cu).
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline long long coef_index(int s, int k, int nc4) {
  return ((((long long)(s >> 5) * nc4 + (k >> 2)) * 32 + (s & 31)) << 2) + (k & 3);
}

struct PlaneGeom {
  int h, w;    // unpadded picture plane
  int ph, pw;  // padded
  int depth;
  __host__ __device__ int band_h(int b) const { return b == 0 ? (ph >> depth) : (ph >> (depth - ((b - 1) / 3 + 1) + 1)); }
  __host__ __device__ int band_w(int b) const { return b == 0 ? (pw >> depth) : (pw >> (depth - ((b - 1) / 3 + 1) + 1)); }
  __host__ __device__ long long band_off(int b) const {
    if (b == 0) return 0;
    const long long n0 = (long long)(ph >> depth) * (pw >> depth);
    const int L = (b - 1) / 3 + 1, type = (b - 1) % 3 + 1;
    return (n0 << (2 * (L - 1))) * type;
  }
  __host__ __device__ long long size() const { return (long long)ph * pw; }
};

// quantiser tables (Quantisation.cpp:40-83), filled by the host at context creation.
// qf = quant_factor, qo = quant_offset, (qm, ql) = Granlund-Montgomery magic for exact u32 / qf,
// (qm31, ql31) = the one-multiply magic that is exact for dividends below 2^31 (vc2_quant_magic31).
struct QuantTables {
  uint32_t qf[128];
  uint32_t qo[128];
  uint32_t qm[128];
  uint32_t ql[128];
  uint32_t qm31[128];
  uint32_t ql31[128];
  uint32_t qmul16[128];   // |q| = (|v| * qmul16) >> qsh16 for |v| < VC2_NARROW_FAST_MAX: one full-rate multiply (0 / 0 when no such pair exists)
  uint32_t qsh16[128];
};

// narrow decode, per picture (written by the parser, read by the inverse lifting kernels)
struct BandScale {
  uint32_t diff;                 // OR over the slices of (qindex ^ qindex of slice 0), 0 = one index for the whole picture
  uint32_t pad;
  uint2 fo[VC2_MAX_BANDS];       // (quant_factor, quant_offset + 2) of every band for the index of slice 0
};

__host__ __device__ inline int band_level(int b) { return b == 0 ? 0 : (b - 1) / 3 + 1; }

}  // namespace vc2
