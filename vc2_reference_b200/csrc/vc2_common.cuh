// Shared device/host definitions for the B200 VC-2 hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vc2_cabi.h"

#define VC2_MAX_DEPTH 6
#define VC2_MAX_BANDS (3 * VC2_MAX_DEPTH + 1)

namespace vc2 {

// ---------------------------------------------------------------------------------------------
// Planar-subband coefficient layout used INSIDE the device pipeline.
//
// A padded component plane (ph x pw, both multiples of 2^depth) is stored as 3*depth+1 compact
// subband planes back to back: band 0 = LL (ph>>d x pw>>d); VC-2 level L = 1..d (L = d finest)
// contributes HL, LH, HH (bands 3(L-1)+1..3), each (ph >> (d-L+1)) x (pw >> (d-L+1)).
// Everything before level L sums to n0*4^(L-1) elements (n0 = LL size), so
//   offset(L, type) = n0 * 4^(L-1) * type,  type = 1 (HL), 2 (LH), 3 (HH).
// This is exactly the order in which a slice's coefficients are coded
// (reference split_into_subbands, WaveletTransform.cpp:428-450), so slice coding reads contiguous
// band rows, and every DWT level reads/writes dense rows.  The reference's in-place interleaved
// Array2D order is produced/consumed only at the Library boundary (layout kernels in dwt.cu).
// ---------------------------------------------------------------------------------------------
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
// qf = quant_factor, qo = quant_offset, (qm, ql) = Granlund-Montgomery magic for exact u32 / qf.
struct QuantTables {
  uint32_t qf[128];
  uint32_t qo[128];
  uint32_t qm[128];
  uint32_t ql[128];
};

__host__ __device__ inline int band_level(int b) { return b == 0 ? 0 : (b - 1) / 3 + 1; }

}  // namespace vc2
