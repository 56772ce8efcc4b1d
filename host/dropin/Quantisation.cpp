// Drop-in body for src/Library/src/Quantisation.cpp of bbc/vc2-reference: the declarations of the reference's own
// Quantisation.h, implemented over the CUDA C-ABI (include/vc2_cabi.h).  See vc2_dropin.h.
//   quantise_transform_np / inverse_quantise_transform_np   Quantisation.cpp:479-558, 586-625 -> vc2_quantise_np / vc2_dequantise_np
//   quantise_transform / inverse_quantise_transform (LD)     :213-379, 560-584                 -> vc2_quantise_ld / vc2_dequantise_ld
//   yss_for_slice                                            :627-642  (quantise + restore on the GPU, the sum of squares here)
//   quant, scale, adjust_quant_index, predictDC              :16-20, 40-95, 191-208  scalar int -> int helpers, host
// Not provided: quantise_block / inverse_quantise_block on array VIEWS (Quantisation.h:30-49) - internal building
// blocks of the reference's own Quantisation.cpp that no caller outside that file uses.
#include <algorithm>
#include <numeric>
#include "Quantisation.h"
#include "vc2_dropin.h"

using vc2dropin::check;
using vc2dropin::ctx;

namespace {
int factor(int q) {
  if (q > 119) throw std::logic_error("quantization index exceeds maximum implemented value.");
  return vc2_quant_factor(q < 0 ? 0 : q);
}
int depthOf(const Array1D& qMatrix) { return ((int)qMatrix.size() - 1) / 3; }
Array2D uniform(int q) {
  Array2D a(extents[1][1]);
  a[0][0] = q;
  return a;
}
typedef int (*QuantFn)(vc2_ctx*, const int32_t*, int, int, int, const int32_t*, const int32_t*, int, int, int32_t*);
Array2D run(QuantFn fn, const Array2D& in, const Array2D& qIndices, const Array1D& qMatrix) {
  Array2D out(extents[in.shape()[0]][in.shape()[1]]);
  check(fn(ctx(), in.data(), (int)in.shape()[0], (int)in.shape()[1], depthOf(qMatrix), qMatrix.data(), qIndices.data(),
           (int)qIndices.shape()[0], (int)qIndices.shape()[1], out.data()));
  return out;
}
Picture perPlane(QuantFn fn, const Picture& p, const Array2D& qIndices, const Array1D& qMatrix) {
  return Picture(p.format(), run(fn, p.y(), qIndices, qMatrix), run(fn, p.c1(), qIndices, qMatrix), run(fn, p.c2(), qIndices, qMatrix));
}
}  // namespace

// ---- scalar helpers ----------------------------------------------------------------------------------------

const int adjust_quant_index(const int qIndex, const int qMatrix) { return std::max(qIndex - qMatrix, 0); }

const Array1D adjust_quant_indices(const Array1D& qIndices, const int qMatrix) {
  Array1D out(extents[qIndices.size()]);
  for (size_t i = 0; i < qIndices.size(); ++i) out[i] = adjust_quant_index(qIndices[i], qMatrix);
  return out;
}

const Array2D adjust_quant_indices(const Array2D& qIndices, const int qMatrix) {
  Array2D out(extents[qIndices.shape()[0]][qIndices.shape()[1]]);
  for (size_t i = 0; i < qIndices.num_elements(); ++i) out.data()[i] = adjust_quant_index(qIndices.data()[i], qMatrix);
  return out;
}

// dead-zone quantiser: the magnitude, in quarter units, divided by the quantisation factor; the sign is kept
const int quant(int value, int q) {
  const int magnitude = (value < 0 ? -value : value) << 2;
  const int level = magnitude / factor(q);
  return value < 0 ? -level : level;
}

// reconstruction: level x factor, plus the offset for non-zero levels, rounded from quarter units
const int scale(int value, int q) {
  int magnitude = (value < 0 ? -value : value) * factor(q);
  if (magnitude > 0) magnitude += vc2_quant_offset(q < 0 ? 0 : q);
  magnitude = (magnitude + 2) / 4;
  return value < 0 ? -magnitude : magnitude;
}

// mean of the three causal neighbours, rounded to nearest away from zero; first row / column: the one neighbour; corner: 0
const int predictDC(const Array2D& llSubband, int y, int x) {
  if (y > 0 && x > 0) {
    const int sum = llSubband[y - 1][x - 1] + llSubband[y - 1][x] + llSubband[y][x - 1];
    return sum >= 0 ? (sum + 1) / 3 : (sum - 1) / 3;
  }
  if (y > 0) return llSubband[y - 1][x];
  if (x > 0) return llSubband[y][x - 1];
  return 0;
}

// ---- whole transforms: GPU ------------------------------------------------------------------------------------

const Array2D quantise_transform(const Array2D& coefficients, const Array2D& qIndices, const Array1D& qMatrix) {
  return run(vc2_quantise_ld, coefficients, qIndices, qMatrix);
}
const Array2D inverse_quantise_transform(const Array2D& qCoeffs, const Array2D& qIndices, const Array1D& qMatrix) {
  return run(vc2_dequantise_ld, qCoeffs, qIndices, qMatrix);
}
const Array2D quantise_transform_np(const Array2D& coefficients, const int qIndex, const Array1D& qMatrix) {
  return run(vc2_quantise_np, coefficients, uniform(qIndex), qMatrix);
}
const Array2D quantise_transform_np(const Array2D& coefficients, const Array2D& qIndices, const Array1D& qMatrix) {
  return run(vc2_quantise_np, coefficients, qIndices, qMatrix);
}
const Array2D inverse_quantise_transform_np(const Array2D& qCoeffs, const Array2D& qIndices, const Array1D& qMatrix) {
  return run(vc2_dequantise_np, qCoeffs, qIndices, qMatrix);
}

const Picture quantise_transform(const Picture& coefficients, const int qIndex, const Array1D& qMatrix) {
  return perPlane(vc2_quantise_ld, coefficients, uniform(qIndex), qMatrix);
}
const Picture quantise_transform(const Picture& coefficients, const Array2D& qIndices, const Array1D& qMatrix) {
  return perPlane(vc2_quantise_ld, coefficients, qIndices, qMatrix);
}
const Picture inverse_quantise_transform(const Picture& qCoeffs, const int qIndex, const Array1D& qMatrix) {
  return perPlane(vc2_dequantise_ld, qCoeffs, uniform(qIndex), qMatrix);
}
const Picture inverse_quantise_transform(const Picture& qCoeffs, const Array2D& qIndices, const Array1D& qMatrix) {
  return perPlane(vc2_dequantise_ld, qCoeffs, qIndices, qMatrix);
}
const Picture quantise_transform_np(const Picture& coefficients, const int qIndex, const Array1D& qMatrix) {
  return perPlane(vc2_quantise_np, coefficients, uniform(qIndex), qMatrix);
}
const Picture quantise_transform_np(const Picture& coefficients, const Array2D& qIndices, const Array1D& qMatrix) {
  return perPlane(vc2_quantise_np, coefficients, qIndices, qMatrix);
}
const Picture inverse_quantise_transform_np(const Picture& qCoeffs, const int qIndex, const Array1D& qMatrix) {
  return perPlane(vc2_dequantise_np, qCoeffs, uniform(qIndex), qMatrix);
}
const Picture inverse_quantise_transform_np(const Picture& qCoeffs, const Array2D& qIndices, const Array1D& qMatrix) {
  return perPlane(vc2_dequantise_np, qCoeffs, qIndices, qMatrix);
}

// luma squared error of one trial index: quantise and restore on the GPU, the reduction over the slice here
const long long yss_for_slice(const Picture& inPicture, const int qIndex, const Array1D& qMatrix) {
  const Array2D q = uniform(qIndex);
  const Array2D restored = run(vc2_dequantise_np, run(vc2_quantise_np, inPicture.y(), q, qMatrix), q, qMatrix);
  long long sum = 0;
  const int* a = inPicture.y().data();
  const int* b = restored.data();
  for (size_t i = 0; i < restored.num_elements(); ++i) {
    const int d = a[i] - b[i];
    sum += d * d;   // int product, as std::multiplies<int> in the reference
  }
  return sum;
}
