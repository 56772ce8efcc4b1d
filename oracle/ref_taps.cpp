// extern "C" taps onto the UNMODIFIED bbc/vc2-reference Library, linked with the
// reference's own object files into oracle/_ref/libvc2ref.so by build_ref.sh.
//
// TEST INFRASTRUCTURE ONLY (parity oracle).  Nothing in the product path may
// load this.  Every function just marshals raw int32 buffers into the
// reference's Array2D/Picture types, calls the reference function named in the
// comment, and copies the result back.  Return 0 on success, -1 when the
// reference threw (message retrievable with ref_last_error()).
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>

#include "Arrays.h"
#include "Picture.h"
#include "Quantisation.h"
#include "Slices.h"
#include "Utils.h"
#include "VLC.h"
#include "WaveletTransform.h"

// defined (non-static, undeclared in the header) in the reference's Quantisation.cpp:40-83
const int quant_factor(int q);
const int quant_offset(int q);
// defined in the reference's EncodeStream.cpp (compiled with -Dmain=... below)
const Array2D quantIndicesCBR(const Picture& coefficients, const Array1D& qMatrix,
                              const Array2D& sliceBytes, const int scalar);
const Array2D quantIndicesLD(const Picture& coefficients, const Array1D& qMatrix, const Array2D& sliceBytes);

namespace {
std::string g_err;

Array2D to_array(const int* p, int h, int w) {
  Array2D a(extents[h][w]);
  if (h * w) std::memcpy(a.data(), p, sizeof(int) * size_t(h) * size_t(w));
  return a;
}
void from_array(const Array2D& a, int* p) {
  if (a.num_elements()) std::memcpy(p, a.data(), sizeof(int) * a.num_elements());
}
Array1D to_array1(const int* p, int n) {
  Array1D a(extents[n]);
  for (int i = 0; i < n; ++i) a[i] = p[i];
  return a;
}
ColourFormat cf_of(int lh, int lw, int ch, int cw) {
  if (ch == lh && cw == lw) return CF444;
  if (ch == lh) return CF422;
  return CF420;
}
Picture to_picture(const int* y, const int* u, const int* v, int lh, int lw, int ch, int cw) {
  PictureFormat f(lh, lw, ch, cw, cf_of(lh, lw, ch, cw));
  return Picture(f, to_array(y, lh, lw), to_array(u, ch, cw), to_array(v, ch, cw));
}
}  // namespace

#define TAP_TRY try {
#define TAP_CATCH                      \
  }                                    \
  catch (const std::exception& e) {    \
    g_err = e.what();                  \
    return -1;                         \
  }                                    \
  return 0;

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// WaveletTransform.cpp:74-77, 116-136
int ref_padded_size(int size, int depth) { return paddedSize(size, depth); }
int ref_slice_size_is_valid(int depth, int luma, int chroma, int n) {
  return sliceSizeIsValid(depth, luma, chroma, n);
}

// Quantisation.cpp:40-66, 78-83
int ref_quant_factor(int q) { try { return quant_factor(q); } catch (...) { return -1; } }
int ref_quant_offset(int q) { try { return quant_offset(q); } catch (...) { return -1; } }

// WaveletTransform.cpp:345-423
int ref_quant_matrix(int kernel, int depth, int* out) {
  TAP_TRY
  const Array1D m = quantMatrix(static_cast<WaveletKernel>(kernel), depth);
  for (unsigned i = 0; i < m.size(); ++i) out[i] = m[i];
  TAP_CATCH
}

// Quantisation.cpp:69-95
int ref_quant(int value, int q, int* out) {
  TAP_TRY
  *out = quant(value, q);
  TAP_CATCH
}
int ref_scale(int value, int q, int* out) {
  TAP_TRY
  *out = scale(value, q);
  TAP_CATCH
}

// WaveletTransform.cpp:262-281 ; dst is paddedSize(h) x paddedSize(w)
int ref_dwt_forward(const int* src, int h, int w, int kernel, int depth, int* dst) {
  TAP_TRY
  from_array(waveletTransform(to_array(src, h, w), static_cast<WaveletKernel>(kernel), depth), dst);
  TAP_CATCH
}

// WaveletTransform.cpp:321-342 ; src is ph x pw, dst is h x w
int ref_dwt_inverse(const int* src, int ph, int pw, int kernel, int depth, int* dst, int h, int w) {
  TAP_TRY
  const Shape2D shape = {{h, w}};
  from_array(inverseWaveletTransform(to_array(src, ph, pw), static_cast<WaveletKernel>(kernel), depth, shape), dst);
  TAP_CATCH
}

// Quantisation.cpp:479-489 / 534-544 (per-slice indices, no DC prediction)
int ref_quantise_np(const int* coef, int ph, int pw, const int* qidx, int ny, int nx,
                    const int* qmatrix, int nbands, int* out) {
  TAP_TRY
  from_array(quantise_transform_np(to_array(coef, ph, pw), to_array(qidx, ny, nx), to_array1(qmatrix, nbands)), out);
  TAP_CATCH
}
int ref_dequantise_np(const int* coef, int ph, int pw, const int* qidx, int ny, int nx,
                      const int* qmatrix, int nbands, int* out) {
  TAP_TRY
  from_array(inverse_quantise_transform_np(to_array(coef, ph, pw), to_array(qidx, ny, nx), to_array1(qmatrix, nbands)), out);
  TAP_CATCH
}
// Quantisation.cpp:357-379 (LD: LL band DC-predicted)
int ref_quantise_ld(const int* coef, int ph, int pw, const int* qidx, int ny, int nx,
                    const int* qmatrix, int nbands, int* out) {
  TAP_TRY
  from_array(quantise_transform(to_array(coef, ph, pw), to_array(qidx, ny, nx), to_array1(qmatrix, nbands)), out);
  TAP_CATCH
}
int ref_dequantise_ld(const int* coef, int ph, int pw, const int* qidx, int ny, int nx,
                      const int* qmatrix, int nbands, int* out) {
  TAP_TRY
  from_array(inverse_quantise_transform(to_array(coef, ph, pw), to_array(qidx, ny, nx), to_array1(qmatrix, nbands)), out);
  TAP_CATCH
}

// Slices.cpp:28-49
int ref_slice_bytes(int ny, int nx, int total, int scalar, int* out) {
  TAP_TRY
  from_array(slice_bytes(ny, nx, total, scalar), out);
  TAP_CATCH
}

// Slices.cpp:97-119
int ref_component_slice_bytes(const int* slice, int h, int w, int depth, int scalar, int* out) {
  TAP_TRY
  *out = component_slice_bytes(to_array(slice, h, w), depth, scalar);
  TAP_CATCH
}

// Slices.cpp:51-96: luma_slice_bits of one slice (v == NULL) or chroma_slice_bits of a U/V slice pair
int ref_slice_bits(const int* u, const int* v, int h, int w, int depth, int* out) {
  TAP_TRY
  *out = v ? chroma_slice_bits(to_array(u, h, w), to_array(v, h, w), depth) : luma_slice_bits(to_array(u, h, w), depth);
  TAP_CATCH
}

// Quantisation.cpp:627-642
int ref_yss_for_slice(const int* y, const int* u, const int* v, int lh, int lw, int ch, int cw,
                      int q, const int* qmatrix, int nbands, long long* out) {
  TAP_TRY
  *out = yss_for_slice(to_picture(y, u, v, lh, lw, ch, cw), q, to_array1(qmatrix, nbands));
  TAP_CATCH
}

// EncodeStream.cpp:73-125
int ref_cbr_qindices(const int* y, const int* u, const int* v, int lh, int lw, int ch, int cw,
                     const int* qmatrix, int nbands, const int* slice_bytes_, int ny, int nx,
                     int scalar, int* out) {
  TAP_TRY
  from_array(quantIndicesCBR(to_picture(y, u, v, lh, lw, ch, cw), to_array1(qmatrix, nbands),
                             to_array(slice_bytes_, ny, nx), scalar), out);
  TAP_CATCH
}

// EncodeStream.cpp:193-245
int ref_ld_qindices(const int* y, const int* u, const int* v, int lh, int lw, int ch, int cw,
                    const int* qmatrix, int nbands, const int* slice_bytes_, int ny, int nx, int* out) {
  TAP_TRY
  from_array(quantIndicesLD(to_picture(y, u, v, lh, lw, ch, cw), to_array1(qmatrix, nbands), to_array(slice_bytes_, ny, nx)), out);
  TAP_CATCH
}

// Slices.cpp:645-660 with the HQ VBR (mode 0) / HQ CBR (mode 1) / LD (mode 2) slice writers
// planes are QUANTISED padded coefficient planes
int ref_pack_slices(const int* y, const int* u, const int* v, int lh, int lw, int ch, int cw,
                    int depth, const int* qidx, int ny, int nx, int mode, int prefix, int scalar,
                    const int* slice_bytes_, unsigned char* out, long cap, long* out_len) {
  TAP_TRY
  const Picture q = to_picture(y, u, v, lh, lw, ch, cw);
  const PictureArray slices = split_into_blocks(q, ny, nx);
  const Array2D qIndices = to_array(qidx, ny, nx);
  const Slices outSlices(slices, depth, qIndices);
  Array2D bytes;
  if (slice_bytes_) bytes = to_array(slice_bytes_, ny, nx);
  std::ostringstream ss;
  if (mode == 0) ss << sliceio::highQualityVBR(prefix, scalar);
  else if (mode == 1) ss << sliceio::highQualityCBR(bytes, prefix, scalar);
  else ss << sliceio::lowDelay(bytes);
  ss << outSlices;
  const std::string s = ss.str();
  *out_len = static_cast<long>(s.size());
  if (static_cast<long>(s.size()) > cap) throw std::length_error("ref_pack_slices: output buffer too small");
  std::memcpy(out, s.data(), s.size());
  TAP_CATCH
}

// Slices.cpp:662-694 with the HQ VBR reader (mode 0) or LD reader (mode 2); output = quantised planes + qidx
int ref_unpack_slices(const unsigned char* in, long len, int lh, int lw, int ch, int cw,
                      int depth, int ny, int nx, int mode, int prefix, int scalar,
                      const int* slice_bytes_, int* y, int* u, int* v, int* qidx) {
  TAP_TRY
  PictureFormat f(lh, lw, ch, cw, cf_of(lh, lw, ch, cw));
  Slices inSlices(f, depth, ny, nx);
  Array2D bytes;
  if (slice_bytes_) bytes = to_array(slice_bytes_, ny, nx);
  std::istringstream ss(std::string(reinterpret_cast<const char*>(in), static_cast<size_t>(len)));
  if (mode == 0) ss >> sliceio::highQualityVBR(prefix, scalar);
  else if (mode == 1) ss >> sliceio::highQualityCBR(bytes, prefix, scalar);
  else ss >> sliceio::lowDelay(bytes);
  ss >> inSlices;
  const Picture p = merge_blocks(inSlices.yuvSlices);
  from_array(p.y(), y);
  from_array(p.c1(), u);
  from_array(p.c2(), v);
  from_array(inSlices.qIndices, qidx);
  TAP_CATCH
}

// VLC.cpp:78-94 : code and length of one signed interleaved exp-Golomb value
int ref_signed_vlc(int value, unsigned* nbits, unsigned* code) {
  TAP_TRY
  SignedVLC c(value);
  *nbits = c.numOfBits();
  *code = c.code();
  TAP_CATCH
}

}  // extern "C"
