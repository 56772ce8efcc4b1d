// vc2_library.cpp - the C++ mirror of the reference Library for the hot path (include/vc2/*.h), a thin layer
// over the C-ABI of libvc2b200.so.  Every transform / quantiser / slice coder call runs on the GPU; status
// codes come back as the reference's exception types and texts (src/Library/src/*.cpp, cited per function).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "vc2/Arrays.h"
#include "vc2/Picture.h"
#include "vc2/Quantisation.h"
#include "vc2/Slices.h"
#include "vc2/WaveletTransform.h"
#include "vc2_cabi.h"

namespace vc2 {

namespace {

// one context per thread (the reference Library is re-entrant across distinct pictures)
struct ThreadCtx {
  vc2_ctx* c = nullptr;
  ~ThreadCtx() { if (c) vc2_destroy(c); }
};
vc2_ctx* ctx() {
  thread_local ThreadCtx t;
  if (!t.c) {
    t.c = vc2_create(0);
    if (!t.c) throw std::runtime_error("vc2: no CUDA device (the hot path has no CPU fallback)");
  }
  return t.c;
}

// status -> the exception the reference throws at the corresponding site
void raise(int st) {
  if (st == VC2_OK) return;
  const std::string msg = vc2_last_error(ctx());
  switch (st) {
    case VC2_ERR_ARG: throw std::invalid_argument(msg);
    case VC2_ERR_CUDA: throw std::runtime_error(msg);
    default: throw std::logic_error(msg);   // Slices.cpp:115-117, 356-366; Quantisation.cpp:60-63; stream errors
  }
}

vc2_geom slice_geom(const PictureFormat& padded, WaveletKernel kernel, int depth, int ySlices, int xSlices, int prefix, int scalar) {
  vc2_geom g;
  g.luma_h = padded.lumaHeight(); g.luma_w = padded.lumaWidth();
  g.chroma_h = padded.chromaHeight(); g.chroma_w = padded.chromaWidth();
  g.kernel = (int)kernel; g.depth = depth; g.slices_y = ySlices; g.slices_x = xSlices; g.prefix = prefix; g.scalar = scalar;
  return g;
}

}  // namespace

// ---- Arrays ----------------------------------------------------------------------------------------
const Array2D clip(const Array2D& v, int lo, int hi) {
  Array2D r(v.shape()[0], v.shape()[1]);
  const int* s = v.data();
  int* d = r.data();
  for (size_t i = 0; i < v.num_elements(); ++i) d[i] = s[i] < lo ? lo : s[i] > hi ? hi : s[i];
  return r;
}

bool readArray(std::istream& in, Array2D& a, const SampleFormat& f) {
  const size_t n = a.num_elements();
  std::vector<unsigned char> buf(n * f.bytes);
  in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
  if ((size_t)in.gcount() != buf.size()) { in.setstate(std::ios_base::failbit); return false; }
  const int shift = f.left_justified ? 8 * f.bytes - f.depth : 0;
  const long long offset = f.offset_binary ? (1ll << (f.depth - 1)) : 0;
  int* d = a.data();
  for (size_t i = 0; i < n; ++i) {
    unsigned long long w = 0;
    for (int b = 0; b < f.bytes; ++b) w = (w << 8) | buf[i * f.bytes + b];
    long long v;
    if (f.offset_binary) v = (long long)(w >> shift) - offset;
    else {   // two's complement word of 8*bytes bits (Arrays.cpp:360-366)
      const int bits = 8 * f.bytes;
      v = (long long)w;
      if (bits < 64 && (w >> (bits - 1))) v -= (1ll << bits);
      v >>= shift;
    }
    d[i] = (int)v;
  }
  return true;
}

bool writeArray(std::ostream& out, const Array2D& a, const SampleFormat& f) {
  const size_t n = a.num_elements();
  std::vector<unsigned char> buf(n * f.bytes);
  const int shift = f.left_justified ? 8 * f.bytes - f.depth : 0;
  const long long offset = f.offset_binary ? (1ll << (f.depth - 1)) : 0;
  const int* s = a.data();
  for (size_t i = 0; i < n; ++i) {
    const unsigned long long w = (unsigned long long)(((long long)s[i] + offset) << shift);   // Arrays.cpp:396-397
    for (int b = 0; b < f.bytes; ++b) buf[i * f.bytes + b] = (unsigned char)(w >> (8 * (f.bytes - 1 - b)));
  }
  out.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)buf.size());
  return (bool)out;
}

// ---- Picture ---------------------------------------------------------------------------------------
PictureFormat::PictureFormat(int height, int width, ColourFormat cf) : h_(height), w_(width), cf_(cf) {
  if (cf != CF444 && cf != CF422 && cf != CF420 && cf != CF_UNSET) throw std::invalid_argument("Invalid colour format");
}

const Picture clip(const Picture& p, int ylo, int yhi, int clo, int chi) {
  return Picture(p.format(), clip(p.y(), ylo, yhi), clip(p.c1(), clo, chi), clip(p.c2(), clo, chi));
}

bool readPicture(std::istream& in, Picture& p, int bytes, int ld, int cd, bool ob) {
  const SampleFormat fy = {bytes, ld, true, ob}, fc = {bytes, cd, true, ob};
  return readArray(in, p.y(), fy) && readArray(in, p.c1(), fc) && readArray(in, p.c2(), fc);
}
bool writePicture(std::ostream& out, const Picture& p, int bytes, int ld, int cd, bool ob) {
  const SampleFormat fy = {bytes, ld, true, ob}, fc = {bytes, cd, true, ob};
  return writeArray(out, p.y(), fy) && writeArray(out, p.c1(), fc) && writeArray(out, p.c2(), fc);
}

// ---- WaveletTransform ----------------------------------------------------------------------------------
std::ostream& operator<<(std::ostream& os, WaveletKernel k) {
  static const char* names[] = {"Deslauriers-Dubuc (9,7) (\"DD97\")", "LeGall (5,3) (\"LeGall\")", "Deslauriers-Dubuc (13,7) (\"DD137\")",
                                "Haar (no shift) (\"Haar0\")", "Haar (one bit shift) (\"Haar1\")", "Fidelity (\"Fidelity\")",
                                "Daubechies (9,7) (\"Daub97\")", "NullKernel (\"NullKernel\")"};
  return os << ((int)k >= 0 && (int)k <= 7 ? names[(int)k] : "Unknown wavelet kernel!");
}
std::istream& operator>>(std::istream& is, WaveletKernel& k) {
  static const char* names[] = {"DD97", "LeGall", "DD137", "Haar0", "Haar1", "Fidelity", "Daub97", "NullKernel"};
  std::string t;
  is >> t;
  for (int i = 0; i < 8; ++i) if (t == names[i]) { k = (WaveletKernel)i; return is; }
  throw std::invalid_argument("invalid wavelet kernel");
}

int paddedSize(int size, int depth) { return vc2_padded_size(size, depth); }
int sliceSizeIsValid(int depth, int luma, int chroma, int n) { return vc2_slice_size_is_valid(depth, luma, chroma, n); }

static int gcd_i(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

bool waveletTransformIsPossible(int depth, int luma, int chroma) {   // WaveletTransform.cpp:96-108
  if (depth <= 0 || depth > 31) return false;
  const int g = gcd_i(paddedSize(luma, depth), paddedSize(chroma, depth));
  return g / (1 << depth) >= 2;
}
int suggestSliceSize(int depth, int luma, int chroma, int start) {   // :197-210
  const int maxSlices = std::min(luma, chroma) / (1 << depth);
  if (start > maxSlices) start = maxSlices;
  for (int n = 0, sgn = 1; n < 2 * maxSlices; ++n, sgn = -sgn) {
    const int t = start + sgn * (n + 1) / 2;
    if (sliceSizeIsValid(depth, luma, chroma, t)) return t;
  }
  throw std::logic_error("It is not possible to encode this picture because of its dimensions.");
}
int suggestWaveletDepth(int lw, int lh, int cw, int ch, int start) {   // :160-178
  const int minDim = std::min(std::min(lh, lw), std::min(ch, cw));
  const double lg = std::log2((double)minDim);
  if (start > lg) start = (int)lg;
  for (int n = 1, sgn = -1; n < 2 * lg; ++n, sgn = -sgn) {
    const int d = start + sgn * (n + 1) / 2;
    if (waveletTransformIsPossible(d, lw, cw) && waveletTransformIsPossible(d, lh, ch)) return d;
  }
  throw std::logic_error("It is not possible to encode this picture because of its dimensions.");
}

const Array2D waveletTransform(const Array2D& picture, WaveletKernel kernel, int depth) {
  const int h = (int)picture.shape()[0], w = (int)picture.shape()[1];
  Array2D out(paddedSize(h, depth), paddedSize(w, depth));
  raise(vc2_dwt_forward(ctx(), picture.data(), h, w, (int)kernel, depth, out.data()));
  return out;
}
const Picture waveletTransform(const Picture& p, WaveletKernel kernel, int depth) {
  const PictureFormat& f = p.format();
  const PictureFormat tf(paddedSize(f.lumaHeight(), depth), paddedSize(f.lumaWidth(), depth), f.chromaFormat());
  // the chroma planes pad on their own size, exactly as three Array2D transforms do (WaveletTransform.cpp:283-290)
  return Picture(tf, waveletTransform(p.y(), kernel, depth), waveletTransform(p.c1(), kernel, depth), waveletTransform(p.c2(), kernel, depth));
}
const Array2D inverseWaveletTransform(const Array2D& t, WaveletKernel kernel, int depth, int height, int width) {
  Array2D out(height, width);
  raise(vc2_dwt_inverse(ctx(), t.data(), (int)t.shape()[0], (int)t.shape()[1], (int)kernel, depth, out.data(), height, width));
  return out;
}
const Picture inverseWaveletTransform(const Picture& t, WaveletKernel kernel, int depth, const PictureFormat& f) {
  return Picture(f, inverseWaveletTransform(t.y(), kernel, depth, f.lumaHeight(), f.lumaWidth()),
                 inverseWaveletTransform(t.c1(), kernel, depth, f.chromaHeight(), f.chromaWidth()),
                 inverseWaveletTransform(t.c2(), kernel, depth, f.chromaHeight(), f.chromaWidth()));
}
const Array1D quantMatrix(WaveletKernel kernel, int depth) {
  if (depth < 0) throw std::domain_error("wavelet depth may not be < 0");   // WaveletTransform.cpp:348 (any depth >= 0 is computed)
  Array1D m(3 * depth + 1);
  if (vc2_quant_matrix((int)kernel, depth, m.data()) != VC2_OK) throw std::invalid_argument("quantMatrix: invalid kernel or depth");
  return m;
}

// ---- Quantisation --------------------------------------------------------------------------------------
namespace {
int depth_of(const Array1D& qMatrix) { return ((int)qMatrix.size() - 1) / 3; }
typedef int (*quant_fn)(vc2_ctx*, const int32_t*, int, int, int, const int32_t*, const int32_t*, int, int, int32_t*);
Array2D quant_plane(quant_fn fn, const Array2D& c, const Array2D& q, const Array1D& m) {
  Array2D out(c.shape()[0], c.shape()[1]);
  raise(fn(ctx(), c.data(), (int)c.shape()[0], (int)c.shape()[1], depth_of(m), m.data(), q.data(), (int)q.shape()[0], (int)q.shape()[1],
           out.data()));
  return out;
}
Picture quant_picture(quant_fn fn, const Picture& p, const Array2D& q, const Array1D& m) {
  return Picture(p.format(), quant_plane(fn, p.y(), q, m), quant_plane(fn, p.c1(), q, m), quant_plane(fn, p.c2(), q, m));
}
Array2D one_index(int q) { Array2D a(1, 1); a[0][0] = q; return a; }
}  // namespace

const Array2D quantise_transform_np(const Array2D& c, const Array2D& q, const Array1D& m) { return quant_plane(vc2_quantise_np, c, q, m); }
const Picture quantise_transform_np(const Picture& c, const Array2D& q, const Array1D& m) { return quant_picture(vc2_quantise_np, c, q, m); }
const Picture quantise_transform_np(const Picture& c, int q, const Array1D& m) { return quant_picture(vc2_quantise_np, c, one_index(q), m); }
const Array2D inverse_quantise_transform_np(const Array2D& c, const Array2D& q, const Array1D& m) { return quant_plane(vc2_dequantise_np, c, q, m); }
const Picture inverse_quantise_transform_np(const Picture& c, const Array2D& q, const Array1D& m) { return quant_picture(vc2_dequantise_np, c, q, m); }
const Picture inverse_quantise_transform_np(const Picture& c, int q, const Array1D& m) { return quant_picture(vc2_dequantise_np, c, one_index(q), m); }
const Picture inverse_quantise_transform(const Picture& c, const Array2D& q, const Array1D& m) { return quant_picture(vc2_dequantise_ld, c, q, m); }

const Array2D quantIndicesConstQ(int ySlices, int xSlices, int qIndex) {
  Array2D q(ySlices, xSlices);
  for (size_t i = 0; i < q.num_elements(); ++i) q.data()[i] = qIndex;
  return q;
}

const Array2D quantIndicesCBR(const Picture& c, const Array1D& qMatrix, const Array2D& sliceBytes, int scalar, int depth) {
  const int ny = (int)sliceBytes.shape()[0], nx = (int)sliceBytes.shape()[1];
  const vc2_geom g = slice_geom(c.format(), DD97 /* unused by the search */, depth, ny, nx, 0, scalar);
  Array2D q(ny, nx);
  raise(vc2_cbr_qindices(ctx(), c.y().data(), c.c1().data(), c.c2().data(), &g, qMatrix.data(), sliceBytes.data(), q.data(), nullptr));
  return q;
}

// ---- Slices --------------------------------------------------------------------------------------------
const Array2D slice_bytes(int ny, int nx, int total, int scalar) {
  Array2D a(ny, nx);
  if (vc2_slice_bytes(ny, nx, total, scalar, a.data()) != VC2_OK) throw std::invalid_argument("slice_bytes: invalid arguments");
  return a;
}

namespace {
std::string write_slices(const Slices& s, WaveletKernel kernel, int mode, const Array2D* bytes, int prefix, int scalar) {
  const int ny = (int)s.qIndices.shape()[0], nx = (int)s.qIndices.shape()[1];
  const vc2_geom g = slice_geom(s.yuvCoeffs.format(), kernel, s.waveletDepth, ny, nx, prefix, scalar);
  size_t cap = (size_t)ny * nx * (size_t)(prefix + 4 + 3 * 255 * scalar) + 64;
  if (bytes) { cap = 64; for (size_t i = 0; i < bytes->num_elements(); ++i) cap += (size_t)bytes->data()[i] + prefix; }
  std::string out(cap, '\0');
  size_t len = 0;
  raise(vc2_hq_pack(ctx(), s.yuvCoeffs.y().data(), s.yuvCoeffs.c1().data(), s.yuvCoeffs.c2().data(), &g, s.qIndices.data(), mode,
                    bytes ? bytes->data() : nullptr, reinterpret_cast<uint8_t*>(&out[0]), cap, &len, nullptr));
  out.resize(len);
  return out;
}
}  // namespace

std::string writeSlicesHQVBR(const Slices& s, WaveletKernel kernel, int prefix, int scalar) {
  return write_slices(s, kernel, VC2_HQ_VBR, nullptr, prefix, scalar);
}
std::string writeSlicesHQCBR(const Slices& s, WaveletKernel kernel, const Array2D& sliceBytes, int prefix, int scalar) {
  return write_slices(s, kernel, VC2_HQ_CBR, &sliceBytes, prefix, scalar);
}

Slices readSlicesHQ(const uint8_t* data, size_t len, const PictureFormat& tf, WaveletKernel kernel, int depth, int ny, int nx, int prefix,
                    int scalar) {
  Slices s;
  s.yuvCoeffs = Picture(tf);
  s.waveletDepth = depth;
  s.qIndices.resize(ny, nx);
  const vc2_geom g = slice_geom(tf, kernel, depth, ny, nx, prefix, scalar);
  raise(vc2_hq_unpack(ctx(), data, len, &g, s.yuvCoeffs.y().data(), s.yuvCoeffs.c1().data(), s.yuvCoeffs.c2().data(), s.qIndices.data()));
  return s;
}

Slices readSlicesLD(const uint8_t* data, size_t len, const PictureFormat& tf, WaveletKernel kernel, int depth, int ny, int nx,
                    const Array2D& sliceBytes) {
  Slices s;
  s.yuvCoeffs = Picture(tf);
  s.waveletDepth = depth;
  s.qIndices.resize(ny, nx);
  const vc2_geom g = slice_geom(tf, kernel, depth, ny, nx, 0, 1);
  raise(vc2_ld_unpack(ctx(), data, len, &g, sliceBytes.data(), s.yuvCoeffs.y().data(), s.yuvCoeffs.c1().data(), s.yuvCoeffs.c2().data(),
                      s.qIndices.data()));
  return s;
}

}  // namespace vc2
