// Drop-in body for src/Library/src/Slices.cpp of bbc/vc2-reference: the declarations of the reference's own Slices.h,
// implemented over the CUDA C-ABI (include/vc2_cabi.h).  See vc2_dropin.h.
//   operator<<(ostream&, Slices / Slice)     Slices.cpp:195-244, 305-382, 469-533, 645-660, 697-713 -> vc2_hq_pack / vc2_ld_pack
//   operator>>(istream&, Slices)             :246-303, 535-612, 662-694                             -> vc2_hq_unpack / vc2_ld_unpack
//   luma_slice_bits / chroma_slice_bits      :51-96                                                 -> vc2_slice_bits
//   component_slice_bytes                    :97-119                                                -> vc2_hq_slice_sizes
//   slice_bytes                              :18-49                                                 -> vc2_slice_bytes / host arithmetic
//   sliceio:: manipulators, Slices, SliceQuantiser bookkeeping  :120-146, 148-193, 613-644, 715-end (stream state, host)
// The coding mode travels in the stream's iword slots exactly as in the reference, so `ss.copyfmt(stream)` in
// DataUnit.cpp carries it over to the temporary streams the wrapped-picture writers use.
#include <iostream>
#include <string>
#include <vector>
#include "Slices.h"
#include "DataUnit.h"
#include "Utils.h"
#include "vc2_dropin.h"

using vc2dropin::check;
using vc2dropin::ctx;

// ---- slice sizes ------------------------------------------------------------------------------------------

const int slice_bytes(int v, int h, const int ySlices, const int xSlices, const int sliceBytesNumerator,
                      const int sliceBytesDenominator) {
  // the bytes up to and including this slice minus the bytes up to the one before, both rounded down
  const long long index = (long long)v * xSlices + h;
  return (int)((index + 1) * sliceBytesNumerator / sliceBytesDenominator - index * sliceBytesNumerator / sliceBytesDenominator);
}

const Array2D slice_bytes(const int ySlices, const int xSlices, const int totalBytes, const int scalar) {
  Array2D bytes(extents[ySlices][xSlices]);
  check(vc2_slice_bytes(ySlices, xSlices, totalBytes, scalar, bytes.data()));
  return bytes;
}

const int luma_slice_bits(const Array2D& lumaSlice, const char waveletDepth) {
  int bits = 0;
  check(vc2_slice_bits(ctx(), lumaSlice.data(), nullptr, (int)lumaSlice.shape()[0], (int)lumaSlice.shape()[1], waveletDepth, 1, 1, &bits));
  return bits;
}

const int chroma_slice_bits(const Array2D& uSlice, const Array2D& vSlice, const char waveletDepth) {
  int bits = 0;
  check(vc2_slice_bits(ctx(), uSlice.data(), vSlice.data(), (int)uSlice.shape()[0], (int)uSlice.shape()[1], waveletDepth, 1, 1, &bits));
  return bits;
}

const int component_slice_bytes(const Array2D& componentSlice, const char waveletDepth, const int scalar) {
  int bytes = 0;
  check(vc2_hq_slice_sizes(ctx(), componentSlice.data(), (int)componentSlice.shape()[0], (int)componentSlice.shape()[1], waveletDepth,
                           1, 1, scalar, &bytes));
  return bytes;
}

// ---- SliceQuantiser: raster walk over the slices of a component (the LD rate control derives from it) -----------

SliceQuantiser::SliceQuantiser(const Array2D& coefficients, int vSlices, int hSlices, const Array1D& quantMatrix)
    : ySlices(vSlices), xSlices(hSlices),
      coeffsHeight((int)coefficients.shape()[0]), coeffsWidth((int)coefficients.shape()[1]),
      sliceHeight(coeffsHeight / vSlices), sliceWidth(coeffsWidth / hSlices),
      numberOfSubbands((int)quantMatrix.size()), waveletDepth((numberOfSubbands - 1) / 3),
      transformSize(1 << waveletDepth), v(0), h(0) {
  qSlice.resize(extents[sliceHeight][sliceWidth]);
}

const bool SliceQuantiser::next_slice() {
  if (h + 1 < xSlices) { ++h; return true; }
  if (v + 1 < ySlices) { h = 0; ++v; return true; }
  return false;
}

// ---- stream state ----------------------------------------------------------------------------------------

namespace {
// one iword slot per item, allocated on first use
enum Item { MODE, SIZES, SCALAR, PREFIX, ONE_SIZE, COUNT, OFFSET_X, OFFSET_Y, ITEMS };
long& item(std::ios_base& stream, Item which) {
  static int slot[ITEMS];
  static bool ready = false;
  if (!ready) {
    for (int i = 0; i < ITEMS; ++i) slot[i] = std::ios_base::xalloc();
    ready = true;
  }
  return stream.iword(slot[which]);
}

struct Geometry {
  vc2_geom g;
  int lh, lw, ch, cw;   // plane sizes of the whole block of slices
};
// geometry of ny x nx slices of the given slice format laid out as one picture
Geometry geometry(const PictureFormat& slice, int depth, int ny, int nx, int prefix, int scalar) {
  Geometry r;
  r.lh = slice.lumaHeight() * ny; r.lw = slice.lumaWidth() * nx;
  r.ch = slice.chromaHeight() * ny; r.cw = slice.chromaWidth() * nx;
  r.g.luma_h = r.lh; r.g.luma_w = r.lw; r.g.chroma_h = r.ch; r.g.chroma_w = r.cw;
  r.g.kernel = 0;   // the slice syntax does not depend on the wavelet kernel
  r.g.depth = depth; r.g.slices_y = ny; r.g.slices_x = nx; r.g.prefix = prefix; r.g.scalar = scalar;
  return r;
}

sliceio::SliceIOMode modeOf(std::ios_base& stream, const char* direction) {
  const long m = item(stream, MODE);
  if (!m) throw std::logic_error(std::string("SliceIO: ") + direction + " Format not set");
  if (m != sliceio::LD && m != sliceio::HQVBR && m != sliceio::HQCBR)
    throw std::logic_error(std::string("SliceIO: Unknown ") + direction + " Format");
  return static_cast<sliceio::SliceIOMode>(m);
}

// code ny x nx slices held as whole planes and append the bytes to the stream
void writePlanes(std::ostream& stream, const Picture& q, int depth, int ny, int nx, const Array2D& qIndices, const int* sizes) {
  const sliceio::SliceIOMode mode = modeOf(stream, "Output");
  const int prefix = mode == sliceio::LD ? 0 : (int)item(stream, PREFIX), scalar = mode == sliceio::LD ? 1 : (int)item(stream, SCALAR);
  const PictureFormat f = q.format();
  const PictureFormat sliceFormat(f.lumaHeight() / ny, f.lumaWidth() / nx, f.chromaHeight() / ny, f.chromaWidth() / nx, f.chromaFormat());
  const Geometry geo = geometry(sliceFormat, depth, ny, nx, prefix, scalar);
  const int n = ny * nx;
  size_t cap = 0;
  if (mode == sliceio::HQVBR) {
    cap = 4 * (q.y().num_elements() + q.c1().num_elements() + q.c2().num_elements()) + (size_t)(prefix + 4) * n + 1024;
  } else {
    if (!sizes) throw std::logic_error("SliceIO: slice sizes not set");
    for (int i = 0; i < n; ++i) cap += (size_t)sizes[i] + prefix;
    cap += 1024;
  }
  std::vector<uint8_t> out(cap);
  size_t len = 0;
  if (mode == sliceio::LD)
    check(vc2_ld_pack(ctx(), q.y().data(), q.c1().data(), q.c2().data(), &geo.g, qIndices.data(), sizes, out.data(), cap, &len));
  else
    check(vc2_hq_pack(ctx(), q.y().data(), q.c1().data(), q.c2().data(), &geo.g, qIndices.data(),
                      mode == sliceio::HQCBR ? VC2_HQ_CBR : VC2_HQ_VBR, mode == sliceio::HQCBR ? sizes : nullptr, out.data(), cap, &len,
                      nullptr));
  stream.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)len);
}

void readBytes(std::istream& stream, std::vector<uint8_t>& buf, size_t n) {
  const size_t at = buf.size();
  buf.resize(at + n);
  if (n && !stream.read(reinterpret_cast<char*>(&buf[at]), (std::streamsize)n)) throw std::logic_error("SliceIO: stream ends inside a slice");
}
}  // namespace

sliceio::SliceIOMode& sliceio::sliceIOMode(std::ios_base& stream) {
  return reinterpret_cast<sliceio::SliceIOMode&>(item(stream, MODE));
}

const Array2D* sliceio::SliceSizes(std::ios_base& stream) { return reinterpret_cast<const Array2D*>(item(stream, SIZES)); }

// ---- Slices -------------------------------------------------------------------------------------------------

Slices::Slices(const PictureArray& s, const int d, const Array2D& i) : yuvSlices(s), waveletDepth(d), qIndices(i) {}

Slices::Slices(const PictureFormat& pictureFormat, int d, int ySlices, int xSlices) : waveletDepth(d) {
  const PictureFormat sliceFormat(pictureFormat.lumaHeight() / ySlices, pictureFormat.lumaWidth() / xSlices,
                                  pictureFormat.chromaHeight() / ySlices, pictureFormat.chromaWidth() / xSlices,
                                  pictureFormat.chromaFormat());
  const Shape2D shape = {{ySlices, xSlices}};
  yuvSlices = PictureArray(shape);
  for (int v = 0; v < ySlices; ++v)
    for (int h = 0; h < xSlices; ++h) yuvSlices[v][h] = Picture(sliceFormat);
  qIndices = Array2D(shape);
}

std::ostream& operator<<(std::ostream& stream, const Slices& s) {
  const int ny = (int)s.yuvSlices.shape()[0], nx = (int)s.yuvSlices.shape()[1];
  const Array2D* sizes = sliceio::SliceSizes(stream);
  writePlanes(stream, merge_blocks(s.yuvSlices), s.waveletDepth, ny, nx, s.qIndices, sizes ? sizes->data() : nullptr);
  return stream;
}

std::ostream& operator<<(std::ostream& stream, const Slice& s) {
  Array2D q(extents[1][1]);
  q[0][0] = s.qIndex;
  const int size = (int)item(stream, ONE_SIZE);
  writePlanes(stream, s.yuvSlice, s.waveletDepth, 1, 1, q, &size);
  return stream;
}

// Reads the whole picture, or - after sliceio::ExpectedSlicesForFragment - the run of slices a fragment carries
std::istream& operator>>(std::istream& stream, Slices& s) {
  const sliceio::SliceIOMode mode = modeOf(stream, "Input");
  const int ny = (int)s.yuvSlices.shape()[0], nx = (int)s.yuvSlices.shape()[1];
  const int prefix = mode == sliceio::LD ? 0 : (int)item(stream, PREFIX), scalar = mode == sliceio::LD ? 1 : (int)item(stream, SCALAR);
  const int first = (int)item(stream, OFFSET_Y) * nx + (int)item(stream, OFFSET_X);
  const int expected = (int)item(stream, COUNT);
  const int count = expected ? std::min(expected, ny * nx - first) : ny * nx - first;
  if (count <= 0) return stream;
  // the bytes of the run: LD slices have their sizes from the manipulator, HQ slices carry three length bytes
  std::vector<uint8_t> buf;
  std::vector<int> sizes(count);
  const Array2D* given = sliceio::SliceSizes(stream);
  for (int i = 0; i < count; ++i) {
    const size_t start = buf.size();
    if (mode == sliceio::LD) {
      if (!given) throw std::logic_error("SliceIO: slice sizes not set");
      readBytes(stream, buf, (size_t)given->data()[first + i]);
    } else {
      readBytes(stream, buf, (size_t)prefix + 1);
      for (int c = 0; c < 3; ++c) {
        readBytes(stream, buf, 1);
        readBytes(stream, buf, (size_t)buf.back() * scalar);
      }
    }
    sizes[i] = (int)(buf.size() - start);
  }
  // a whole picture keeps its geometry; a partial run is decoded as one row of `count` slices
  const bool whole = first == 0 && count == ny * nx;
  const int gy = whole ? ny : 1, gx = whole ? nx : count;
  const PictureFormat sliceFormat = s.yuvSlices[0][0].format();
  const Geometry geo = geometry(sliceFormat, s.waveletDepth, gy, gx, prefix, scalar);
  Array2D Y(extents[geo.lh][geo.lw]), U(extents[geo.ch][geo.cw]), V(extents[geo.ch][geo.cw]), Q(extents[gy][gx]);
  if (mode == sliceio::LD)
    check(vc2_ld_unpack(ctx(), buf.data(), buf.size(), &geo.g, sizes.data(), Y.data(), U.data(), V.data(), Q.data()));
  else
    check(vc2_hq_unpack(ctx(), buf.data(), buf.size(), &geo.g, Y.data(), U.data(), V.data(), Q.data()));
  const int sh = sliceFormat.lumaHeight(), sw = sliceFormat.lumaWidth(), csh = sliceFormat.chromaHeight(), csw = sliceFormat.chromaWidth();
  Array2D y(extents[sh][sw]), u(extents[csh][csw]), v(extents[csh][csw]);
  for (int i = 0; i < count; ++i) {
    const int sv = (first + i) / nx, shh = (first + i) % nx;   // where the slice belongs
    const int by = whole ? sv : 0, bx = whole ? shh : i;       // where it sits in the decoded block
    for (int r = 0; r < sh; ++r)
      for (int c = 0; c < sw; ++c) y[r][c] = Y[by * sh + r][bx * sw + c];
    for (int r = 0; r < csh; ++r)
      for (int c = 0; c < csw; ++c) { u[r][c] = U[by * csh + r][bx * csw + c]; v[r][c] = V[by * csh + r][bx * csw + c]; }
    s.yuvSlices[sv][shh].y(y);
    s.yuvSlices[sv][shh].c1(u);
    s.yuvSlices[sv][shh].c2(v);
    s.qIndices[sv][shh] = Q[by][bx];
  }
  return stream;
}

std::istream& operator>>(std::istream& stream, Slice& s) {
  Slices one(s.yuvSlice.format(), s.waveletDepth, 1, 1);
  Array2D size(extents[1][1]);
  size[0][0] = (int)item(stream, ONE_SIZE);
  const long keepSizes = item(stream, SIZES), keepCount = item(stream, COUNT), keepX = item(stream, OFFSET_X), keepY = item(stream, OFFSET_Y);
  item(stream, SIZES) = reinterpret_cast<long>(&size);
  item(stream, COUNT) = 0; item(stream, OFFSET_X) = 0; item(stream, OFFSET_Y) = 0;
  stream >> one;
  item(stream, SIZES) = keepSizes; item(stream, COUNT) = keepCount; item(stream, OFFSET_X) = keepX; item(stream, OFFSET_Y) = keepY;
  s.yuvSlice = one.yuvSlices[0][0];
  s.qIndex = one.qIndices[0][0];
  return stream;
}

// ---- manipulators ---------------------------------------------------------------------------------------------

sliceio::ExpectedSlicesForFragment::ExpectedSlicesForFragment(Fragment& frag)
    : n_slices(frag.n_slices()), slice_offset_x(frag.slice_offset_x()), slice_offset_y(frag.slice_offset_y()) {}

void sliceio::ExpectedSlicesForFragment::operator()(std::ios_base& stream) const {
  item(stream, COUNT) = n_slices;
  item(stream, OFFSET_X) = slice_offset_x;
  item(stream, OFFSET_Y) = slice_offset_y;
}

void sliceio::lowDelay::operator()(std::ios_base& stream) const {
  item(stream, MODE) = LD;
  item(stream, SIZES) = reinterpret_cast<long>(&bytes);
}

void sliceio::highQualityCBR::operator()(std::ios_base& stream) const {
  item(stream, MODE) = HQCBR;
  item(stream, SIZES) = reinterpret_cast<long>(&bytes);
  item(stream, PREFIX) = prefix;
  item(stream, SCALAR) = scalar;
}

void sliceio::highQualityVBR::operator()(std::ios_base& stream) const {
  item(stream, MODE) = HQVBR;
  item(stream, PREFIX) = prefix;
  item(stream, SCALAR) = scalar;
}

void sliceio::setBytes::operator()(std::ios_base& stream) const { item(stream, ONE_SIZE) = bytes; }

std::istream& operator>>(std::istream& stream, sliceio::ExpectedSlicesForFragment esf) { esf(stream); return stream; }
std::ostream& operator<<(std::ostream& stream, sliceio::setBytes arg) { arg(stream); return stream; }
std::istream& operator>>(std::istream& stream, sliceio::setBytes arg) { arg(stream); return stream; }
std::ostream& operator<<(std::ostream& stream, sliceio::lowDelay arg) { arg(stream); return stream; }
std::istream& operator>>(std::istream& stream, sliceio::lowDelay arg) { arg(stream); return stream; }
std::ostream& operator<<(std::ostream& stream, sliceio::highQualityCBR arg) { arg(stream); return stream; }
std::istream& operator>>(std::istream& stream, sliceio::highQualityCBR arg) { arg(stream); return stream; }
std::ostream& operator<<(std::ostream& stream, sliceio::highQualityVBR arg) { arg(stream); return stream; }
std::istream& operator>>(std::istream& stream, sliceio::highQualityVBR arg) { arg(stream); return stream; }
