// Drop-in body for src/Library/src/WaveletTransform.cpp of bbc/vc2-reference: the declarations of the reference's
// own WaveletTransform.h, implemented over the CUDA C-ABI (include/vc2_cabi.h).  See vc2_dropin.h.
//   waveletTransform / inverseWaveletTransform   WaveletTransform.cpp:262-281, 321-342  -> vc2_dwt_forward / vc2_dwt_inverse
//   quantMatrix                                  :345-423                               -> vc2_quant_matrix
//   paddedSize, sliceSizeIsValid                 :74-77, 116-136                        -> vc2_padded_size, vc2_slice_size_is_valid
//   split_into_subbands / merge_subbands         :428-476   (pure index permutation, host)
//   kernel names, suggest* helpers               :23-72, 96-207 (command-line conveniences, host)
#include <cmath>
#include <iostream>
#include <numeric>
#include "WaveletTransform.h"
#include "vc2_dropin.h"

using vc2dropin::check;
using vc2dropin::ctx;

namespace {
struct KernelName { WaveletKernel k; const char* tag; const char* title; };
const KernelName kNames[] = {
    {DD97, "DD97", "Deslauriers-Dubuc (9,7)"}, {LeGall, "LeGall", "LeGall (5,3)"}, {DD137, "DD137", "Deslauriers-Dubuc (13,7)"},
    {Haar0, "Haar0", "Haar (no shift)"},       {Haar1, "Haar1", "Haar (one bit shift)"}, {Fidelity, "Fidelity", "Fidelity"},
    {Daub97, "Daub97", "Daubechies (9,7)"},    {NullKernel, "NullKernel", "NullKernel"}};

int gcd_of(int a, int b) {
  while (b) { const int t = a % b; a = b; b = t; }
  return a;
}
bool bothPossible(int depth, int lw, int lh, int cw, int ch) {
  return waveletTransformIsPossible(depth, lw, cw) && waveletTransformIsPossible(depth, lh, ch);
}
[[noreturn]] void impossible() { throw std::logic_error("It is not possible to encode this picture because of its dimensions."); }
}  // namespace

std::ostream& operator<<(std::ostream& os, WaveletKernel kernel) {
  for (const KernelName& n : kNames)
    if (n.k == kernel) return os << n.title << " (\"" << n.tag << "\")";
  return os << "Unknown wavelet kernel!";
}

std::istream& operator>>(std::istream& strm, WaveletKernel& kernel) {
  std::string text;
  strm >> text;
  for (const KernelName& n : kNames)
    if (text == n.tag) { kernel = n.k; return strm; }
  throw std::invalid_argument("invalid wavelet kernel");
}

const int paddedSize(int size, int depth) { return vc2_padded_size(size, depth); }

const bool waveletTransformIsPossible(const int waveletDepth, const int lengthLuma, const int lengthChroma) {
  if (waveletDepth <= 0 || waveletDepth > 31) return false;
  // at least two slices of 2^depth must fit the common divisor of the padded lengths
  const int g = gcd_of(paddedSize(lengthLuma, waveletDepth), paddedSize(lengthChroma, waveletDepth));
  return (g >> waveletDepth) >= 2;
}

const int sliceSizeIsValid(const int waveletDepth, const int lengthLuma, const int lengthChroma, const int nSize) {
  return vc2_slice_size_is_valid(waveletDepth, lengthLuma, lengthChroma, nSize);
}

const int suggestWaveletDepth(const int lumaWidth, const int lumaHeight, const int chromaWidth, const int chromaHeight) {
  const int smallest = std::min(std::min(lumaHeight, lumaWidth), std::min(chromaHeight, chromaWidth));
  for (int depth = 1; depth < log2(smallest); ++depth)
    if (bothPossible(depth, lumaWidth, lumaHeight, chromaWidth, chromaHeight)) return depth;
  impossible();
}

const int suggestWaveletDepth(const int lumaWidth, const int lumaHeight, const int chromaWidth, const int chromaHeight,
                              int startingDepth) {
  const int smallest = std::min(std::min(lumaHeight, lumaWidth), std::min(chromaHeight, chromaWidth));
  if (startingDepth > log2(smallest)) startingDepth = log2(smallest);
  // candidates alternate below / above the starting depth: -1, +1, -2, +2 ... (the reference's search order)
  int sign = -1;
  for (int n = 1; n < 2 * log2(smallest); ++n, sign = -sign) {
    const int depth = startingDepth + sign * (n + 1) / 2;
    if (bothPossible(depth, lumaWidth, lumaHeight, chromaWidth, chromaHeight)) return depth;
  }
  impossible();
}

const int suggestSliceSize(const int waveletDepth, const int lengthLuma, const int lengthChroma) {
  const int pl = paddedSize(lengthLuma, waveletDepth), pc = paddedSize(lengthChroma, waveletDepth);
  return pl / gcd_of(pl, pc);
}

const int suggestSliceSize(const int waveletDepth, const int lengthLuma, const int lengthChroma, int startingSliceSize) {
  const int most = std::min(lengthLuma, lengthChroma) >> waveletDepth;
  if (startingSliceSize > most) startingSliceSize = most;
  int sign = 1;
  for (int n = 0; n < 2 * most; ++n, sign = -sign) {   // 0, -1, +1, -2, +2 ...
    const int candidate = startingSliceSize + sign * (n + 1) / 2;
    if (sliceSizeIsValid(waveletDepth, lengthLuma, lengthChroma, candidate)) return candidate;
  }
  impossible();
}

// ---- the transforms: GPU ------------------------------------------------------------------------------

const Array2D waveletTransform(const Array2D& picture, WaveletKernel kernel, int depth) {
  const int h = (int)picture.shape()[0], w = (int)picture.shape()[1];
  if (kernel == NullKernel) throw std::invalid_argument("vc2: NullKernel is a reference test stub, not a VC-2 kernel");
  Array2D out(extents[paddedSize(h, depth)][paddedSize(w, depth)]);
  check(vc2_dwt_forward(ctx(), picture.data(), h, w, (int)kernel, depth, out.data()));
  return out;
}

const Array2D inverseWaveletTransform(const Array2D& transform, WaveletKernel kernel, int depth, Shape2D shape) {
  if (kernel == NullKernel) throw std::invalid_argument("vc2: NullKernel is a reference test stub, not a VC-2 kernel");
  Array2D out(extents[shape[0]][shape[1]]);
  check(vc2_dwt_inverse(ctx(), transform.data(), (int)transform.shape()[0], (int)transform.shape()[1], (int)kernel, depth,
                        out.data(), (int)shape[0], (int)shape[1]));
  return out;
}

const Picture waveletTransform(const Picture& picture, enum WaveletKernel kernel, int depth) {
  const PictureFormat f = picture.format();
  const PictureFormat padded(paddedSize(f.lumaHeight(), depth), paddedSize(f.lumaWidth(), depth),
                             paddedSize(f.chromaHeight(), depth), paddedSize(f.chromaWidth(), depth), f.chromaFormat());
  return Picture(padded, waveletTransform(picture.y(), kernel, depth), waveletTransform(picture.c1(), kernel, depth),
                 waveletTransform(picture.c2(), kernel, depth));
}

const Picture inverseWaveletTransform(const Picture& transform, enum WaveletKernel kernel, int depth, PictureFormat format) {
  return Picture(format, inverseWaveletTransform(transform.y(), kernel, depth, format.lumaShape()),
                 inverseWaveletTransform(transform.c1(), kernel, depth, format.chromaShape()),
                 inverseWaveletTransform(transform.c2(), kernel, depth, format.chromaShape()));
}

const Array1D quantMatrix(WaveletKernel kernel, int depth) {
  if (depth < 0) throw std::domain_error("wavelet depth may not be < 0");
  Array1D m(extents[3 * depth + 1]);
  check(vc2_quant_matrix((int)kernel, depth, m.data()));
  return m;
}

// ---- in-place order <-> list of subbands: an index permutation ---------------------------------------------
// band 0 is LL (every 2^depth-th sample); level L = 1..depth contributes HL, LH, HH at stride 2^(depth+1-L), phase stride/2

namespace {
struct BandLattice { int stride, oy, ox; };
BandLattice lattice(int band, int depth) {
  if (band == 0) return {1 << depth, 0, 0};
  const int level = (band - 1) / 3 + 1, kind = (band - 1) % 3;   // 0 HL, 1 LH, 2 HH
  const int stride = 1 << (depth + 1 - level), half = stride / 2;
  return {stride, kind == 0 ? 0 : half, kind == 1 ? 0 : half};
}
}  // namespace

const BlockVector split_into_subbands(const Array2D& picture, const char waveletDepth) {
  const int depth = waveletDepth, h = (int)picture.shape()[0], w = (int)picture.shape()[1];
  BlockVector bands(extents[3 * depth + 1]);
  for (int b = 0; b < 3 * depth + 1; ++b) {
    const BandLattice l = lattice(b, depth);
    const int bh = h / l.stride, bw = w / l.stride;
    bands[b].resize(extents[bh][bw]);
    for (int y = 0; y < bh; ++y)
      for (int x = 0; x < bw; ++x) bands[b][y][x] = picture[l.oy + y * l.stride][l.ox + x * l.stride];
  }
  return bands;
}

const Array2D merge_subbands(const BlockVector& subbands) {
  const int depth = ((int)subbands.size() - 1) / 3;
  const int h = (int)subbands[0].shape()[0] << depth, w = (int)subbands[0].shape()[1] << depth;
  Array2D picture(extents[h][w]);
  for (int b = 0; b < 3 * depth + 1; ++b) {
    const BandLattice l = lattice(b, depth);
    const int bh = h / l.stride, bw = w / l.stride;
    for (int y = 0; y < bh; ++y)
      for (int x = 0; x < bw; ++x) picture[l.oy + y * l.stride][l.ox + x * l.stride] = subbands[b][y][x];
  }
  return picture;
}
