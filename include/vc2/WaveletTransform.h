// vc2/WaveletTransform.h - mirrors src/Library/WaveletTransform.h:26-77.  The transforms run on the GPU
// through libvc2b200.so (include/vc2_cabi.h); there is no CPU implementation behind these functions.
#ifndef VC2_WAVELETTRANSFORM_H
#define VC2_WAVELETTRANSFORM_H
#include <iosfwd>
#include "Arrays.h"
#include "Picture.h"

namespace vc2 {

// same enumerators and values as WaveletTransform.h:26 (the value is the wavelet_index on the wire)
enum WaveletKernel { DD97, LeGall, DD137, Haar0, Haar1, Fidelity, Daub97, NullKernel };
std::ostream& operator<<(std::ostream& os, WaveletKernel k);   // WaveletTransform.cpp:26-57
std::istream& operator>>(std::istream& is, WaveletKernel& k);  // :59-72 (unknown text sets failbit)

int paddedSize(int size, int depth);                                              // :74-77
int sliceSizeIsValid(int depth, int lumaLength, int chromaLength, int nSize);     // :116-136
bool waveletTransformIsPossible(int depth, int lumaLength, int chromaLength);     // :138-150
int suggestSliceSize(int depth, int lumaLength, int chromaLength, int nSize);     // :152-180
int suggestWaveletDepth(int lumaWidth, int lumaHeight, int chromaWidth, int chromaHeight, int depth);   // :182-222

// waveletTransform(Array2D / Picture, kernel, depth) - :262-281: result is the PADDED array in the
// reference's in-place interleaved coefficient order
const Array2D waveletTransform(const Array2D& picture, WaveletKernel kernel, int depth);
const Picture waveletTransform(const Picture& picture, WaveletKernel kernel, int depth);
// inverseWaveletTransform(..., shape) - :321-342: crops to `height x width` / `format`
const Array2D inverseWaveletTransform(const Array2D& transform, WaveletKernel kernel, int depth, int height, int width);
const Picture inverseWaveletTransform(const Picture& transform, WaveletKernel kernel, int depth, const PictureFormat& format);
// quantMatrix(kernel, depth) - :345-423 (std::domain_error for depth < 0 as in :348; depth is bounded by VC2_MAX_DEPTH here)
const Array1D quantMatrix(WaveletKernel kernel, int depth);

}  // namespace vc2
#endif
