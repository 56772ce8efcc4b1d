// vc2/Slices.h - slice sizes and the HQ / LD slice wire formats (src/Library/Slices.h, Slices.cpp).
// The reference serialises through iostream manipulators (sliceio::highQualityVBR(...) << Slices); here the
// coding mode is an explicit argument and the byte string is the value, the semantics are the same.
#ifndef VC2_SLICES_H
#define VC2_SLICES_H
#include <cstdint>
#include <string>
#include "Arrays.h"
#include "Picture.h"
#include "WaveletTransform.h"

namespace vc2 {

// slice_bytes(ySlices, xSlices, totalBytes, scalar) - Slices.cpp:28-49
const Array2D slice_bytes(int ySlices, int xSlices, int totalBytes, int scalar);

// A picture's worth of slices: the quantised transform (padded, in-place order; the slices are its tiles,
// Picture.cpp:231-271) and one quantiser index per slice - the data of the reference's `Slices` (Slices.h:40-60).
struct Slices {
  Picture yuvCoeffs;
  int waveletDepth;
  Array2D qIndices;    // ySlices x xSlices
};

// operator<<(ostream&, Slices) after sliceio::highQualityVBR(prefix, scalar) - Slices.cpp:469-533, 645-660
std::string writeSlicesHQVBR(const Slices& s, WaveletKernel kernel, int slicePrefix, int sliceScalar);
// ... after sliceio::highQualityCBR(sliceBytes, prefix, scalar) - Slices.cpp:305-382
std::string writeSlicesHQCBR(const Slices& s, WaveletKernel kernel, const Array2D& sliceBytes, int slicePrefix, int sliceScalar);
// operator>>(istream&, Slices) HQ reader - Slices.cpp:535-612, 662-694.  format = PADDED transform format.
Slices readSlicesHQ(const uint8_t* data, size_t len, const PictureFormat& transformFormat, WaveletKernel kernel, int waveletDepth,
                    int ySlices, int xSlices, int slicePrefix, int sliceScalar);
// LD reader - Slices.cpp:246-303
Slices readSlicesLD(const uint8_t* data, size_t len, const PictureFormat& transformFormat, WaveletKernel kernel, int waveletDepth,
                    int ySlices, int xSlices, const Array2D& sliceBytes);

}  // namespace vc2
#endif
