// vc2/Quantisation.h - mirrors the "no prediction" entry points of src/Library/Quantisation.h:16-118
// used by the HQ path, the LD inverse used by DecodeStream, and the rate control of EncodeStream.cpp:73-138.
#ifndef VC2_QUANTISATION_H
#define VC2_QUANTISATION_H
#include "Arrays.h"
#include "Picture.h"

namespace vc2 {

// Quantisation.cpp:586-605 / 479-519: per-slice index array (ySlices x xSlices) or one index for all slices
const Picture quantise_transform_np(const Picture& coefficients, const Array2D& qIndices, const Array1D& qMatrix);
const Picture quantise_transform_np(const Picture& coefficients, int qIndex, const Array1D& qMatrix);
const Array2D quantise_transform_np(const Array2D& coefficients, const Array2D& qIndices, const Array1D& qMatrix);
// Quantisation.cpp:607-625 / 534-558
const Picture inverse_quantise_transform_np(const Picture& qCoeffs, const Array2D& qIndices, const Array1D& qMatrix);
const Picture inverse_quantise_transform_np(const Picture& qCoeffs, int qIndex, const Array1D& qMatrix);
const Array2D inverse_quantise_transform_np(const Array2D& qCoeffs, const Array2D& qIndices, const Array1D& qMatrix);
// LD inverse with DC prediction of the LL band - Quantisation.cpp:369-379, 287-306
const Picture inverse_quantise_transform(const Picture& qCoeffs, const Array2D& qIndices, const Array1D& qMatrix);

// EncodeStream.cpp:73-125 (throws the reference's std::logic_error texts) and :128-138
const Array2D quantIndicesCBR(const Picture& coefficients, const Array1D& qMatrix, const Array2D& sliceBytes, int scalar,
                              int waveletDepth);
const Array2D quantIndicesConstQ(int ySlices, int xSlices, int qIndex);

}  // namespace vc2
#endif
