// HQ / LD slice coding kernels: dead-zone quantisation, interleaved exp-Golomb packing and
// parsing, HQ_CBR rate control.  Reference semantics (paths relative to /root/reference):
//   quant / scale            src/Library/src/Quantisation.cpp:40-95
//   component_slice_bytes    src/Library/src/Slices.cpp:97-119
//   HQ slice writers/readers src/Library/src/Slices.cpp:305-382 (CBR), 469-533 (VBR), 535-612 (read)
//   LD slice reader          src/Library/src/Slices.cpp:246-303
//   SignedVLC / bounded IO   src/Library/src/VLC.cpp:21-94, 151-213, 229-257
//   quantIndicesCBR          src/EncodeStream/EncodeStream.cpp:73-125 ; yss_for_slice Quantisation.cpp:627-642
#pragma once
#include "vc2_common.cuh"

namespace vc2 {

struct SliceGeom {
  PlaneGeom plane[3];        // Y, C1, C2 padded plane geometry
  long long coef_pic_stride; // elements between pictures = 32 * ceil(slices / 32) * comp_start[3]
  int depth, nbands;
  int slices_x, slices_y;
  int prefix, scalar;
  int qmatrix[VC2_MAX_BANDS];
  // per component class: slice part of each band (rows x cols) and scan-order start of each band
  int part_h[3][VC2_MAX_BANDS];
  int part_w[3][VC2_MAX_BANDS];
  int band_start[3][VC2_MAX_BANDS + 1];  // band_start[c][nbands] = coefficients per slice component
  int comp_start[4];                     // start of each component in the per-slice coefficient list
};

struct PackParams {
  SliceGeom g;
  const int32_t* coef;          // group-interleaved coefficients [pic]
  int mode;                     // VC2_HQ_VBR / VC2_HQ_CBR
  int quantise;                 // 1: coefficients are unquantised, apply quant(); 0: already quantised
  int search;                   // 1: run quantIndicesCBR per slice and use (and store) its result
  int const_q;                  // >= 0: every slice uses this index (HQ_ConstQ); < 0: read qidx[]
  int emit;                     // 0: rate control only (vc2_cbr_qindices)
  int after_search;             // 1: a rate control launch has left qidx[] and err_flags[] (a slice it flagged is not packed)
  int32_t* qidx;                // [pic][slices] in (search == 0 && const_q < 0) or out
  const int32_t* slice_bytes;   // [slices] per-slice byte budget (CBR), same for every picture
  uint32_t* staging;            // [pic][slices][wcap] slice images as MSB-first 32-bit words
  int wcap;                     // staging words per slice (worst case: every code 32 bits)
  uint32_t* sizes;              // [pic][slices] out: bytes of each coded slice
  uint32_t* err_flags;          // [pic][slices] out (VC2_FLAG_*)
  int narrow;                   // 1: coef is the narrow block (16-bit sign-magnitude, already quantised; HQ_ConstQ only)
  // fused scan + gather (narrow packer): the CTA scans its slice sizes, learns the bytes in front of it from the CTAs
  // before it (decoupled look-back over tile_state) and copies its slice images - still in L2 - to their place in the
  // payload: no second pass over the staging buffer, no scan / gather launches
  int fuse;
  unsigned long long* tile_state;   // [pic][tiles]: (state << 32) | bytes; state 0 = nothing, 1 = the tile's own bytes, 2 = all bytes up to and with it
  uint32_t* tile_ticket;            // [pic] next tile to hand out (tiles are taken in arrival order: a tile never waits for one that has not started)
  int tiles;                        // CTAs per picture
  uint8_t* out;                     // payload [pic]
  long long out_pic_stride, out_capacity;
  uint32_t* slice_off;              // [pic][slices + 1] out
  uint32_t* total_len;              // [pic] out (may be NULL)
};

struct AssembleParams {         // exclusive scan of the slice sizes and the gather into the payload
  int nslices;
  const uint32_t* sizes;        // [pic][slices]
  const uint32_t* fixed_off;    // [slices + 1] slice offsets known a priori (CBR) or NULL (scan the sizes)
  uint32_t* slice_off;          // [pic][slices + 1] out
  uint32_t* total_len;          // [pic] out: payload bytes of each picture (NULL: not wanted)
  const uint32_t* staging;      // [pic][slices][wcap]
  int wcap;
  uint8_t* out;                 // payload [pic]
  long long out_pic_stride;
  long long out_capacity;       // bytes available per picture
  uint32_t* err_flags;          // [pic][slices]: VC2_FLAG_STREAM when the payload does not fit
};

struct UnpackParams {
  SliceGeom g;
  const uint8_t* in;            // payload [pic]
  long long in_pic_stride;
  const uint32_t* slice_off;    // [pic][slices + 1]
  long long slice_off_pic_stride;  // 0 when every picture shares one table (CBR / LD)
  int32_t* coef;                // group-interleaved coefficients out [pic]
  int32_t* qidx;                // [pic][slices] out
  uint32_t* err_flags;          // [pic][slices]
  int dequantise;               // 1: store scale(v, q'); 0: store the quantised value
  int ld;                       // 1: LD slice syntax (Slices.cpp:246-303); LL band left quantised
  int narrow;                   // 1: coef is the narrow block: 16-bit sign-magnitude words of the QUANTISED coefficients (HQ only)
  uint32_t* narrow_ovf;         // [pic] set when a magnitude does not fit the narrow block
  BandScale* band_scale;        // [pic] narrow: one index for the whole picture? and its scale factors (for the inverse lifting kernels)
};

// slice index of HQ payloads on the device (the reader's walk over the length bytes, Slices.cpp:544-605)
#define VC2_INDEX_MAX_PICTURES 8
struct IndexParams {
  const uint8_t* in;            // payload [pic]
  long long in_pic_stride;      // bytes; also the readable size of one picture's buffer
  uint32_t len[VC2_INDEX_MAX_PICTURES];   // payload bytes per picture (len_dev == NULL)
  const uint32_t* len_dev;      // [pic] payload bytes per picture in device memory; lifts the picture limit
  uint32_t* slice_off;          // [pic][slices + 1] out
  int nslices, prefix, scalar;
};

struct QuantParams {            // stand-alone quantise / dequantise on IN-PLACE ordered planes
  const int32_t* src;
  int32_t* dst;
  const int32_t* qidx;          // [slices_y][slices_x]
  int ph, pw, depth;
  int slices_y, slices_x;
  int qmatrix[VC2_MAX_BANDS];
  int inverse;                  // 0: quant, 1: scale
  int skip_ll;                  // LD: leave the LL band to the DC-prediction kernel
};

struct LdDcParams {             // LD LL-band reconstruction with DC prediction (Quantisation.cpp:287-306)
  int32_t* base;                // in-place plane (interleaved == 0) or one picture's group-interleaved block
  const int32_t* qidx;
  int H, W;                     // LL band dims of the whole picture
  int interleaved;              // 0: LL sample (y, x) at base[(y * pitch + x) << depth]; 1: coef_index(slice, k0 + ...)
  long long pitch;              // in-place: padded plane width
  int depth;
  int bh, bw;                   // interleaved: LL part of one slice
  int k0, nc4;                  // interleaved: comp_start of the component, NC / 4
  int slices_y, slices_x;
  int qm0;
  // a batch: grid.x = picture, grid.y = plane; picture i of plane j starts at base + i * base_pic_stride (+ the plane's own base)
  long long base_pic_stride, qidx_pic_stride;
};
struct LdDcBatch { LdDcParams c[3]; int nplanes; };
cudaError_t ld_dc_batch_launch(cudaStream_t s, const LdDcBatch& b, int npictures);

// LD encoder (EncodeStream.cpp:139-245 quantIndicesLD, Quantisation.cpp:213-282 predictive quantiser,
// Slices.cpp:51-96 slice bit counts, :195-244 LD slice writer)
struct LdEncParams {
  SliceGeom g;
  const int32_t* coef;          // [pic] group-interleaved transform coefficients
  int32_t* qcoef;               // [pic] same layout, out: quantised coefficients (LL band: quantised prediction residual)
  uint32_t* acbits;             // [pic][slice][128][2]: bits up to the last non-zero AC coefficient, luma | chroma pair
  int32_t* qidx;                // [pic][slices] out
  int32_t* restored;            // [pic][3][LL band] scratch: locally decoded LL band
  long long ll_stride;          // elements per picture of `restored` (3 planes)
  int ll_off[3], ll_w[3];       // start and width of each component's LL band inside a picture's block
  const int32_t* slice_bytes;   // [slices]
  uint32_t* staging;            // [pic][slices][wcap]
  int wcap;
  uint32_t* sizes;              // [pic][slices] out
  uint32_t* err_flags;          // [pic][slices] out
  int prequantised;             // 1 (ld_pack_launch): qcoef already holds every quantised coefficient, qidx is an input
};
cudaError_t ld_encode_launch(cudaStream_t s, const LdEncParams& p, int npictures);
// the LD slice writer alone, on an already quantised block (operator<<(ostream&, Slices) in LD mode, Slices.cpp:195-244)
cudaError_t ld_pack_launch(cudaStream_t s, const LdEncParams& p, int npictures);

// bits of every slice's code list up to and including its last non-zero coefficient, on IN-PLACE ordered quantised planes
// (luma_slice_bits / chroma_slice_bits / the count inside component_slice_bytes, Slices.cpp:51-119); q2 != NULL: the two
// planes are walked interleaved, as the LD chroma list is
struct SliceBitsParams {
  const int32_t* q;
  const int32_t* q2;
  int ph, pw, depth, slices_y, slices_x;
  int32_t* bits;                // [slices_y][slices_x] out
};
cudaError_t slice_bits_launch(cudaStream_t s, const SliceBitsParams& p);

// forward LD LL-band quantiser with DC prediction (quantise_LLSubband, Quantisation.cpp:213-236) on an in-place plane
struct LdDcQuantParams {
  const int32_t* src;           // in-place plane of transform coefficients
  int32_t* dst;                 // in-place plane: the LL positions receive the quantised prediction residuals
  int32_t* restored;            // [H][W] scratch: the locally decoded LL band
  const int32_t* qidx;          // [slices_y][slices_x]
  int H, W;                     // LL band dims
  long long pitch;              // padded plane width
  int depth, slices_y, slices_x, qm0;
};
cudaError_t ld_dc_quant_launch(cudaStream_t s, const LdDcQuantParams& p);

cudaError_t upload_quant_tables(const QuantTables& t);
cudaError_t pack_launch(cudaStream_t s, const PackParams& p, int npictures);
cudaError_t assemble_launch(cudaStream_t s, const AssembleParams& p, int npictures);
cudaError_t unpack_launch(cudaStream_t s, const UnpackParams& p, int npictures, const uint2* scale_tab = nullptr);
cudaError_t index_launch(cudaStream_t s, const IndexParams& p, int npictures);
cudaError_t layout_launch(cudaStream_t s, bool to_slice_major, const int32_t* src, int32_t* dst, const SliceGeom& g, int c);
cudaError_t quant_launch(cudaStream_t s, const QuantParams& p);
cudaError_t ld_dc_launch(cudaStream_t s, const LdDcParams& p);

}  // namespace vc2
