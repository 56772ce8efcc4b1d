/* vc2_cabi.h - C-ABI of the B200-native VC-2 HQ/LD hot path (libvc2b200.so).
 *
 * Plain C: pointers, sizes and PODs only.  No exceptions cross this boundary;
 * every call returns VC2_OK (0) or a negative vc2_status, and the matching
 * reference exception text can be fetched with vc2_last_error().
 *
 * bbc/vc2-reference has no FFI: its boundary is the set of free functions in
 * the src/Library headers called by EncodeStream.cpp / DecodeStream.cpp.  Each entry
 * point below names the reference function(s) it replaces (paths relative to
 * /root/reference).  The C++ mirror of the Library headers (the include/vc2/ headers)
 * is a thin layer over these calls; INTEGRATION.md shows the binding.
 *
 * Layout contract (same as the reference's Array2D, src/Library/Arrays.h:28-31):
 * a plane is row-major contiguous int32, [y][x]; transformed planes are the
 * PADDED size (each dimension rounded up to a multiple of 2^depth) in the
 * reference's IN-PLACE INTERLEAVED coefficient order
 * (src/Library/src/WaveletTransform.cpp:262-281, 428-450).
 *
 * "host" entry points take host pointers, do their own H2D/D2H and return
 * after the result is in the caller's buffer.  "_dev" entry points take device
 * pointers, are asynchronous on the context stream and do no allocation on the
 * hot path.  There is NO CPU fallback: without a CUDA device every compute
 * call fails with VC2_ERR_CUDA.
 */
#ifndef VC2_CABI_H
#define VC2_CABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums ---------------------------------------------------------------- */

/* WaveletKernel, same numeric values as src/Library/WaveletTransform.h:26 and the
 * wavelet_index on the wire (src/Library/src/DataUnit.cpp:246). */
enum vc2_kernel {
  VC2_DD97 = 0, VC2_LEGALL = 1, VC2_DD137 = 2, VC2_HAAR0 = 3, VC2_HAAR1 = 4, VC2_FIDELITY = 5, VC2_DAUB97 = 6
};

/* slice coding modes (src/Library/Slices.h:98 sliceio::SliceIOMode) */
enum vc2_slice_mode { VC2_HQ_VBR = 0, VC2_HQ_CBR = 1, VC2_LD = 2 };

enum vc2_status {
  VC2_OK = 0,
  VC2_ERR_ARG = -1,            /* bad argument (null pointer, bad kernel/depth/geometry)        */
  VC2_ERR_CUDA = -2,           /* CUDA runtime error, or no device: there is no CPU fallback    */
  VC2_ERR_SCALAR_TOO_SMALL = -3,   /* Slices.cpp:115-117 "Slice scalar is too small, ..."        */
  VC2_ERR_QUANT_INDEX = -4,    /* Quantisation.cpp:60-63 "quantization index exceeds ..."       */
  VC2_ERR_CBR_TOO_MANY_BYTES = -5, /* Slices.cpp:356-358 "SliceIO, HQ CBR mode: Too many bytes..." */
  VC2_ERR_CBR_COMPONENT_LENGTH = -6, /* Slices.cpp:359-366 "Slice component length exceeds 1 byte..." */
  VC2_ERR_CAPACITY = -7,       /* caller's output buffer too small                                */
  VC2_ERR_VLC_RANGE = -8,      /* |quantised coefficient| >= 65535: reference VLC is UB (VLC.h:27) */
  VC2_ERR_STREAM = -9,         /* malformed / truncated slice data on decode                       */
  VC2_ERR_LD_TOO_MANY_BYTES = -10  /* Slices.cpp:209-211 "SliceIO, LD mode: Too many bytes..."   */
};

/* per-slice error flag bits written by the slice kernels (err_flags arrays) */
#define VC2_FLAG_SCALAR_TOO_SMALL   0x01u
#define VC2_FLAG_QUANT_INDEX        0x02u
#define VC2_FLAG_CBR_TOO_MANY_BYTES 0x04u
#define VC2_FLAG_CBR_COMP_LENGTH    0x08u
#define VC2_FLAG_VLC_RANGE          0x10u
#define VC2_FLAG_STREAM             0x20u
#define VC2_FLAG_LD_TOO_MANY_BYTES  0x40u
#define VC2_FLAG_SEARCH_PHASE       0x80u  /* raised inside quantIndicesCBR (before any slice is written) */

/* ---- PODs ----------------------------------------------------------------- */

/* Geometry of one coded picture.  Mirrors PictureFormat (src/Library/Picture.h:23-70)
 * plus the transform / slice parameters EncodeStream derives (EncodeStream.cpp:368-375). */
typedef struct vc2_geom {
  int32_t luma_h, luma_w;       /* picture (unpadded) luma size                       */
  int32_t chroma_h, chroma_w;   /* picture (unpadded) chroma size                     */
  int32_t kernel;               /* enum vc2_kernel                                    */
  int32_t depth;                /* wavelet depth, 1..6                                */
  int32_t slices_y, slices_x;   /* slices per picture (sliceSizeIsValid results)      */
  int32_t prefix, scalar;       /* HQ slice prefix bytes, slice size scalar           */
} vc2_geom;

/* Sample format of raw planar picture files (src/Library/src/Arrays.cpp:333-426,
 * EncodeStream.cpp:319-322): big-endian words, MSB-justified, offset binary. */
typedef struct vc2_sample_format {
  int32_t bytes_per_sample;     /* 1 or 2 on the fused path (3,4 via the int32 Library path) */
  int32_t luma_depth;           /* bits */
  int32_t chroma_depth;         /* bits */
} vc2_sample_format;

typedef struct vc2_ctx vc2_ctx;         /* one per (process, GPU): stream, scratch, tables   */
typedef struct vc2_codec vc2_codec;     /* batched fused encoder/decoder for one geometry    */

/* ---- context --------------------------------------------------------------- */

vc2_ctx* vc2_create(int device);                    /* NULL if the device cannot be opened   */
void vc2_destroy(vc2_ctx* ctx);
int vc2_set_stream(vc2_ctx* ctx, void* cuda_stream); /* run on the caller's cudaStream_t (e.g. torch's) */
int vc2_synchronize(vc2_ctx* ctx);
const char* vc2_last_error(vc2_ctx* ctx);           /* reference exception text of the last failure */
const char* vc2_status_message(int status);         /* same text, by status code                */
int vc2_device_count(void);
int vc2_kernel_launches(vc2_ctx* ctx, int reset);   /* kernels launched through this context    */
/* page-locked (pinned, portable across devices) host memory for the buffers of the *_host entry points; NULL when it
 * cannot be had (the caller may then use ordinary memory: same results, slower copies) */
void* vc2_host_alloc(size_t bytes);
/* pin the calling thread (and threads it starts afterwards) to the CPUs local to a GPU (sysfs local_cpulist of its PCI
 * device): host buffers allocated and filled afterwards sit on the GPU's NUMA node.  VC2_ERR_ARG when the topology is
 * not readable; nothing is changed then */
int vc2_bind_thread_to_device(int device);
void vc2_host_free(void* p);

/* per-kernel timing for the roofline report: CUDA events recorded on the launch stream around every
 * kernel launched through this context (no reference counterpart; measurement only) */
enum vc2_stage {
  VC2_STAGE_DWT_L0 = 0,    /* forward lifting, finest level (reads the picture)          */
  VC2_STAGE_DWT_DEEP = 1,  /* forward lifting, remaining levels                          */
  VC2_STAGE_PACK = 2,      /* quantise + [CBR search] + exp-Golomb slice packing         */
  VC2_STAGE_UNPACK = 3,    /* slice parsing + inverse quantisation                       */
  VC2_STAGE_IDWT_DEEP = 4, /* inverse lifting, coarse levels                             */
  VC2_STAGE_IDWT_L0 = 5,   /* inverse lifting, finest level (writes the picture)         */
  VC2_STAGE_LD_DC = 6,     /* LD LL-band DC prediction wavefront                         */
  VC2_STAGE_ASSEMBLE = 7,  /* slice size scan + gather of the slice images into the payload */
  VC2_STAGE_INDEX = 8,     /* HQ slice index: the walk over the slice length bytes of a payload */
  VC2_STAGE_SEARCH = 9,    /* HQ_CBR rate control: quantIndicesCBR per slice (its own launch, before the packing launch) */
  VC2_NUM_STAGES = 10
};
int vc2_profile_enable(vc2_ctx* ctx, int on);
int vc2_profile_read(vc2_ctx* ctx, float* ms, int* launches, int nstages);

/* ---- host-side helpers (pure host code, no GPU needed) ---------------------- */

/* paddedSize  - WaveletTransform.cpp:74-77 */
int vc2_padded_size(int size, int depth);
/* sliceSizeIsValid - WaveletTransform.cpp:116-136 (returns number of slices or 0) */
int vc2_slice_size_is_valid(int depth, int luma_len, int chroma_len, int n_size);
/* quantMatrix - WaveletTransform.cpp:345-423 ; out[3*depth+1] */
int vc2_quant_matrix(int kernel, int depth, int32_t* out);
/* slice_bytes - Slices.cpp:28-49 ; out[ny*nx] */
int vc2_slice_bytes(int ny, int nx, int total_bytes, int scalar, int32_t* out);
/* quant_factor / quant_offset tables - Quantisation.cpp:40-83 */
int vc2_quant_factor(int q);
int vc2_quant_offset(int q);
/* multiplier and shift the slice coders divide by quant_factor(q) with: a / quant_factor(q) == mulhi(a, m) >> shift
 * for every a < 2^31 (the reference's own domain: (abs(v) << 2) in int, Quantisation.cpp:69-76); for the tests */
int vc2_quant_magic31(int q, uint32_t* m, uint32_t* shift);
/* fill a vc2_geom from picture size + colour format (0=4:4:4, 1=4:2:2, 2=4:2:0) and -u/-a slice sizes;
 * returns VC2_ERR_ARG when sliceSizeIsValid rejects the combination (EncodeStream.cpp:374-405) */
int vc2_make_geom(int height, int width, int chroma_format, int kernel, int depth,
                  int v_slice_size, int h_slice_size, int prefix, int scalar, vc2_geom* out);
/* walk the length bytes of an HQ picture payload and produce slice start offsets
 * (Slices.cpp:535-612 read order); offsets[n_slices+1]; VC2_ERR_STREAM if it runs off the end */
int vc2_hq_index_slices(const uint8_t* payload, size_t len, int n_slices, int prefix, int scalar,
                        uint32_t* offsets);

/* ---- Library-surface operations, host buffers -------------------------------- */

/* waveletTransform(Array2D,kernel,depth) - WaveletTransform.cpp:262-281 (pads: :79-94) */
int vc2_dwt_forward(vc2_ctx*, const int32_t* src, int h, int w, int kernel, int depth,
                    int32_t* dst /* paddedH x paddedW */);
/* inverseWaveletTransform(Array2D,kernel,depth,shape) - WaveletTransform.cpp:321-342 */
int vc2_dwt_inverse(vc2_ctx*, const int32_t* src, int ph, int pw, int kernel, int depth,
                    int32_t* dst, int h, int w);
/* quantise_transform_np(Array2D, Array2D qIndices, qMatrix) - Quantisation.cpp:479-489 */
int vc2_quantise_np(vc2_ctx*, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                    const int32_t* qidx, int ny, int nx, int32_t* out);
/* inverse_quantise_transform_np - Quantisation.cpp:534-544 */
int vc2_dequantise_np(vc2_ctx*, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                      const int32_t* qidx, int ny, int nx, int32_t* out);
/* inverse_quantise_transform (LD, DC-predicted LL band) - Quantisation.cpp:369-379, 287-306 */
int vc2_dequantise_ld(vc2_ctx*, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                      const int32_t* qidx, int ny, int nx, int32_t* out);
/* operator<<(ostream&, Slices) with the HQ VBR / HQ CBR writers - Slices.cpp:645-660, 305-382, 469-533.
 * qY/qU/qV: QUANTISED padded planes (in-place order).  slice_bytes: per-slice budget (CBR) or NULL.
 * slice_off (optional): n_slices+1 byte offsets of the slices inside out. */
int vc2_hq_pack(vc2_ctx*, const int32_t* qY, const int32_t* qU, const int32_t* qV, const vc2_geom* g,
                const int32_t* qidx, int mode, const int32_t* slice_bytes,
                uint8_t* out, size_t cap, size_t* out_len, uint32_t* slice_off);
/* operator>>(istream&, Slices) with the HQ VBR reader - Slices.cpp:662-694, 535-612 */
int vc2_hq_unpack(vc2_ctx*, const uint8_t* in, size_t len, const vc2_geom* g,
                  int32_t* qY, int32_t* qU, int32_t* qV, int32_t* qidx);
/* operator>>(istream&, Slices) with the LD reader - Slices.cpp:246-303 ; slice_bytes[ny*nx] */
int vc2_ld_unpack(vc2_ctx*, const uint8_t* in, size_t len, const vc2_geom* g, const int32_t* slice_bytes,
                  int32_t* qY, int32_t* qU, int32_t* qV, int32_t* qidx);
/* operator<<(ostream&, Slices) with the LD writer - Slices.cpp:195-244, 645-660.  qY/qU/qV: quantised padded planes
 * (LL band: the quantised prediction residuals of quantise_transform); slice_bytes[ny*nx] */
int vc2_ld_pack(vc2_ctx*, const int32_t* qY, const int32_t* qU, const int32_t* qV, const vc2_geom* g,
                const int32_t* qidx, const int32_t* slice_bytes, uint8_t* out, size_t cap, size_t* out_len);
/* quantise_transform(Array2D, qIndices, qMatrix), the LD quantiser with DC prediction of the LL band -
 * Quantisation.cpp:213-282, 353-367 */
int vc2_quantise_ld(vc2_ctx*, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                    const int32_t* qidx, int ny, int nx, int32_t* out);
/* luma_slice_bits (q2 == NULL) / chroma_slice_bits (q, q2 walked interleaved) - Slices.cpp:51-96: for every slice of
 * the quantised in-place plane(s), the bits of its code list up to and including its last non-zero coefficient;
 * bits[ny*nx] */
int vc2_slice_bits(vc2_ctx*, const int32_t* q, const int32_t* q2, int ph, int pw, int depth, int ny, int nx,
                   int32_t* bits);
/* component_slice_bytes - Slices.cpp:97-119, for every slice of one quantised in-place plane; bytes[ny*nx];
 * VC2_ERR_SCALAR_TOO_SMALL as the reference throws */
int vc2_hq_slice_sizes(vc2_ctx*, const int32_t* q, int ph, int pw, int depth, int ny, int nx, int scalar,
                       int32_t* bytes);
/* quantIndicesCBR - EncodeStream.cpp:73-125 ; cY/cU/cV: UNQUANTISED padded planes */
int vc2_cbr_qindices(vc2_ctx*, const int32_t* cY, const int32_t* cU, const int32_t* cV, const vc2_geom* g,
                     const int32_t* qmatrix, const int32_t* slice_bytes, int32_t* qidx_out,
                     uint32_t* err_flags /* optional, ny*nx */);

/* ---- fused, batched picture codec (the hot path proper) ----------------------- */

typedef struct vc2_codec_params {
  vc2_geom geom;
  vc2_sample_format fmt;
  int32_t mode;                 /* VC2_HQ_VBR (HQ_ConstQ), VC2_HQ_CBR or VC2_LD (decode only)        */
  int32_t qindex;               /* HQ_ConstQ: the fixed quantiser index (EncodeStream.cpp:128-138)    */
  int32_t picture_bytes;        /* HQ_CBR / LD: compressed bytes per picture (-s)                     */
  int32_t max_pictures;         /* batch capacity: pictures resident on the device at once            */
} vc2_codec_params;

vc2_codec* vc2_codec_create(vc2_ctx*, const vc2_codec_params*);
void vc2_codec_destroy(vc2_codec*);
size_t vc2_codec_picture_in_bytes(const vc2_codec*);     /* raw planar bytes of one picture           */
size_t vc2_codec_payload_capacity(const vc2_codec*);     /* worst-case slice payload bytes / picture  */

/* device-resident stages (asynchronous on the context stream).  Slot = picture index in the batch.
 *   encode: raw samples (device) -> DWT -> [CBR search] -> quantise + slice pack -> payload (device)
 *     replaces EncodeStream.cpp:456-565 + Slices.cpp:645-660 for n pictures
 *   decode: payload + payload length (device) -> slice index from the length bytes (HQ, Slices.cpp:544-605)
 *     -> parse + dequantise -> IDWT -> clip -> raw samples
 *     replaces DecodeStream.cpp:512-605 for n pictures.  The slice offsets are always rebuilt from the payload;
 *     the table the encoder left behind is overwritten, not used. */
int vc2_codec_encode_dev(vc2_codec*, int n_pictures);
int vc2_codec_decode_dev(vc2_codec*, int n_pictures);
/* pipelined mode (default off): consecutive encode_dev / decode_dev calls with the same n_pictures are ordered per
 * sub-batch (same slots, same internal stream) instead of call by call, so the serial slice-index walk of one
 * sub-batch overlaps the kernels of the others.  The context stream still waits for every sub-batch of every call,
 * and the codec's own upload / host-buffer calls re-synchronise the sub-batches.  Work the CALLER puts on the
 * context stream between two calls (e.g. its own kernels writing vc2_codec_samples_dev) is only ordered before the
 * next call when pipelined mode is off. */
int vc2_codec_set_pipelined(vc2_codec*, int on);

/* device buffers owned by the codec (for device-resident use and for tests) */
void* vc2_codec_samples_dev(vc2_codec*, int slot);       /* raw planar picture bytes: encoder input   */
void* vc2_codec_recon_dev(vc2_codec*, int slot);         /* raw planar picture bytes: decoder output  */
uint8_t* vc2_codec_payload_dev(vc2_codec*, int slot);    /* slice payload                                   */
int32_t* vc2_codec_coeffs_dev(vc2_codec*, int slot);     /* group-interleaved coefficient block */
uint32_t* vc2_codec_slice_offsets_dev(vc2_codec*, int slot);   /* n_slices+1                              */

/* host <-> slot transfers (pinned staging inside; asynchronous, ordered on the context stream) */
int vc2_codec_upload_picture(vc2_codec*, int slot, const void* raw_planar);
int vc2_codec_download_picture(vc2_codec*, int slot, void* raw_planar);
int vc2_codec_upload_payload(vc2_codec*, int slot, const uint8_t* payload, size_t len);
/* blocks until the slot's encode finished; returns payload length, per-slice qindex and offsets (optional) */
int vc2_codec_download_payload(vc2_codec*, int slot, uint8_t* payload, size_t cap, size_t* len,
                               int32_t* qidx /* optional ny*nx */, uint32_t* slice_off /* optional n+1 */);
/* taps used by the parity tests (reference: EncodeStream -o Transform/Quantised/Indices) -
 * in-place interleaved padded planes, host pointers */
int vc2_codec_read_transform(vc2_codec*, int slot, int32_t* y, int32_t* u, int32_t* v);
int vc2_codec_read_quantised(vc2_codec*, int slot, int32_t* y, int32_t* u, int32_t* v);
int vc2_codec_read_indices(vc2_codec*, int slot, int32_t* qidx);
/* first failing slice of the last encode/decode of this slot -> vc2_status (reference throw order) */
int vc2_codec_slot_status(vc2_codec*, int slot);

/* end-to-end, host buffers in and out, copies overlapped with compute on internal streams:
 *   pictures[i]: raw planar picture bytes;  payloads[i]: caller buffer of payload_cap bytes */
int vc2_codec_encode_host(vc2_codec*, int n, const void* const* pictures,
                          uint8_t* const* payloads, size_t payload_cap, size_t* payload_len);
int vc2_codec_decode_host(vc2_codec*, int n, const uint8_t* const* payloads, const size_t* payload_len,
                          void* const* pictures);

#ifdef __cplusplus
}
#endif
#endif /* VC2_CABI_H */
