/* vc2_host.h - C-ABI of the host-side stream framing in libvc2host.so (pure host code, no GPU).
 *
 * The C++ interface is include/vc2/DataUnit.h; these entry points expose the same writer / reader to C,
 * ctypes and the CPU parity tests.  Reference: src/Library/src/DataUnit.cpp (parse info :80-123, HQ picture
 * header :236-266, end of sequence :364-368, sequence header :435-881 and :883-1060, 1203-1312, picture
 * preamble :1314-1410).  Every function returns a length / count >= 0, or a negative value on error.
 */
#ifndef VC2_HOST_H
#define VC2_HOST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Sequence header data unit (parse info + video format) as EncodeStream writes it (EncodeStream.cpp:437-447):
 * profile_hq 1 = HQ profile, 0 = LD; colour_format 0/1/2 = 4:4:4 / 4:2:2 / 4:2:0; frame_rate = FrameRate index. */
int vc2host_sequence_header(int profile_hq, int height, int width, int colour_format, int interlace, int frame_rate,
                            int top_field_first, int bitdepth, uint8_t* out, int cap);

/* A whole non-fragmented HQ stream: sequence header, one HQ picture data unit per payload (picture numbers
 * 0..n-1), end of sequence - EncodeStream.cpp:437-447, 583-608, 773-775.  Returns the stream length. */
long long vc2host_wrap_hq_stream(int height, int width, int colour_format, int frame_rate, int top_field_first, int bitdepth,
                                 int kernel, int depth, int slices_x, int slices_y, int prefix, int scalar,
                                 int n, const uint8_t* const* payloads, const size_t* payload_len, uint8_t* out, size_t cap);

/* Walk the data units of a stream (DecodeStream.cpp:203-230).  units: 4 values per unit = parse code, offset of the
 * parse info, next_parse_offset, prev_parse_offset.  Returns the number of units found. */
int vc2host_parse_units(const uint8_t* data, size_t len, int max_units, int64_t* units);

/* Read the sequence header that starts at data + offset (just behind its parse info).
 * fields[10] = major_version, profile (0 LD / 3 HQ), height, width, colour_format, interlace, frame_rate,
 * top_field_first, bitdepth, bytes consumed. */
int vc2host_read_sequence_header(const uint8_t* data, size_t len, size_t offset, int32_t* fields);

/* Read a picture header + transform parameters at data + offset (just behind the parse info).
 * fields[9] = picture number, wavelet index, depth, slices_x, slices_y, prefix (HQ) or slice-bytes numerator (LD),
 * scalar (HQ) or slice-bytes denominator (LD), bytes consumed, 0.  major_version: from the sequence header. */
int vc2host_read_picture_header(const uint8_t* data, size_t len, size_t offset, int ld, int major_version, int64_t* fields);

#ifdef __cplusplus
}
#endif
#endif
