// vc2/DataUnit.h - VC-2 stream framing around the slice payload (src/Library/DataUnit.h, DataUnit.cpp):
// parse info headers with their next/previous offsets, sequence header (video format with base-format
// matching), HQ / LD picture headers.  Pure host code.  The reference serialises through iostream state;
// here a StreamWriter / StreamReader object carries that state (previous parse offset, major version).
#ifndef VC2_DATAUNIT_H
#define VC2_DATAUNIT_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include "Picture.h"
#include "WaveletTransform.h"

namespace vc2 {

enum DataUnitType { UNKNOWN_DATA_UNIT, SEQUENCE_HEADER, END_OF_SEQUENCE, AUXILIARY_DATA, PADDING_DATA, HQ_PICTURE, LD_PICTURE, HQ_FRAGMENT, LD_FRAGMENT };
// DataUnit.h:89-93, same enumerators and values
enum FrameRate { FR_UNSET = -1, FR0, FR24000_1001, FR24, FR25, FR30000_1001, FR30, FR50, FR60000_1001, FR60, FR15000_1001, FR25_2, FR48, FR48_1001, FR96, FR100, FR120_1001, FR120 };
enum Profile { PROFILE_UNKNOWN, PROFILE_LD, PROFILE_HQ };
const FrameRate MAX_V2_FRAMERATE = FR48;

// The fields of SequenceHeader (DataUnit.h:101-150) that EncodeStream sets / DecodeStream reads.
struct SequenceHeader {
  SequenceHeader();
  // DataUnit.cpp:382-433: major_version 2 for HQ, 3 when frameRate > FR48 or bitdepth > 12 or use_v3
  SequenceHeader(Profile profile, int height, int width, ColourFormat chromaFormat, bool interlace, FrameRate frameRate,
                 bool topFieldFirst, int bitdepth, bool use_v3 = false);
  int major_version, minor_version;
  Profile profile;
  int width, height;
  ColourFormat chromaFormat;
  bool interlace;
  FrameRate frameRate;
  unsigned frameRateNumer, frameRateDenom;
  bool topFieldFirst;
  int bitdepth;
  // wire-level view after base video format matching (video_format, DataUnit.cpp:592-786)
  int level, base_video_format;
};

struct Rational { int numerator, denominator; };
Rational rationalise(int numerator, int denominator);   // Utils.cpp:34-50

// transform parameters of one picture (PicturePreamble, DataUnit.h:212-227)
struct PicturePreamble {
  WaveletKernel wavelet_kernel;
  int depth, slices_x, slices_y, slice_prefix, slice_size_scalar;
  Rational slice_bytes;   // LD only
};

// ---- writing ---------------------------------------------------------------------------------------
class StreamWriter {
 public:
  StreamWriter() : prev_(0), major_(0) {}
  // dataunitio::start_sequence (DataUnit.cpp:359-362) + operator<<(SequenceHeader) (:1043-1060)
  void startSequence(std::string& out, const SequenceHeader& hdr);
  // HQWrappedPictureIO (DataUnit.cpp:236-266): parse info + picture number + transform parameters + slice bytes
  void hqPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices, size_t len);
  // the fragmented form of HQWrappedPictureIO (DataUnit.cpp:267-342, parse code 0xEC): one fragment that carries
  // the transform parameters, then fragments of whole slices - a slice is appended to the current fragment unless
  // that would take it over fragmentLength bytes.  slice_off[n_slices + 1] are the byte offsets of the slices
  // inside `slices`.  The sequence header must have been written with major version 3 (fragmentedPictures, :1412-1421).
  void hqFragmentedPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices,
                           const uint32_t* slice_off, int fragmentLength);
  // LDWrappedPictureIO (DataUnit.cpp:125-234, parse codes 0xC8 / 0xCC): as above with the LD transform parameters
  // (slice bytes numerator / denominator from PicturePreamble::slice_bytes)
  void ldPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices, size_t len);
  void ldFragmentedPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices,
                           const uint32_t* slice_off, int fragmentLength);
  // dataunitio::end_sequence (:364-368)
  void endSequence(std::string& out);
 private:
  void parseInfo(std::string& out, unsigned char code, unsigned next);
  void transformParameters(std::string& out, const PicturePreamble& p, bool ld, bool asymFlags);
  void fragmented(std::string& out, unsigned char code, bool ld, unsigned long pictureNumber, const PicturePreamble& p,
                  const uint8_t* slices, const uint32_t* slice_off, int fragmentLength);
  unsigned prev_;
  int major_;
};

// ---- reading ---------------------------------------------------------------------------------------
// picture number + fragment header of an HQ / LD fragment data unit (operator>>(Fragment), DataUnit.cpp:1146-1163)
struct FragmentHeader {
  unsigned long picture_number;
  int fragment_length;       // bytes of fragment data that follow the header
  int n_slices;              // 0: the fragment carries the transform parameters
  int slice_offset_x, slice_offset_y;
};

struct DataUnit {
  DataUnitType type;
  size_t offset;             // of the parse info header in the stream
  unsigned next_parse_offset, prev_parse_offset;
};

class StreamReader {
 public:
  StreamReader(const uint8_t* data, size_t len) : d_(data), n_(len), pos_(0), major_(0) {}
  // dataunitio::synchronise (DataUnit.cpp:1086-1105): position on the next parse info prefix; false at end
  bool synchronise();
  bool atEnd() const { return pos_ >= n_; }
  // operator>>(DataUnit) (:1107-1144); throws std::logic_error on a bad prefix / unknown parse code
  DataUnit readDataUnit();
  SequenceHeader readSequenceHeader();                       // :883-1041, 1203-1312
  // picture number + PicturePreamble (:1314-1410); ld selects the LD parameter set
  PicturePreamble readPictureHeader(bool ld, unsigned long& pictureNumber);
  PicturePreamble readTransformParameters(bool ld);          // the PicturePreamble alone (:1327-1410)
  FragmentHeader readFragmentHeader();                       // :1146-1163 (after the parse info)
  size_t pos() const { return pos_; }
  void setMajorVersion(int m) { major_ = m; }   // normally taken from the sequence header just read
  void seek(size_t p) { pos_ = p; }
  const uint8_t* data() const { return d_; }
  size_t size() const { return n_; }
 private:
  const uint8_t* d_;
  size_t n_, pos_;
  int major_;
};

}  // namespace vc2
#endif
