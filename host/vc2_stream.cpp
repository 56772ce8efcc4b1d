// vc2_stream.cpp - VC-2 stream framing (host only): parse info chain, sequence header, picture headers.
// Behaviour follows /root/reference/src/Library/src/DataUnit.cpp (line numbers cited per function); the
// implementation is table driven and works on byte strings instead of iostream state.
#include <cstring>
#include <stdexcept>
#include <string>

#include "vc2/DataUnit.h"

namespace vc2 {

namespace {

// ---- MSB-first bit string (VLC.cpp:119-180 without the bounded mode, which framing never uses) ----
struct BitSink {
  std::string& s;
  unsigned acc;
  int n;
  explicit BitSink(std::string& out) : s(out), acc(0), n(0) {}
  void bit(bool b) {
    acc = (acc << 1) | (b ? 1u : 0u);
    if (++n == 8) { s.push_back((char)acc); acc = 0; n = 0; }
  }
  // unsigned interleaved exp-Golomb (VLC.cpp:21-52): m = v+1, k = floor(log2 m): 0 b(k-1) 0 b(k-2) ... 0 b0 1
  void uvlc(unsigned v) {
    const unsigned m = v + 1;
    int k = 0;
    while ((m >> (k + 1)) != 0) ++k;
    for (int i = k - 1; i >= 0; --i) { bit(false); bit((m >> i) & 1u); }
    bit(true);
  }
  void align() { while (n) bit(false); }   // vlc::align pads with zero bits (VLC.cpp:229-236)
};

struct BitSource {
  const uint8_t* d;
  size_t n, pos;   // pos in bits
  BitSource(const uint8_t* data, size_t len, size_t byte_pos) : d(data), n(len), pos(byte_pos * 8) {}
  bool bit() {
    if ((pos >> 3) >= n) throw std::logic_error("Stream Error: unexpected end of data unit");
    const bool b = (d[pos >> 3] >> (7 - (pos & 7))) & 1;
    ++pos;
    return b;
  }
  unsigned uvlc() {   // VLC.cpp:283-295
    unsigned m = 1;
    while (!bit()) m = (m << 1) | (bit() ? 1u : 0u);
    return m - 1;
  }
  void align() { pos = (pos + 7) & ~(size_t)7; }
  size_t byte_pos() const { return pos >> 3; }
};

void be(std::string& out, unsigned v, int bytes) {
  for (int i = bytes - 1; i >= 0; --i) out.push_back((char)((v >> (8 * i)) & 0xFF));
}

// base video formats 1..22 (DataUnit.cpp:438-466; index 0 is the 640x480 fallback): the fields the matcher reads
struct BaseFormat { int h, w; ColourFormat cf; bool interlace; FrameRate fr; bool tff; int bits; };
const BaseFormat kBase[23] = {
    {480, 640, CF420, false, FR24000_1001, false, 8},
    {120, 176, CF420, false, FR15000_1001, false, 8},   {144, 176, CF420, false, FR25_2, true, 8},
    {240, 352, CF420, false, FR15000_1001, false, 8},   {288, 352, CF420, false, FR25_2, true, 8},
    {480, 704, CF420, false, FR15000_1001, false, 8},   {576, 704, CF420, false, FR25_2, true, 8},
    {480, 720, CF422, true, FR30000_1001, false, 10},   {576, 720, CF422, true, FR25, true, 10},
    {720, 1280, CF422, false, FR60000_1001, true, 10},  {720, 1280, CF422, false, FR50, true, 10},
    {1080, 1920, CF422, true, FR30000_1001, true, 10},  {1080, 1920, CF422, true, FR25, true, 10},
    {1080, 1920, CF422, false, FR60000_1001, true, 10}, {1080, 1920, CF422, false, FR50, true, 10},
    {1080, 2048, CF444, false, FR24, true, 12},         {2160, 4096, CF444, false, FR24, true, 12},
    {2160, 3840, CF422, false, FR60000_1001, true, 10}, {2160, 3840, CF422, false, FR50, true, 10},
    {4320, 7680, CF422, false, FR60000_1001, true, 10}, {4320, 7680, CF422, false, FR50, true, 10},
    {1080, 1920, CF422, false, FR24000_1001, true, 10}, {486, 720, CF422, true, FR30000_1001, false, 10},
};

// the wire-level video format (video_format, DataUnit.cpp:592-786) for a header whose optional fields are unset
struct WireFormat {
  int level = 0, base = 0;
  bool custom_dims = false; int width = 0, height = 0;
  bool custom_cf = false; int cf = 0;
  bool custom_scan = false; int source_sampling = 0;
  bool custom_fr = false; int fr = 0; unsigned fr_num = 0, fr_den = 0;
  bool custom_clean = false; int clean_w = 0, clean_h = 0, left = 0, top = 0;
  bool custom_range = false; int range_index = 0;
};

bool matches_all(const SequenceHeader& f, int i) {   // PictureFormatMatches(fmt, index), :482-502
  const BaseFormat& b = kBase[i];
  return f.width == b.w && f.height == b.h && f.chromaFormat == b.cf && f.frameRate == b.fr && f.bitdepth == b.bits &&
         f.interlace == b.interlace && f.topFieldFirst == b.tff;
}
bool matches(const SequenceHeader& f, int w, int h, ColourFormat cf, FrameRate r, int bd, bool tff) {   // :468-480
  return f.width == w && f.height == h && f.chromaFormat == cf && f.frameRate == r && f.bitdepth == bd && f.topFieldFirst == tff;
}

WireFormat to_wire(const SequenceHeader& f) {
  WireFormat v;
  auto pick = [&](int base, int level) { v.base = base; v.level = level; };
  if (f.interlace) {   // :640-662
    if (matches_all(f, 7)) pick(7, 2);
    else if (matches_all(f, 8)) pick(8, 2);
    else if (matches_all(f, 22)) pick(22, 2);
    else if (f.chromaFormat == CF422 && f.width == 720 && f.height >= 480 && f.height <= 486 && f.frameRate == FR30000_1001 &&
             f.bitdepth == 10) {
      pick(7, 2);
      v.custom_dims = true; v.width = f.width; v.height = f.height;
    }
    else if (matches_all(f, 11)) pick(11, 3);
    else if (matches_all(f, 12)) pick(12, 3);
  } else {             // :663-701
    auto scan = [&](int base, int level) { pick(base, level); v.custom_scan = true; v.source_sampling = 0; };
    if (matches_all(f, 1)) pick(1, 1);
    else if (matches_all(f, 2)) pick(2, 1);
    else if (matches_all(f, 3)) pick(3, 1);
    else if (matches_all(f, 4)) pick(4, 1);
    else if (matches_all(f, 5)) pick(5, 1);
    else if (matches_all(f, 6)) pick(6, 1);
    else if (matches(f, 720, 480, CF422, FR30000_1001, 10, false)) scan(7, 2);
    else if (matches(f, 720, 576, CF422, FR25, 10, true)) scan(8, 2);
    else if (matches(f, 720, 486, CF422, FR30000_1001, 10, false)) scan(22, 2);
    else if (matches_all(f, 9)) pick(9, 3);
    else if (matches_all(f, 10)) pick(10, 3);
    else if (matches(f, 1920, 1080, CF422, FR30000_1001, 10, true)) scan(11, 3);
    else if (matches(f, 1920, 1080, CF422, FR25, 10, true)) scan(12, 3);
    else if (matches_all(f, 13)) pick(13, 3);
    else if (matches_all(f, 14)) pick(14, 3);
    else if (matches_all(f, 21)) pick(21, 3);
    else if (matches_all(f, 15)) pick(15, 4);
    else if (matches(f, 2048, 1080, CF444, FR48, 12, true)) { pick(15, 4); v.custom_fr = true; v.fr = FR48; }
    else if (matches_all(f, 16)) pick(16, 5);
    else if (matches_all(f, 17)) pick(17, 6);
    else if (matches_all(f, 18)) pick(18, 6);
    else if (matches_all(f, 19)) pick(19, 7);
    else if (matches_all(f, 20)) pick(20, 7);
  }
  if (v.base != 0) return v;

  // no exact match (:703-785): the base format with the fewest differing fields among those with the same field
  // order, first one wins; everything that differs becomes a custom override
  v.level = 0;
  int best = 999;
  for (int i = 1; i <= 22; ++i) {
    const BaseFormat& b = kBase[i];
    if (f.topFieldFirst != b.tff) continue;   // CheckMatch returns -1 (:532-534)
    const int diff = (f.width != b.w) + (f.height != b.h) + (f.chromaFormat != b.cf) + (f.frameRate != b.fr) + (f.bitdepth != b.bits) +
                     (f.interlace != b.interlace);
    if (diff < best) { v.base = i; best = diff; }
  }
  const BaseFormat& b = kBase[v.base];
  if (f.interlace != b.interlace) { v.custom_scan = true; v.source_sampling = f.interlace ? 1 : 0; }
  if (f.width != b.w || f.height != b.h) { v.custom_dims = true; v.width = f.width; v.height = f.height; }
  if (f.chromaFormat != b.cf) { v.custom_cf = true; v.cf = (int)f.chromaFormat; }
  if (f.frameRate != b.fr) {
    v.custom_fr = true; v.fr = (int)f.frameRate;
    if (f.frameRate == FR0) { v.fr_num = f.frameRateNumer; v.fr_den = f.frameRateDenom; }
  }
  if (f.bitdepth != b.bits) {
    v.custom_range = true;
    switch (f.bitdepth) {
      case 8: v.range_index = 1; break;
      case 10: v.range_index = 3; break;
      case 12: v.range_index = 4; break;
      case 16: v.range_index = 7; break;
      default: throw std::logic_error("DataUnitIO: invalid bit depth");
    }
  }
  if (v.custom_dims) {   // no clean area given: it becomes the whole picture (:754-765)
    v.custom_clean = true; v.clean_w = v.width; v.clean_h = v.height; v.left = 0; v.top = 0;
  }
  return v;
}

}  // namespace

SequenceHeader::SequenceHeader()
    : major_version(1), minor_version(0), profile(PROFILE_UNKNOWN), width(0), height(0), chromaFormat(CF444), interlace(false),
      frameRate(FR0), frameRateNumer(0), frameRateDenom(0), topFieldFirst(false), bitdepth(0), level(0), base_video_format(0) {}

SequenceHeader::SequenceHeader(Profile p, int h, int w, ColourFormat cf, bool il, FrameRate fr, bool tff, int bits, bool use_v3)
    : major_version(1), minor_version(0), profile(p), width(w), height(h), chromaFormat(cf), interlace(il), frameRate(fr),
      frameRateNumer(0), frameRateDenom(0), topFieldFirst(tff), bitdepth(bits), level(0), base_video_format(0) {
  if (p == PROFILE_HQ) major_version = 2;                                     // DataUnit.cpp:425-427
  if (use_v3 || fr > MAX_V2_FRAMERATE || bits > 12) major_version = 3;         // :428-432
}

Rational rationalise(int numerator, int denominator) {
  int a = numerator < 0 ? -numerator : numerator, b = denominator < 0 ? -denominator : denominator;
  while (b) { const int t = a % b; a = b; b = t; }
  Rational r;
  r.numerator = a ? numerator / a : numerator;
  r.denominator = a ? denominator / a : denominator;
  return r;
}

// ---- writer ------------------------------------------------------------------------------------------
void StreamWriter::parseInfo(std::string& out, unsigned char code, unsigned next) {   // DataUnit.cpp:112-123
  out.push_back(0x42); out.push_back(0x42); out.push_back(0x43); out.push_back(0x44);
  out.push_back((char)code);
  be(out, next, 4);
  be(out, prev_, 4);
  prev_ = next;
}

void StreamWriter::startSequence(std::string& out, const SequenceHeader& hdr) {
  prev_ = 0;   // start_sequence (:359-362)
  const WireFormat v = to_wire(hdr);
  std::string body;
  BitSink b(body);
  major_ = hdr.major_version;
  b.uvlc(hdr.major_version); b.uvlc(hdr.minor_version);
  b.uvlc(hdr.profile == PROFILE_HQ ? 3 : 0);   // :627-638
  b.uvlc(v.level); b.uvlc(v.base);
  b.bit(v.custom_dims); if (v.custom_dims) { b.uvlc(v.width); b.uvlc(v.height); }
  b.bit(v.custom_cf); if (v.custom_cf) b.uvlc(v.cf);
  b.bit(v.custom_scan); if (v.custom_scan) b.uvlc(v.source_sampling);
  b.bit(v.custom_fr);
  if (v.custom_fr) { b.uvlc(v.fr); if (v.fr == FR0) { b.uvlc(v.fr_num); b.uvlc(v.fr_den); } }
  b.bit(false);                                 // custom_pixel_aspect_ratio_flag: never set by EncodeStream
  b.bit(v.custom_clean); if (v.custom_clean) { b.uvlc(v.clean_w); b.uvlc(v.clean_h); b.uvlc(v.left); b.uvlc(v.top); }
  b.bit(v.custom_range); if (v.custom_range) b.uvlc(v.range_index);
  b.bit(false);                                 // custom_color_spec_flag
  b.uvlc(v.source_sampling);                    // picture coding mode follows the source sampling (:872-877)
  b.align();
  parseInfo(out, 0x00, (unsigned)body.size() + 13);
  out += body;
}

void StreamWriter::transformParameters(std::string& out, const PicturePreamble& p, bool ld, bool asymFlags) {
  BitSink b(out);
  b.uvlc((unsigned)p.wavelet_kernel); b.uvlc(p.depth);
  if (asymFlags) { b.bit(false); b.bit(false); }   // asym_transform_index_flag, asym_transform_flag
  b.uvlc(p.slices_x); b.uvlc(p.slices_y);
  if (ld) { b.uvlc(p.slice_bytes.numerator); b.uvlc(p.slice_bytes.denominator); }
  else { b.uvlc(p.slice_prefix); b.uvlc(p.slice_size_scalar); }
  b.bit(false);                                      // no custom quantisation matrix
  b.align();
}

void StreamWriter::hqPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices, size_t len) {
  std::string hdr;
  be(hdr, (unsigned)pictureNumber, 4);
  transformParameters(hdr, p, false, major_ >= 3);   // :249-252
  parseInfo(out, 0xE8, (unsigned)(hdr.size() + len) + 13);
  out += hdr;
  if (slices) out.append(reinterpret_cast<const char*>(slices), len);   // NULL: headers only, the caller writes the slice bytes itself
}

void StreamWriter::ldPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices, size_t len) {
  std::string hdr;
  be(hdr, (unsigned)pictureNumber, 4);
  transformParameters(hdr, p, true, major_ >= 3);    // :138-141
  parseInfo(out, 0xC8, (unsigned)(hdr.size() + len) + 13);
  out += hdr;
  if (slices) out.append(reinterpret_cast<const char*>(slices), len);
}

void StreamWriter::fragmented(std::string& out, unsigned char code, bool ld, unsigned long pictureNumber, const PicturePreamble& p,
                              const uint8_t* slices, const uint32_t* slice_off, int fragmentLength) {
  {   // the parameter fragment (:155-178, :268-293): slice count 0; the asymmetric transform flags are always present here
    std::string prm;
    transformParameters(prm, p, ld, true);
    parseInfo(out, code, (unsigned)prm.size() + 8 + 13);
    be(out, (unsigned)pictureNumber, 4);
    be(out, (unsigned)prm.size(), 2);
    be(out, 0, 2);
    out += prm;
  }
  auto emit = [&](int first, int count) {                // :198-205, :224-231, :306-313, :331-338
    const unsigned bytes = slice_off[first + count] - slice_off[first];
    parseInfo(out, code, bytes + 12 + 13);
    be(out, (unsigned)pictureNumber, 4);
    be(out, bytes, 2);
    be(out, (unsigned)count, 2);
    be(out, (unsigned)(first % p.slices_x), 2);
    be(out, (unsigned)(first / p.slices_x), 2);
    out.append(reinterpret_cast<const char*>(slices + slice_off[first]), bytes);
  };
  const int n = p.slices_x * p.slices_y;
  int first = 0, count = 0;
  for (int s = 0; s < n; ++s) {
    const unsigned have = slice_off[s] - slice_off[first], sz = slice_off[s + 1] - slice_off[s];
    if (count > 0 && (int)(have + sz) > fragmentLength) { emit(first, count); first = s; count = 0; }
    ++count;
  }
  emit(first, count);
}

void StreamWriter::hqFragmentedPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices,
                                       const uint32_t* slice_off, int fragmentLength) {
  fragmented(out, 0xEC, false, pictureNumber, p, slices, slice_off, fragmentLength);
}

void StreamWriter::ldFragmentedPicture(std::string& out, unsigned long pictureNumber, const PicturePreamble& p, const uint8_t* slices,
                                       const uint32_t* slice_off, int fragmentLength) {
  fragmented(out, 0xCC, true, pictureNumber, p, slices, slice_off, fragmentLength);
}

void StreamWriter::endSequence(std::string& out) {   // :364-368
  parseInfo(out, 0x10, 0);
  prev_ = 0;
}

// ---- reader ------------------------------------------------------------------------------------------
bool StreamReader::synchronise() {
  while (pos_ + 4 <= n_) {
    if (d_[pos_] == 0x42 && d_[pos_ + 1] == 0x42 && d_[pos_ + 2] == 0x43 && d_[pos_ + 3] == 0x44) return true;
    ++pos_;
  }
  pos_ = n_;
  return false;
}

DataUnit StreamReader::readDataUnit() {
  if (pos_ + 13 > n_) throw std::logic_error("Stream Error: truncated parse info header");
  const uint8_t* p = d_ + pos_;
  if (p[0] != 0x42 || p[1] != 0x42 || p[2] != 0x43 || p[3] != 0x44)
    throw std::logic_error("Read bytes do not match expected parse_info_header.");
  DataUnit du;
  du.offset = pos_;
  switch (p[4]) {
    case 0x00: du.type = SEQUENCE_HEADER; break;
    case 0x10: du.type = END_OF_SEQUENCE; break;
    case 0x20: du.type = AUXILIARY_DATA; break;
    case 0x30: du.type = PADDING_DATA; break;
    case 0xC8: du.type = LD_PICTURE; break;
    case 0xE8: du.type = HQ_PICTURE; break;
    case 0xCC: du.type = LD_FRAGMENT; break;
    case 0xEC: du.type = HQ_FRAGMENT; break;
    default: throw std::logic_error("Stream Error: Unknown data unit type.");
  }
  du.next_parse_offset = (unsigned)p[5] << 24 | (unsigned)p[6] << 16 | (unsigned)p[7] << 8 | p[8];
  du.prev_parse_offset = (unsigned)p[9] << 24 | (unsigned)p[10] << 16 | (unsigned)p[11] << 8 | p[12];
  pos_ += 13;
  return du;
}

SequenceHeader StreamReader::readSequenceHeader() {
  BitSource b(d_, n_, pos_);
  SequenceHeader h;
  h.major_version = (int)b.uvlc();
  h.minor_version = (int)b.uvlc();
  const unsigned profile = b.uvlc();
  h.level = (int)b.uvlc();
  h.base_video_format = (int)b.uvlc();
  if (h.base_video_format > 22) throw std::logic_error("DataUnitIO: unknown base video format");
  const BaseFormat& base = kBase[h.base_video_format];   // copy_video_fmt_to_hdr (:1203-1226)
  h.width = base.w; h.height = base.h; h.chromaFormat = base.cf; h.interlace = base.interlace; h.frameRate = base.fr;
  h.topFieldFirst = base.tff; h.bitdepth = base.bits;
  h.profile = profile == 0 ? PROFILE_LD : profile == 3 ? PROFILE_HQ : PROFILE_UNKNOWN;
  if (b.bit()) { h.width = (int)b.uvlc(); h.height = (int)b.uvlc(); }
  if (b.bit()) {
    const unsigned cf = b.uvlc();
    if (cf > 2) throw std::logic_error("DataUnitIO: Invalid Frame Rate on Input: invalid colour format");
    h.chromaFormat = (ColourFormat)cf;
  }
  if (b.bit()) h.interlace = b.uvlc() != 0;
  if (b.bit()) {
    const unsigned fr = b.uvlc();
    if (fr > FR120) throw std::logic_error("DataUnitIO: Invalid Frame Rate on Input");
    h.frameRate = (FrameRate)fr;
    if (h.frameRate == FR0) { h.frameRateNumer = b.uvlc(); h.frameRateDenom = b.uvlc(); }
    if (h.frameRate > MAX_V2_FRAMERATE && h.major_version < 3) h.major_version = 3;
  }
  if (b.bit()) { if (b.uvlc() == 0) { b.uvlc(); b.uvlc(); } }     // pixel aspect ratio (+ custom ratio)
  if (b.bit()) { b.uvlc(); b.uvlc(); b.uvlc(); b.uvlc(); }        // clean area
  if (b.bit()) {                                                  // signal range (:1273-1297)
    const unsigned idx = b.uvlc();
    static const int bits[9] = {0, 8, 8, 10, 12, 10, 12, 16, 16};
    if (idx <= 8) h.bitdepth = bits[idx];
    if (idx == 0) { b.uvlc(); b.uvlc(); b.uvlc(); b.uvlc(); }
    if (idx > 4 && h.major_version < 3) h.major_version = 3;
  }
  if (b.bit()) {                                                  // colour spec
    if (b.uvlc() == 0) {
      if (b.bit()) b.uvlc();
      if (b.bit()) b.uvlc();
      if (b.bit()) b.uvlc();
    }
  }
  b.uvlc();   // picture coding mode
  b.align();
  pos_ = b.byte_pos();
  major_ = h.major_version;
  return h;
}

FragmentHeader StreamReader::readFragmentHeader() {
  if (pos_ + 8 > n_) throw std::logic_error("Stream Error: truncated fragment header");
  const uint8_t* q = d_ + pos_;
  FragmentHeader f;
  f.picture_number = (unsigned long)q[0] << 24 | (unsigned long)q[1] << 16 | (unsigned long)q[2] << 8 | q[3];
  f.fragment_length = q[4] << 8 | q[5];
  f.n_slices = q[6] << 8 | q[7];
  f.slice_offset_x = f.slice_offset_y = 0;
  pos_ += 8;
  if (f.n_slices != 0) {
    if (pos_ + 4 > n_) throw std::logic_error("Stream Error: truncated fragment header");
    f.slice_offset_x = q[8] << 8 | q[9];
    f.slice_offset_y = q[10] << 8 | q[11];
    pos_ += 4;
  }
  return f;
}

PicturePreamble StreamReader::readPictureHeader(bool ld, unsigned long& pictureNumber) {
  if (pos_ + 4 > n_) throw std::logic_error("Stream Error: truncated picture header");
  pictureNumber = (unsigned long)d_[pos_] << 24 | (unsigned long)d_[pos_ + 1] << 16 | (unsigned long)d_[pos_ + 2] << 8 | d_[pos_ + 3];
  pos_ += 4;
  return readTransformParameters(ld);
}

PicturePreamble StreamReader::readTransformParameters(bool ld) {
  BitSource b(d_, n_, pos_);
  PicturePreamble p;
  const unsigned wi = b.uvlc();
  p.wavelet_kernel = wi <= 6 ? (WaveletKernel)wi : NullKernel;
  p.depth = (int)b.uvlc();
  if (major_ >= 3) {   // asymmetric transform flags are parsed and ignored (:1342-1367)
    if (b.bit()) b.uvlc();
    if (b.bit()) b.uvlc();
  }
  p.slices_x = (int)b.uvlc();
  p.slices_y = (int)b.uvlc();
  if (ld) {
    const int num = (int)b.uvlc(), den = (int)b.uvlc();
    p.slice_prefix = 0; p.slice_size_scalar = 0;
    p.slice_bytes = rationalise(num, den);
  } else {
    p.slice_prefix = (int)b.uvlc();
    p.slice_size_scalar = (int)b.uvlc();
    p.slice_bytes = rationalise(0, 1);
  }
  if (b.bit()) throw std::logic_error("DataUnitIO: Custom Quantisation Matrix flag not supported");
  b.align();
  pos_ = b.byte_pos();
  return p;
}

}  // namespace vc2

// ---- C-ABI (include/vc2_host.h) ----------------------------------------------------------------------
#include "vc2_host.h"

namespace {
vc2::ColourFormat cf_of(int c) { return c == 0 ? vc2::CF444 : c == 1 ? vc2::CF422 : vc2::CF420; }
}

extern "C" int vc2host_sequence_header(int profile_hq, int height, int width, int cf, int interlace, int frame_rate, int tff, int bitdepth,
                                       uint8_t* out, int cap) {
  try {
    vc2::StreamWriter w;
    std::string s;
    w.startSequence(s, vc2::SequenceHeader(profile_hq ? vc2::PROFILE_HQ : vc2::PROFILE_LD, height, width, cf_of(cf), interlace != 0,
                                           (vc2::FrameRate)frame_rate, tff != 0, bitdepth));
    if ((int)s.size() > cap) return -2;
    memcpy(out, s.data(), s.size());
    return (int)s.size();
  } catch (const std::exception&) { return -1; }
}

extern "C" long long vc2host_wrap_hq_stream(int height, int width, int cf, int frame_rate, int tff, int bitdepth, int kernel, int depth,
                                            int slices_x, int slices_y, int prefix, int scalar, int n, const uint8_t* const* payloads,
                                            const size_t* payload_len, uint8_t* out, size_t cap) {
  try {
    vc2::StreamWriter w;
    std::string s;
    w.startSequence(s, vc2::SequenceHeader(vc2::PROFILE_HQ, height, width, cf_of(cf), false, (vc2::FrameRate)frame_rate, tff != 0, bitdepth));
    vc2::PicturePreamble p;
    p.wavelet_kernel = (vc2::WaveletKernel)kernel; p.depth = depth; p.slices_x = slices_x; p.slices_y = slices_y;
    p.slice_prefix = prefix; p.slice_size_scalar = scalar; p.slice_bytes = vc2::rationalise(0, 1);
    for (int i = 0; i < n; ++i) w.hqPicture(s, (unsigned long)i, p, payloads[i], payload_len[i]);
    w.endSequence(s);
    if (s.size() > cap) return -2;
    memcpy(out, s.data(), s.size());
    return (long long)s.size();
  } catch (const std::exception&) { return -1; }
}

extern "C" int vc2host_parse_units(const uint8_t* data, size_t len, int max_units, int64_t* units) {
  try {
    vc2::StreamReader r(data, len);
    int n = 0;
    if (!r.synchronise()) return 0;
    while (!r.atEnd() && n < max_units) {
      const vc2::DataUnit du = r.readDataUnit();
      units[4 * n] = data[du.offset + 4]; units[4 * n + 1] = (int64_t)du.offset;
      units[4 * n + 2] = du.next_parse_offset; units[4 * n + 3] = du.prev_parse_offset;
      ++n;
      if (du.next_parse_offset == 0) break;   // end of sequence (or an unterminated unit): nothing to chain to
      r.seek(du.offset + du.next_parse_offset);
    }
    return n;
  } catch (const std::exception&) { return -1; }
}

extern "C" int vc2host_read_sequence_header(const uint8_t* data, size_t len, size_t offset, int32_t* f) {
  try {
    vc2::StreamReader r(data, len);
    r.seek(offset);
    const vc2::SequenceHeader h = r.readSequenceHeader();
    f[0] = h.major_version; f[1] = h.profile == vc2::PROFILE_HQ ? 3 : 0; f[2] = h.height; f[3] = h.width; f[4] = (int)h.chromaFormat;
    f[5] = h.interlace; f[6] = (int)h.frameRate; f[7] = h.topFieldFirst; f[8] = h.bitdepth; f[9] = (int)(r.pos() - offset);
    return 0;
  } catch (const std::exception&) { return -1; }
}

extern "C" int vc2host_read_picture_header(const uint8_t* data, size_t len, size_t offset, int ld, int major_version, int64_t* f) {
  try {
    vc2::StreamReader r(data, len);
    r.setMajorVersion(major_version);
    r.seek(offset);
    unsigned long picnum = 0;
    const vc2::PicturePreamble p = r.readPictureHeader(ld != 0, picnum);
    f[0] = (int64_t)picnum; f[1] = (int)p.wavelet_kernel; f[2] = p.depth; f[3] = p.slices_x; f[4] = p.slices_y;
    f[5] = ld ? p.slice_bytes.numerator : p.slice_prefix; f[6] = ld ? p.slice_bytes.denominator : p.slice_size_scalar;
    f[7] = (int64_t)(r.pos() - offset); f[8] = 0;
    return 0;
  } catch (const std::exception&) { return -1; }
}
