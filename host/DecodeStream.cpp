// DecodeStream - drop-in for the reference command line (src/DecodeStream/DecodeStream.cpp, DecodeParams.cpp)
// on the B200 hot path: same flags, same output bytes.  The stream is split into data units on the host
// (parse-info chain), HQ / LD pictures are decoded in batches by the fused CUDA codec
// (vc2_codec_decode_host); with --gpus N consecutive batches go to different GPUs and the pictures are
// written in stream order.
//
// Fragmented pictures (DecodeStream.cpp:614-977) are reassembled on the host: the slices of the fragments of a
// picture, in slice order, are exactly the slice payload of the unfragmented picture.  Interlaced streams decode as
// field pictures of half the height; two consecutive pictures are woven into one output frame (Frame.cpp:62-110).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <cstring>
#include <deque>
#include <iterator>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <chrono>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "cmdline.h"
#include "pipeline.h"
#include "vc2/Codec.h"
#include "vc2/DataUnit.h"
#include "vc2/Quantisation.h"
#include "vc2/Slices.h"
#include "vc2/WaveletTransform.h"

using namespace vc2;
using std::clog;
using std::endl;

namespace {

enum Output { TRANSFORM, QUANTISED, INDICES, DECODED };

// DecodeStream.cpp:312: bytes of a whole LD picture from the slice-bytes ratio (the header is untrusted: no division by zero)
static int ld_picture_bytes(const PicturePreamble& pre) {
  if (pre.slice_bytes.denominator == 0) throw std::logic_error("Stream Error: slice bytes denominator is zero");
  if (pre.slices_y < 1 || pre.slices_x < 1) throw std::logic_error("Stream Error: slice counts must be positive");
  return (int)((long long)pre.slice_bytes.numerator * pre.slices_y * pre.slices_x / pre.slice_bytes.denominator);
}

// VC2_CLI_TIMING=1: where the wall-clock time of a run goes (standard error; measurement only)
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct PictureUnit {
  const uint8_t* data;
  size_t len;
};

struct Config {   // everything the codec geometry depends on
  bool ld = false;
  int height = 0, width = 0, bits = 0;
  ColourFormat cf = CF_UNSET;
  PicturePreamble pre;
  int compressedBytes = 0;
  bool interlace = false, tff = true;   // height is the PICTURE height (a field when interlaced)
  bool same(const Config& o) const {
    return interlace == o.interlace && tff == o.tff && ld == o.ld && height == o.height && width == o.width && bits == o.bits && cf == o.cf &&
           pre.wavelet_kernel == o.pre.wavelet_kernel && pre.depth == o.pre.depth && pre.slices_x == o.pre.slices_x &&
           pre.slices_y == o.pre.slices_y && pre.slice_prefix == o.pre.slice_prefix && pre.slice_size_scalar == o.pre.slice_size_scalar &&
           compressedBytes == o.compressedBytes;
  }
};

void write_be32_plane(std::ostream& out, const int* v, size_t n) {
  std::string buf(n * 4, '\0');
  for (size_t i = 0; i < n; ++i) {
    const uint32_t w = (uint32_t)v[i];
    buf[4 * i] = (char)(w >> 24); buf[4 * i + 1] = (char)(w >> 16); buf[4 * i + 2] = (char)(w >> 8); buf[4 * i + 3] = (char)w;
  }
  out.write(buf.data(), (std::streamsize)buf.size());
}

// end of an HQ picture's slice data when the parse info carries no next offset: walk the length bytes
size_t hq_payload_length(const uint8_t* p, size_t avail, int nslices, int prefix, int scalar) {
  std::vector<uint32_t> off(nslices + 1);
  if (vc2_hq_index_slices(p, avail, nslices, prefix, scalar, off.data()) != VC2_OK) throw std::logic_error("Stream Error: HQ picture runs past the end of the stream");
  return off[nslices];
}

class Decoder {
 public:
  // decoded pictures go through the positional writer `pw` (large pieces, several threads), the tap outputs through `out`
  Decoder(std::ostream& out, vc2cli::PositionalWriter* pw, Output output, int gpus, int batch, bool verbose)
      : out_(out), pw_(pw), output_(output), G_(gpus), B_(batch), verbose_(verbose), frames_(0) {}

  // a payload this object keeps alive until it has been decoded (reassembled fragments)
  const uint8_t* keep(std::vector<uint8_t>&& bytes) {
    owned_.emplace_back(std::move(bytes));
    return owned_.back().data();
  }

  void add(const Config& c, const PictureUnit& u) {
    if (!pending_.empty() && !cfg_.same(c)) flush();
    if (pending_.empty() && (codecs_.empty() || !cfg_.same(c))) open(c);
    pending_.push_back(u);
    if ((int)pending_.size() == G_ * B_) flush();
  }

  void flush() {
    if (pending_.empty()) return;
    if (output_ != DECODED) { taps(); pending_.clear(); owned_.clear(); return; }
    const int n = (int)pending_.size();
    const double t0 = now_s();
    std::vector<std::string> errors(G_);
    std::vector<std::thread> th;
    for (int g = 0; g < G_; ++g) {
      const int first = g * B_, cnt = std::min(B_, n - first);
      if (cnt <= 0) break;
      th.emplace_back([&, g, first, cnt]() {
        if (G_ > 1) vc2_bind_thread_to_device(g);
        try {
          std::vector<const uint8_t*> pay(cnt);
          std::vector<size_t> len(cnt);
          std::vector<void*> pics(cnt);
          for (int i = 0; i < cnt; ++i) { pay[i] = pending_[first + i].data; len[i] = pending_[first + i].len; pics[i] = recon_[cur_][first + i].data(); }
          codecs_[g]->decode(cnt, pay.data(), len.data(), pics.data());
        } catch (const std::exception& e) { errors[g] = e.what(); }
      });
    }
    for (auto& t : th) t.join();
    t_decode_ += now_s() - t0;
    for (int g = 0; g < G_; ++g) if (!errors[g].empty()) throw std::logic_error(errors[g]);
    // the pictures of this batch are written by a second thread while the next batch is parsed and decoded into the other
    // set of buffers (the reference decodes and writes one picture at a time, DecodeStream.cpp:512-605)
    finishWrite();
    const int set = cur_;
    const Config cfg = cfg_;
    writer_ = std::thread([this, set, n, cfg]() {
      try { writeSet(set, n, cfg); } catch (const std::exception& e) { writeError_ = e.what(); }
    });
    cur_ ^= 1;
    pending_.clear();
    owned_.clear();
  }

  // wait for the batch being written; called before its buffers are reused, on a geometry change and at the end
  void finishWrite() {
    if (writer_.joinable()) writer_.join();
    if (!writeError_.empty()) { const std::string e = writeError_; writeError_.clear(); throw std::runtime_error(e); }
  }
  ~Decoder() { if (writer_.joinable()) writer_.join(); }

  long frames() { finishWrite(); return frames_; }
  double decodeSeconds() const { return t_decode_; }
  double writeSeconds() const { return t_write_; }

 private:
  void writeSet(int set, int n, const Config& cfg) {
    const double t0 = now_s();
    const size_t bytes = recon_[set][0].size();
    std::vector<vc2cli::PositionalWriter::Piece> pieces;
    for (int i = 0; i < n; ++i) {
      const uint8_t* pic = recon_[set][i].data();
      if (cfg.interlace) {
        // DecodeStream.cpp:566-583: the first picture of a pair is the first field, the second completes the frame
        if (field_.empty()) { field_.assign(pic, pic + bytes); continue; }
        weave(cfg, field_.data(), pic, bytes);
        field_.clear();
        pw_->append({{frame_.data(), frame_.size()}});
      } else {
        pieces.push_back({pic, bytes});
      }
      if (verbose_) clog << "Decoded frame number " << frames_ << endl;
      ++frames_;
    }
    if (!pieces.empty()) pw_->append(pieces);
    if (!pw_->ok()) throw std::runtime_error("Failed to write output file");
    t_write_ += now_s() - t0;
  }

  // two field pictures -> the rows of one frame (Frame::firstField / secondField, Frame.cpp:40-110)
  void weave(const Config& cfg, const uint8_t* first, const uint8_t* second, size_t fieldBytes) {
    const PictureFormat ff(cfg.height, cfg.width, cfg.cf);
    const int bytes = cfg.bits == 8 ? 1 : 2;
    const int h[3] = {ff.lumaHeight(), ff.chromaHeight(), ff.chromaHeight()};
    const size_t w[3] = {(size_t)ff.lumaWidth() * bytes, (size_t)ff.chromaWidth() * bytes, (size_t)ff.chromaWidth() * bytes};
    frame_.resize(2 * fieldBytes);
    const uint8_t* top = cfg.tff ? first : second;
    const uint8_t* bot = cfg.tff ? second : first;
    uint8_t* dst = frame_.data();
    for (int c = 0; c < 3; ++c)
      for (int y = 0; y < h[c]; ++y) {
        memcpy(dst, top, w[c]); top += w[c]; dst += w[c];
        memcpy(dst, bot, w[c]); bot += w[c]; dst += w[c];
      }
  }

  void open(const Config& c) {
    finishWrite();   // the writer thread still reads the old geometry's buffers
    if (!cfg_.same(c)) field_.clear();
    cfg_ = c;
    codecs_.clear();
    vc2_codec_params cp;
    // untrusted header fields are checked before anything is computed from them
    const int depth = c.pre.depth;
    if (depth < 1 || depth > 6) throw std::logic_error("Stream Error: wavelet depth outside 1..6");
    if (c.pre.slices_y < 1 || c.pre.slices_x < 1) throw std::logic_error("Stream Error: slice counts must be positive");
    if (c.ld && c.pre.slice_bytes.denominator == 0) throw std::logic_error("Stream Error: slice bytes denominator is zero");
    const PictureFormat f(c.height, c.width, c.cf);
    // the decoder takes the slice counts as they are in the stream: sliceSizeIsValid (Utils.cpp) is an encoder-side rule, the
    // codec itself only needs every slice to be a whole number of 2^depth cells in both components (geom_ok in cabi.cu)
    cp.geom.luma_h = f.lumaHeight(); cp.geom.luma_w = f.lumaWidth();
    cp.geom.chroma_h = f.chromaHeight(); cp.geom.chroma_w = f.chromaWidth();
    cp.geom.kernel = (int)c.pre.wavelet_kernel; cp.geom.depth = depth;
    cp.geom.slices_y = c.pre.slices_y; cp.geom.slices_x = c.pre.slices_x;
    cp.geom.prefix = c.ld ? 0 : c.pre.slice_prefix; cp.geom.scalar = c.ld ? 1 : c.pre.slice_size_scalar;
    if (cp.geom.kernel < 0 || cp.geom.kernel > 6 || cp.geom.prefix < 0 || cp.geom.scalar < 1)
      throw std::logic_error("Stream Error: unsupported picture / slice geometry");
    cp.fmt.bytes_per_sample = c.bits == 8 ? 1 : 2;            // DecodeStream.cpp:268-271
    cp.fmt.luma_depth = c.bits; cp.fmt.chroma_depth = c.bits;   // :265-266: the luma depth serves both
    cp.mode = c.ld ? VC2_LD : VC2_HQ_VBR;
    cp.qindex = 0; cp.picture_bytes = c.compressedBytes; cp.max_pictures = B_;
    geom_ = cp.geom;
    if (output_ == DECODED) {
      const int G = std::min(G_, std::max(1, vc2_device_count()));
      G_ = G;
      // one thread per GPU: a CUDA context, the codec's device buffers and the pinned pictures of its batches (allocated next
      // to the GPU, vc2_bind_thread_to_device) take most of a second per device
      codecs_.resize(G);
      for (int set = 0; set < 2; ++set) { recon_[set].clear(); recon_[set].resize((size_t)G * B_); }
      std::vector<std::thread> th;
      std::vector<std::string> err(G);
      for (int g = 0; g < G; ++g)
        th.emplace_back([&, g]() {
          try {
            if (G > 1) vc2_bind_thread_to_device(g);
            codecs_[g].reset(new Codec(g, cp));
            for (int set = 0; set < 2; ++set)
              for (int i = 0; i < B_; ++i) recon_[set][(size_t)g * B_ + i].resize(codecs_[g]->pictureBytes());
          } catch (const std::exception& e) { err[g] = e.what(); }
        });
      for (auto& t : th) t.join();
      for (int g = 0; g < G; ++g)
        if (!err[g].empty()) throw std::invalid_argument(err[g]);
    }
  }

  // -o Quantised / Transform / Indices: the Library-surface calls (DecodeStream.cpp:512-575)
  void taps() {
    const int depth = cfg_.pre.depth;
    const PictureFormat f(cfg_.height, cfg_.width, cfg_.cf);
    const PictureFormat tf(paddedSize(f.lumaHeight(), depth), paddedSize(f.lumaWidth(), depth), cfg_.cf);
    const Array1D qm = quantMatrix(cfg_.pre.wavelet_kernel, depth);
    for (size_t i = 0; i < pending_.size(); ++i, ++frames_) {
      const PictureUnit& u = pending_[i];
      Slices s = cfg_.ld ? readSlicesLD(u.data, u.len, tf, cfg_.pre.wavelet_kernel, depth, cfg_.pre.slices_y, cfg_.pre.slices_x,
                                        slice_bytes(cfg_.pre.slices_y, cfg_.pre.slices_x, cfg_.compressedBytes, 1))
                         : readSlicesHQ(u.data, u.len, tf, cfg_.pre.wavelet_kernel, depth, cfg_.pre.slices_y, cfg_.pre.slices_x,
                                        cfg_.pre.slice_prefix, cfg_.pre.slice_size_scalar);
      if (output_ == INDICES) {
        std::string b(s.qIndices.num_elements(), '\0');
        for (size_t j = 0; j < b.size(); ++j) b[j] = (char)s.qIndices.data()[j];
        out_.write(b.data(), (std::streamsize)b.size());
        continue;
      }
      Picture p = s.yuvCoeffs;
      if (output_ == TRANSFORM)
        p = cfg_.ld ? inverse_quantise_transform(s.yuvCoeffs, s.qIndices, qm) : inverse_quantise_transform_np(s.yuvCoeffs, s.qIndices, qm);
      write_be32_plane(out_, p.y().data(), p.y().num_elements());
      write_be32_plane(out_, p.c1().data(), p.c1().num_elements());
      write_be32_plane(out_, p.c2().data(), p.c2().num_elements());
    }
  }

  std::ostream& out_;
  vc2cli::PositionalWriter* pw_;
  Output output_;
  int G_, B_;
  bool verbose_;
  long frames_;
  Config cfg_;
  vc2_geom geom_;
  std::vector<std::unique_ptr<Codec>> codecs_;
  std::vector<vc2cli::HostBuf> recon_[2];   // two sets of decoded pictures: one is written out while the other is decoded into
  int cur_ = 0;
  double t_decode_ = 0, t_write_ = 0;
  std::thread writer_;
  std::string writeError_;
  std::vector<PictureUnit> pending_;
  std::deque<std::vector<uint8_t>> owned_;
  std::vector<uint8_t> field_, frame_;   // interlaced output: the first field waiting for its partner, the woven frame
};

// the fragments of one picture (FragmentedPictureData, DecodeStream.cpp:203): slice data keyed by first slice index
struct FragmentedPicture {
  Config cfg;
  int needed = 0, have = 0;
  std::map<int, std::pair<const uint8_t*, size_t>> parts;
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    clog << "DecodeStream (B200 hot path): decodes a VC-2 HQ / LD stream to planar raw video\n\nFor more details and useage use -h or --help" << endl;
    return EXIT_SUCCESS;
  }
  try {
    vc2cli::CmdLine c;
    c.add("v", "verbose", true);
    c.add("o", "output", false);
    c.add("G", "gpus", false);
    c.add("B", "batch", false);
    std::string inName, outName;
    Output output = DECODED;
    bool verbose = false;
    int gpus = 1, batch = 8;
    try {
      c.parse(argc, argv);
      if (c.positional().size() != 2) throw std::invalid_argument("Command line error: Required argument missing: inFile / outFile");
      inName = c.positional()[0]; outName = c.positional()[1];
      verbose = c.isSet("v");
      const std::string o = c.str("o", "Decoded");
      if (o == "Transform") output = TRANSFORM; else if (o == "Quantised") output = QUANTISED; else if (o == "Indices") output = INDICES;
      else if (o == "Decoded") output = DECODED;
      else throw std::invalid_argument("Command line error: Couldn't read argument value from string '" + o + "' for arg -o");
      gpus = std::max(1, c.integer("G", getenv("VC2_GPUS") ? atoi(getenv("VC2_GPUS")) : 1));
      batch = std::max(1, c.integer("B", getenv("VC2_BATCH") ? atoi(getenv("VC2_BATCH")) : 8));
    } catch (const std::exception& e) {
      std::cerr << "Error: " << e.what() << endl;
      return EXIT_FAILURE;
    }
    const bool timing = getenv("VC2_CLI_TIMING") != nullptr;
    const double t_start = now_s();
    // The whole stream in memory.  A regular file is mapped (no copy; the mapping is laid over an anonymous one that is 64
    // bytes longer, so the word reads of the parsers behind the last payload stay inside mapped memory); standard input and
    // pipes are read in large pieces.
    std::vector<uint8_t> streamBuf;
    const uint8_t* streamData = nullptr;
    size_t streamLen = 0;
    {
      int fd = 0;
      bool mapped = false;
      if (inName != "-") {
        fd = ::open(inName.c_str(), O_RDONLY);
        if (fd < 0) { perror(("Failed to open input file \"" + inName + "\"").c_str()); return EXIT_FAILURE; }
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
          const size_t n = (size_t)st.st_size;
          void* base = mmap(nullptr, n + 4096 + 64, PROT_READ, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
          if (base != MAP_FAILED && mmap(base, n, PROT_READ, MAP_PRIVATE | MAP_FIXED, fd, 0) == base) {
            streamData = static_cast<const uint8_t*>(base);
            streamLen = n;
            mapped = true;
          } else if (base != MAP_FAILED) {
            munmap(base, n + 4096 + 64);
          }
        }
      }
      if (!mapped) {
        const size_t piece = 16u << 20;
        for (size_t have = 0;;) {
          if (streamBuf.size() < have + piece) streamBuf.resize(have + piece);
          const ssize_t r = ::read(fd, streamBuf.data() + have, piece);
          if (r <= 0) { streamBuf.resize(have + 64, 0); streamLen = have; break; }   // + slack behind the last payload
          have += (size_t)r;
        }
        streamData = streamBuf.data();
      }
      if (fd > 0) ::close(fd);
    }
    const double t_loaded = now_s();
    std::ofstream outF;
    std::ostream* out = &std::cout;
    vc2cli::PositionalWriter pw;
    if (output == DECODED) {
      if (!pw.open(outName.c_str())) { perror(("Failed to open output file \"" + outName + "\"").c_str()); return EXIT_FAILURE; }
    } else if (outName != "-") {
      outF.open(outName.c_str(), std::ios::out | std::ios::binary);
      if (!outF) { perror(("Failed to open output file \"" + outName + "\"").c_str()); return EXIT_FAILURE; }
      out = &outF;
    }

    Decoder dec(*out, &pw, output, gpus, batch, verbose);
    StreamReader rd(streamData, streamLen);
    rd.synchronise();   // DecodeStream.cpp:177-178
    bool haveSeq = false;
    SequenceHeader seq;
    std::map<unsigned long, FragmentedPicture> fragments;
    while (!rd.atEnd()) {
      const DataUnit du = rd.readDataUnit();
      if (verbose) clog << endl << "Have read data unit of type: " << (int)du.type << endl;
      const size_t unitEnd = du.next_parse_offset ? du.offset + du.next_parse_offset : 0;
      switch (du.type) {
        case SEQUENCE_HEADER:
          seq = rd.readSequenceHeader();
          haveSeq = true;
          if (verbose) {
            clog << "height        = " << seq.height << endl << "width         = " << seq.width << endl;
            clog << "interlaced    = " << std::boolalpha << seq.interlace << endl;
          }
          break;
        case END_OF_SEQUENCE:
          if (verbose) clog << "End of Sequence after " << dec.frames() << " frames" << endl;
          break;
        case AUXILIARY_DATA:
        case PADDING_DATA:
          if ((long)du.next_parse_offset - 13 < 0) throw std::logic_error("Auxilliary data length is less than zero.");
          rd.seek(unitEnd);
          break;
        case HQ_PICTURE:
        case LD_PICTURE: {
          const bool ld = du.type == LD_PICTURE;
          unsigned long picnum = 0;
          const PicturePreamble pre = rd.readPictureHeader(ld, picnum);
          if (verbose) clog << "Picture number      : " << picnum << endl;
          if (!haveSeq) { clog << "Cannot decode frame, no previous sequence header!" << endl; if (unitEnd) rd.seek(unitEnd); break; }
          Config cfg;
          cfg.ld = ld; cfg.height = seq.interlace ? seq.height / 2 : seq.height; cfg.width = seq.width; cfg.bits = seq.bitdepth;
          cfg.cf = seq.chromaFormat; cfg.pre = pre; cfg.interlace = seq.interlace; cfg.tff = seq.topFieldFirst;
          PictureUnit u;
          u.data = streamData + rd.pos();
          if (ld) {
            // DecodeStream.cpp:312: bytes of the whole picture from the slice-bytes ratio
            cfg.compressedBytes = ld_picture_bytes(pre);
            u.len = (size_t)cfg.compressedBytes;
            if (rd.pos() + u.len > streamLen) throw std::logic_error("Stream Error: LD picture runs past the end of the stream");
          } else {
            u.len = unitEnd ? unitEnd - rd.pos()
                            : hq_payload_length(u.data, streamLen - rd.pos(), pre.slices_x * pre.slices_y, pre.slice_prefix, pre.slice_size_scalar);
            if (rd.pos() + u.len > streamLen) throw std::logic_error("Stream Error: HQ picture runs past the end of the stream");
          }
          dec.add(cfg, u);
          rd.seek(rd.pos() + u.len);
          break;
        }
        case HQ_FRAGMENT:
        case LD_FRAGMENT: {   // DecodeStream.cpp:614-977
          const bool ld = du.type == LD_FRAGMENT;
          const FragmentHeader fh = rd.readFragmentHeader();
          if (fh.n_slices == 0) {
            const PicturePreamble pre = rd.readTransformParameters(ld);
            if (verbose) clog << "Picture number      : " << fh.picture_number << endl;
            if (!haveSeq) { clog << "Cannot decode frame, no previous sequence header!" << endl; if (unitEnd) rd.seek(unitEnd); break; }
            FragmentedPicture fp;
            fp.cfg.ld = ld; fp.cfg.height = seq.interlace ? seq.height / 2 : seq.height; fp.cfg.width = seq.width; fp.cfg.bits = seq.bitdepth;
            fp.cfg.cf = seq.chromaFormat; fp.cfg.pre = pre; fp.cfg.interlace = seq.interlace; fp.cfg.tff = seq.topFieldFirst;
            if (ld) fp.cfg.compressedBytes = ld_picture_bytes(pre);
            fp.needed = pre.slices_x * pre.slices_y;
            fragments[fh.picture_number] = fp;
            if (unitEnd) rd.seek(unitEnd);
            break;
          }
          if (rd.pos() + (size_t)fh.fragment_length > streamLen) throw std::logic_error("Stream Error: fragment runs past the end of the stream");
          auto it = fragments.find(fh.picture_number);
          if (it == fragments.end()) {
            clog << "Cannot decode slices as no picture header yet read for picture number " << fh.picture_number << endl;
          } else {
            FragmentedPicture& fp = it->second;
            if (verbose) clog << "Picture " << fh.picture_number << ": Reading " << fh.n_slices << " slices, starting from (" << fh.slice_offset_x
                              << ", " << fh.slice_offset_y << ")" << endl;
            fp.parts[fh.slice_offset_y * fp.cfg.pre.slices_x + fh.slice_offset_x] = std::make_pair(streamData + rd.pos(), (size_t)fh.fragment_length);
            fp.have += fh.n_slices;
            if (fp.have >= fp.needed) {
              std::vector<uint8_t> payload;
              for (auto& part : fp.parts) payload.insert(payload.end(), part.second.first, part.second.first + part.second.second);
              const size_t len = payload.size();
              if (ld && len < (size_t)fp.cfg.compressedBytes) throw std::logic_error("Stream Error: LD picture fragments are incomplete");
              payload.resize(len + 64, 0);   // slack for the parser's word reads
              PictureUnit u;
              const Config cfg = fp.cfg;
              fragments.erase(it);
              u.len = ld ? (size_t)cfg.compressedBytes : len;
              u.data = dec.keep(std::move(payload));
              dec.add(cfg, u);
            }
          }
          rd.seek(rd.pos() + (size_t)fh.fragment_length);
          break;
        }
        default:
          throw std::logic_error("Stream Error: Unknown data unit type.");
      }
    }
    dec.flush();
    dec.finishWrite();
    out->flush();
    if (timing)
      std::cerr << "timing: stream read " << t_loaded - t_start << " s, parse + decode + write " << now_s() - t_loaded << " s for " << dec.frames()
                << " frames (busy: decode calls " << dec.decodeSeconds() << ", writer " << dec.writeSeconds() << ")" << std::endl;
    clog << "End of data stream reached successfully, exiting." << endl;
    if (output == DECODED) {
      // everything is written: leave without unpinning the host buffers and tearing the CUDA contexts down one by one
      if (!pw.ok()) { std::cerr << "Failed to write output file" << endl; return EXIT_FAILURE; }
      pw.close();
      std::cout.flush(); std::clog.flush(); std::cerr.flush();
      fflush(nullptr);
      _exit(EXIT_SUCCESS);
    }
  } catch (const std::exception& ex) {   // DecodeStream.cpp:985-988
    std::cout << "Error: " << ex.what() << endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
