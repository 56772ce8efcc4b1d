// EncodeStream - drop-in for the reference command line (src/EncodeStream/EncodeStream.cpp, EncodeParams.cpp)
// on the B200 hot path: same flags, same stream bytes.  Pictures are encoded in batches by the fused CUDA
// codec (vc2_codec_encode_host); with --gpus N consecutive batches go to different GPUs (the codec is
// intra-only) and this thread reassembles the data units in picture order.
//
// Interlaced input (-i, Frame.cpp:40-110) is coded as two field pictures per frame, each a batch slot of its own;
// fragmented pictures (-F, DataUnit.cpp:267-342) are a host-side re-chunking of the packed slices.
// -o PSNR (EncodeStream.cpp:676-767) reports, per frame, the mean / standard deviation of the slice quantiser indices
// and the PSNR of the local decode, with the reference's float arithmetic.
// LD mode (-m LD; quantIndicesLD EncodeStream.cpp:139-245, LD slice and data unit syntax) runs on the same codec.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
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

enum Output { TRANSFORM, QUANTISED, INDICES, PACKAGED, STREAM, DECODED, PSNR };
enum Mode { HQ_CBR, HQ_ConstQ, LD };

struct Params {
  std::string inFile, outFile;
  bool verbose = false;
  int height = 0, width = 0, bytes = 2, lumaDepth = 0, chromaDepth = 0;
  int cbytes = 2;   // bytes per sample the codec is fed with: -n 3 / 4 words are narrowed to their two leading bytes by the reader
  ColourFormat cf = CF_UNSET;
  bool interlaced = false, topFieldFirst = true;
  WaveletKernel kernel = NullKernel;
  int depth = 0, ySize = 0, xSize = 0;
  Output output = STREAM;
  Mode mode = HQ_ConstQ;
  int frameRate = 3, scalar = 1, prefix = 0, fragment = 0, compressedBytes = 0, qIndex = 0;
  int gpus = 1, batch = 8;
};

// EncodeParams.cpp:52-250: the same flags, checks and messages
Params parse(int argc, char** argv) {
  vc2cli::CmdLine c;
  const char* value_flags[][2] = {{"m", "mode"}, {"o", "output"}, {"a", "hSlice"}, {"u", "vSlice"}, {"d", "waveletDepth"}, {"k", "kernel"},
                                  {"c", "chromaDepth"}, {"l", "lumaDepth"}, {"z", "bitDepth"}, {"n", "bytes"}, {"f", "format"},
                                  {"x", "width"}, {"y", "height"}, {"r", "framerate"}, {"S", "scalar"}, {"P", "prefix"},
                                  {"F", "fragmentLength"}, {"s", "compressedBytes"}, {"q", "quantIndex"}, {"G", "gpus"}, {"B", "batch"}};
  for (auto& f : value_flags) c.add(f[0], f[1], false);
  const char* switches[][2] = {{"v", "verbose"}, {"b", "bottomFieldFirst"}, {"t", "topFieldFirst"}, {"i", "interlace"}, {"p", "progressive"}};
  for (auto& f : switches) c.add(f[0], f[1], true);
  c.parse(argc, argv);
  if (c.positional().size() != 2) throw std::invalid_argument("Command line error: Required argument missing: inFile / outFile");
  for (const char* req : {"m", "a", "u", "d", "k", "f", "x", "y"})
    if (!c.isSet(req)) throw std::invalid_argument(std::string("Command line error: Required argument missing: ") + req);
  Params p;
  p.inFile = c.positional()[0]; p.outFile = c.positional()[1];
  p.verbose = c.isSet("v");
  p.height = c.integer("y", 0); p.width = c.integer("x", 0);
  const std::string f = c.str("f");
  if (f == "4:4:4") p.cf = CF444; else if (f == "4:2:2") p.cf = CF422; else if (f == "4:2:0") p.cf = CF420;
  else throw std::invalid_argument("invalid colour format");
  p.bytes = c.integer("n", 2);
  int bitDepth = c.integer("z", 0);
  p.lumaDepth = c.integer("l", 0); p.chromaDepth = c.integer("c", 0);
  p.interlaced = c.isSet("i");
  p.topFieldFirst = !c.isSet("b");
  { std::istringstream ss(c.str("k")); ss >> p.kernel; }
  p.depth = c.integer("d", 0); p.ySize = c.integer("u", 0); p.xSize = c.integer("a", 0);
  const std::string o = c.str("o", "Stream");
  if (o == "Transform") p.output = TRANSFORM; else if (o == "Quantised") p.output = QUANTISED; else if (o == "Indices") p.output = INDICES;
  else if (o == "Packaged") p.output = PACKAGED; else if (o == "Stream") p.output = STREAM; else if (o == "Decoded") p.output = DECODED;
  else if (o == "PSNR") p.output = PSNR;
  else throw std::invalid_argument("Command line error: Couldn't read argument value from string '" + o + "' for arg -o");
  const std::string m = c.str("m");
  if (m == "HQ_ConstQ") p.mode = HQ_ConstQ; else if (m == "HQ_CBR") p.mode = HQ_CBR; else if (m == "LD") p.mode = LD;
  else throw std::invalid_argument("Command line error: Couldn't read argument value from string '" + m + "' for arg -m");
  p.frameRate = c.integer("r", 3);
  p.scalar = c.integer("S", 1); p.prefix = c.integer("P", 0); p.fragment = c.integer("F", 0);
  p.compressedBytes = c.integer("s", 0); p.qIndex = c.integer("q", 0);
  p.gpus = c.integer("G", getenv("VC2_GPUS") ? atoi(getenv("VC2_GPUS")) : 1);
  // default batch: 4 pictures per GPU and round for the pipelined outputs (three rounds of pinned buffers are in flight, and
  // pinning memory is the largest fixed cost of a run), 8 for the others
  const bool streamOut = c.str("o", "Stream") == "Stream" || c.str("o", "Stream") == "Packaged";
  p.batch = c.integer("B", getenv("VC2_BATCH") ? atoi(getenv("VC2_BATCH")) : (streamOut ? 4 : 8));

  if (c.isSet("z") && (c.isSet("l") || c.isSet("c")))
    throw std::invalid_argument("bitDepth is incompatible with luma depth (and/or chroma depth): use one or the other");
  if (c.isSet("p") && c.isSet("i")) throw std::invalid_argument("image can't be both interlaced and progressive: specify one or the other");
  if (c.isSet("p") && (c.isSet("t") || c.isSet("b"))) throw std::invalid_argument("field parity is incompatible with progressive image");
  if (c.isSet("t") && c.isSet("b"))
    throw std::invalid_argument("image can't be both top field first and bottom field first: specify one or the other");
  if (!c.isSet("z")) bitDepth = 8 * p.bytes;
  if (!c.isSet("l")) p.lumaDepth = bitDepth;
  if (!c.isSet("c")) p.chromaDepth = p.lumaDepth;
  if (p.height < 1) throw std::invalid_argument("picture height must be > 0");
  if (p.width < 1) throw std::invalid_argument("picture width must be > 0");
  if (p.bytes < 1 || p.bytes > 4) throw std::invalid_argument("bytes must be in range 1 to 4");
  if (c.isSet("z")) {
    if (bitDepth < 1 || bitDepth > 8 * p.bytes) throw std::invalid_argument("bit depth must be in range 1 to 8*(bytes per sample)");
  } else {
    if (p.lumaDepth < 1 || p.lumaDepth > 8 * p.bytes) throw std::invalid_argument("luma bit depth must be in range 1 to 8*(bytes per sample)");
    if (p.chromaDepth < 1 || p.chromaDepth > 8 * p.bytes)
      throw std::invalid_argument("chroma bit depth must be in range 1 to 8*(bytes per sample)");
  }
  if (p.kernel == NullKernel) throw std::invalid_argument("invalid wavelet kernel");
  if (p.depth < 1) throw std::invalid_argument("wavelet depth must be 1 or more");
  const bool hq = p.mode == HQ_CBR || p.mode == HQ_ConstQ;
  if (!hq && c.isSet("S")) throw std::invalid_argument("Slice Scalar is only used in HQ_CBR and HQ_ConstQ modes");
  if (!hq && c.isSet("P")) throw std::invalid_argument("Slice Prefix is only used in HQ_CBR and HQ_ConstQ modes");
  if (p.mode == HQ_ConstQ && c.isSet("F")) throw std::invalid_argument("Fragment length is only used in HQ_CBR and LD modes");
  if (p.mode == HQ_ConstQ && c.isSet("s")) throw std::invalid_argument("Compressed bytes is only used in HQ_CBR and LD modes");
  if (p.mode != HQ_ConstQ && c.isSet("q")) throw std::invalid_argument("Quantisation index is only used in HQ_ConstQ mode");
  if (p.mode != HQ_ConstQ && !c.isSet("s")) throw std::invalid_argument("Compressed bytes must be set in HQ_CBR and LD modes");
  if (p.mode == HQ_ConstQ && !c.isSet("q")) throw std::invalid_argument("Quantisation index must be set in HQ_ConstQ mode");
  if (hq && p.scalar < 1) throw std::invalid_argument("slice scalar must be >=1");
  if (hq && p.prefix < 0) throw std::invalid_argument("slice prefix must be >=0");
  if (p.mode != HQ_ConstQ && p.compressedBytes < 1) throw std::invalid_argument("number of compressed bytes must be >0");
  if (p.mode == HQ_ConstQ && (p.qIndex < 0 || p.qIndex > 119)) throw std::invalid_argument("quantisation index must be in the range 0 to 119");
  if (p.frameRate < 0 || p.frameRate > 16) throw std::invalid_argument("Invalid Frame Rate: ");
  // scope of this build
  // -n 3 / 4 (Arrays.cpp:333-379 reads 1..4 byte words): the samples are MSB justified, so for depths up to 16 bits the two
  // leading bytes of a word carry the whole sample and the codec reads those
  if (p.bytes > 2 && std::max(p.lumaDepth, p.chromaDepth) > 16) throw std::invalid_argument("this build codes at most 16 bits per sample");
  p.cbytes = std::min(p.bytes, 2);
  if (p.gpus < 1) p.gpus = 1;
  if (p.batch < 1) p.batch = 1;
  return p;
}

void write_be32_plane(std::ostream& out, const int32_t* v, size_t n) {   // pictureio::wordWidth(4) << signed_binary
  std::string buf(n * 4, '\0');
  for (size_t i = 0; i < n; ++i) {
    const uint32_t w = (uint32_t)v[i];
    buf[4 * i] = (char)(w >> 24); buf[4 * i + 1] = (char)(w >> 16); buf[4 * i + 2] = (char)(w >> 8); buf[4 * i + 3] = (char)w;
  }
  out.write(buf.data(), (std::streamsize)buf.size());
}

// rows of one parity of a planar frame -> a field picture and back (Frame::topField / bottomField, Frame.cpp:40-96)
void split_fields(const uint8_t* frame, uint8_t* first, uint8_t* second, const PictureFormat& ff, int bytes, bool tff) {
  const int h[3] = {ff.lumaHeight(), ff.chromaHeight(), ff.chromaHeight()};
  const size_t w[3] = {(size_t)ff.lumaWidth() * bytes, (size_t)ff.chromaWidth() * bytes, (size_t)ff.chromaWidth() * bytes};
  uint8_t* top = tff ? first : second;
  uint8_t* bot = tff ? second : first;
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < h[c]; ++y) {   // h = field height; frame rows 2y (top field) and 2y + 1 (bottom field)
      memcpy(top, frame, w[c]); top += w[c]; frame += w[c];
      memcpy(bot, frame, w[c]); bot += w[c]; frame += w[c];
    }
}
void merge_fields(uint8_t* frame, const uint8_t* first, const uint8_t* second, const PictureFormat& ff, int bytes, bool tff) {
  const int h[3] = {ff.lumaHeight(), ff.chromaHeight(), ff.chromaHeight()};
  const size_t w[3] = {(size_t)ff.lumaWidth() * bytes, (size_t)ff.chromaWidth() * bytes, (size_t)ff.chromaWidth() * bytes};
  const uint8_t* top = tff ? first : second;
  const uint8_t* bot = tff ? second : first;
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < h[c]; ++y) {
      memcpy(frame, top, w[c]); top += w[c]; frame += w[c];
      memcpy(frame, bot, w[c]); bot += w[c]; frame += w[c];
    }
}

// VC2_CLI_TIMING=1: where the wall-clock time of a run goes (standard error; measurement only)
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Worker {
  std::unique_ptr<Codec> codec;
  std::string error;
};

// One round of the pipeline: up to G batches of B pictures.  A round is filled by the reader, encoded by one host thread
// per GPU and written out in picture order; with three rounds in flight the three stages overlap (the reference reads,
// codes and writes one frame at a time, EncodeStream.cpp:452-770).
struct Round {
  std::vector<std::vector<vc2cli::HostBuf>> frames;    // [G][B] raw planar pictures
  std::vector<std::vector<vc2cli::HostBuf>> payload;   // [G][B] coded slices
  std::vector<std::vector<size_t>> len;                // [G][B]
  std::vector<int> count;                              // [G] pictures in each batch
  bool last = false;                                   // the input ended in (or before) this round
  bool noFrame0 = false;                               // not even the first frame could be read
};

// input frames: a regular file is read with pread() by several threads, anything else (a pipe, standard input) in order
class FrameSource {
 public:
  FrameSource(const std::string& name, size_t frameBytes) : fd_(0), frameBytes_(frameBytes), total_(-1), next_(0) {
    if (name != "-") {
      fd_ = ::open(name.c_str(), O_RDONLY);
      if (fd_ < 0) return;
      struct stat st;
      if (fstat(fd_, &st) == 0 && S_ISREG(st.st_mode)) total_ = (long long)(st.st_size / (off_t)frameBytes);
    }
  }
  ~FrameSource() { if (fd_ > 0) ::close(fd_); }
  bool ok() const { return fd_ >= 0; }
  bool seekable() const { return total_ >= 0; }
  // frames still to come, when known
  long long remaining() const { return total_ - next_; }
  // claim the next n frame numbers (seekable input)
  long long claim(int n) { const long long f = next_; next_ += n; return f; }
  bool readAt(long long frame, uint8_t* dst) const {
    size_t got = 0;
    while (got < frameBytes_) {
      const ssize_t r = ::pread(fd_, dst + got, frameBytes_ - got, (off_t)(frame * (long long)frameBytes_ + (long long)got));
      if (r <= 0) return false;
      got += (size_t)r;
    }
    return true;
  }
  bool readNext(uint8_t* dst) {
    size_t got = 0;
    while (got < frameBytes_) {
      const ssize_t r = ::read(fd_, dst + got, frameBytes_ - got);
      if (r <= 0) return false;
      got += (size_t)r;
    }
    return true;
  }
 private:
  int fd_;
  size_t frameBytes_;
  long long total_, next_;
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    clog << "EncodeStream (B200 hot path): encodes planar raw video into a VC-2 HQ stream\n\nFor more details and useage use -h or --help" << endl;
    return EXIT_SUCCESS;
  }
  try {
    Params p;
    try { p = parse(argc, argv); }
    catch (const std::exception& e) {   // EncodeStream.cpp:250-256
      std::cerr << "Error: " << e.what() << endl;
      return EXIT_FAILURE;
    }
    if (p.inFile != "-") {
      const int probe = ::open(p.inFile.c_str(), O_RDONLY);
      if (probe < 0) { perror(("Failed to open input file \"" + p.inFile + "\"").c_str()); return EXIT_FAILURE; }
      ::close(probe);
    }
    // -o Stream / Packaged (the pipelined outputs) go through a positional writer: large pieces, several threads; the other
    // outputs through an ostream
    const bool pipelined = p.output == STREAM || p.output == PACKAGED;
    std::ofstream outF;
    std::ostream* out = &std::cout;
    vc2cli::PositionalWriter pw;
    if (pipelined) {
      if (!pw.open(p.outFile.c_str())) { perror(("Failed to open output file \"" + p.outFile + "\"").c_str()); return EXIT_FAILURE; }
    } else if (p.outFile != "-") {
      outF.open(p.outFile.c_str(), std::ios::out | std::ios::binary);
      if (!outF) { perror(("Failed to open output file \"" + p.outFile + "\"").c_str()); return EXIT_FAILURE; }
      out = &outF;
    }
    const PictureFormat frameFormat(p.height, p.width, p.cf);
    // interlaced: every picture is a field of half the height (EncodeStream.cpp:368-377)
    const PictureFormat format(p.interlaced ? p.height / 2 : p.height, p.width, p.cf);
    const int pictureBytes = p.interlaced ? p.compressedBytes / 2 : p.compressedBytes;
    const int framePics = p.interlaced ? 2 : 1;
    const int ySlices = sliceSizeIsValid(p.depth, format.lumaHeight(), format.chromaHeight(), p.ySize);
    const int xSlices = sliceSizeIsValid(p.depth, format.lumaWidth(), format.chromaWidth(), p.xSize);
    if (ySlices == 0 || xSlices == 0) {   // EncodeStream.cpp:379-405
      if (waveletTransformIsPossible(p.depth, format.lumaWidth(), format.chromaWidth()) &&
          waveletTransformIsPossible(p.depth, format.lumaHeight(), format.chromaHeight())) {
        clog << "Consider setting --hSlice (-a) to " << suggestSliceSize(p.depth, format.lumaWidth(), format.chromaWidth(), p.xSize)
             << " and --vSlice (-u) to " << suggestSliceSize(p.depth, format.lumaHeight(), format.chromaHeight(), p.ySize) << "." << endl;
      } else {
        const int d = suggestWaveletDepth(format.lumaWidth(), format.lumaHeight(), format.chromaWidth(), format.chromaHeight(), p.depth);
        clog << "It is not possible to encode this input with a wavelet depth of " << p.depth << "." << endl;
        clog << "Consider setting --waveletDepth (-d) to " << d << " and --hSlice (-a) to "
             << suggestSliceSize(d, format.lumaWidth(), format.chromaWidth(), p.xSize) << " and --vSlice (-u) to "
             << suggestSliceSize(d, format.lumaHeight(), format.chromaHeight(), p.ySize) << "." << endl;
      }
      throw std::logic_error("The given waveletDepth, hSlice, and vSlice parameters cannot encode this input. See above for suggested parameters.");
    }
    const Array1D qMatrix = quantMatrix(p.kernel, p.depth);
    if (p.verbose) {
      clog << "Vertical slices per picture          = " << ySlices << endl;
      clog << "Horizontal slices per picture        = " << xSlices << endl;
      clog << "Quantisation matrix = " << qMatrix[0];
      for (size_t i = 1; i < qMatrix.size(); ++i) clog << ", " << qMatrix[i];
      clog << endl;
    }

    vc2_codec_params cp;
    const bool ld = p.mode == LD;
    if (vc2_make_geom(format.lumaHeight(), p.width, (int)p.cf, (int)p.kernel, p.depth, p.ySize, p.xSize, ld ? 0 : p.prefix, ld ? 1 : p.scalar,
                      &cp.geom) != VC2_OK)
      throw std::logic_error("The given waveletDepth, hSlice, and vSlice parameters cannot encode this input. See above for suggested parameters.");
    cp.fmt.bytes_per_sample = p.cbytes; cp.fmt.luma_depth = p.lumaDepth; cp.fmt.chroma_depth = p.chromaDepth;
    cp.mode = ld ? VC2_LD : p.mode == HQ_CBR ? VC2_HQ_CBR : VC2_HQ_VBR;
    cp.qindex = p.qIndex; cp.picture_bytes = pictureBytes;
    const bool taps = p.output == TRANSFORM || p.output == QUANTISED || p.output == INDICES;
    const int G = taps ? 1 : std::min(p.gpus, std::max(1, vc2_device_count()));
    const int B = p.interlaced ? (p.batch + 1) / 2 * 2 : p.batch;   // both fields of a frame in one batch
    cp.max_pictures = B;
    const bool timing = getenv("VC2_CLI_TIMING") != nullptr;
    const double t_start = now_s();
    double t_read = 0, t_code = 0, t_write = 0;   // busy seconds of the three stages
    std::vector<Worker> workers(G);
    {
      // one thread per GPU: creating a CUDA context and the codec's device buffers takes most of a second each
      std::vector<std::thread> th;
      std::vector<std::string> err(G);
      for (int g = 0; g < G; ++g)
        th.emplace_back([&, g]() {
          try { workers[g].codec.reset(new Codec(g, cp)); } catch (const std::exception& e) { err[g] = e.what(); }
        });
      for (auto& t : th) t.join();
      for (int g = 0; g < G; ++g)
        if (!err[g].empty()) throw std::invalid_argument(err[g]);
    }
    const double t_codecs = now_s();
    const size_t picBytes = workers[0].codec->pictureBytes();
    const size_t cap = workers[0].codec->payloadCapacity();

    StreamWriter writer;
    std::string unit;
    if (p.output == STREAM) {
      if (p.verbose) clog << endl << "Writing Sequence Header" << endl << endl;
      // fragmentedPictures raises the stream to major version 3 (DataUnit.cpp:1062-1067, 1412-1421)
      writer.startSequence(unit, SequenceHeader(ld ? PROFILE_LD : PROFILE_HQ, frameFormat.lumaHeight(), frameFormat.lumaWidth(), frameFormat.chromaFormat(),
                                                p.interlaced, (FrameRate)p.frameRate, p.topFieldFirst, p.lumaDepth, p.fragment > 0));
      pw.append({{unit.data(), unit.size()}});
    }
    PicturePreamble pre;
    pre.wavelet_kernel = p.kernel; pre.depth = p.depth; pre.slices_x = xSlices; pre.slices_y = ySlices;
    pre.slice_prefix = p.prefix; pre.slice_size_scalar = p.scalar;
    pre.slice_bytes = ld ? rationalise(pictureBytes, ySlices * xSlices) : rationalise(0, 1);   // EncodeStream.cpp:628-635
    // fragments (HQ_CBR only, EncodeParams.cpp:181): the slice sizes are known a priori (slice_bytes, Slices.cpp:28-49)
    std::vector<uint32_t> sliceOff;
    if (p.fragment > 0) {
      const Array2D sb = slice_bytes(ySlices, xSlices, pictureBytes, ld ? 1 : p.scalar);
      sliceOff.assign((size_t)ySlices * xSlices + 1, 0);
      for (int i = 0; i < ySlices * xSlices; ++i) sliceOff[i + 1] = sliceOff[i] + (uint32_t)(sb.data()[i] + (ld ? 0 : p.prefix));
    }

    // one round = up to G batches of B pictures; frames[g][i] is picture number frame0 + g*B + i.
    // -o Stream / Packaged: three rounds in flight - the reader fills one while the GPUs code the second and this thread
    // writes the third.  The other outputs read codec slots back after the encode, so they keep to one round.
    const int NR = pipelined ? 3 : 1;
    std::vector<Round> rounds(NR);
    for (Round& r : rounds) { r.frames.resize(G); r.payload.resize(G); r.len.assign(G, std::vector<size_t>(B, 0)); r.count.assign(G, 0); }
    // payload buffers start at the size of a raw picture (pinning memory costs about half a second per gigabyte, and the
    // worst case - every coefficient a 32-bit code - is 1.5 x raw for C3); a batch that does not fit is redone with full-size ones
    const size_t paycap = std::min(cap, std::max(picBytes, (size_t)1 << 20));
    {
      // the buffers of GPU g are allocated by a thread that runs next to it (first touch decides the NUMA node)
      std::vector<std::thread> th;
      for (int g = 0; g < G; ++g)
        th.emplace_back([&, g]() {
          if (G > 1) vc2_bind_thread_to_device(g);
          for (Round& r : rounds)
            for (int i = 0; i < B; ++i) { r.frames[g].emplace_back(picBytes); r.payload[g].emplace_back(paycap); }
        });
      for (auto& t : th) t.join();
    }
    std::vector<std::vector<uint8_t>> recon;
    if (p.output == DECODED || p.output == PSNR) recon.assign(B, std::vector<uint8_t>(picBytes));
    std::vector<uint8_t> frameBuf(p.interlaced ? 2 * picBytes : 0), wideOut;
    // file words of 3 / 4 bytes: a frame is read whole and narrowed to the codec's 2-byte words
    const int wide = p.bytes > 2 ? p.bytes : 0;
    const size_t fileFrameBytes = wide ? picBytes / 2 * (size_t)wide * framePics : picBytes * framePics;
    auto narrow = [wide](const uint8_t* src, uint8_t* dst, size_t samples) {
      for (size_t i = 0; i < samples; ++i) { dst[2 * i] = src[(size_t)wide * i]; dst[2 * i + 1] = src[(size_t)wide * i + 1]; }
    };
    auto widen = [wide](const uint8_t* src, uint8_t* dst, size_t samples) {   // -o Decoded: back to the file's word width
      for (size_t i = 0; i < samples; ++i) {
        dst[(size_t)wide * i] = src[2 * i]; dst[(size_t)wide * i + 1] = src[2 * i + 1];
        for (int k = 2; k < wide; ++k) dst[(size_t)wide * i + k] = 0;
      }
    };
    const double t_buffers = now_s();
    FrameSource source(p.inFile, fileFrameBytes);
    vc2cli::Channel<int> freeQ, readQ, codedQ;
    for (int r = 0; r < NR; ++r) freeQ.push(r);
    std::atomic<bool> stop(false);

    // stage 1: read
    std::thread reader([&]() {
      bool eof = false, first = true;
      std::vector<uint8_t> fieldsIn(p.interlaced ? 2 * picBytes : 0), fileIn(wide ? fileFrameBytes : 0);
      while (!eof) {
        const int ri = freeQ.pop();
        Round& r = rounds[ri];
        std::fill(r.count.begin(), r.count.end(), 0);
        r.last = false; r.noFrame0 = false;
        if (stop) { r.last = true; readQ.push(ri); return; }
        const double t0 = now_s();
        const int framesPerBatch = B / framePics;
        if (source.seekable()) {
          // frame f of this round goes to batch f / framesPerBatch; the reads are independent: several threads
          const long long want = (long long)G * framesPerBatch, have = std::min(want, std::max(0LL, source.remaining()));
          const long long f0 = source.claim((int)have);
          if (have < want) eof = true;
          const int T = (int)std::min<long long>(have, 4);
          std::vector<std::thread> th;
          std::vector<char> bad(T > 0 ? T : 1, 0);
          for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
              std::vector<uint8_t> both(p.interlaced ? 2 * picBytes : 0), file(wide ? fileFrameBytes : 0);
              for (long long f = t; f < have; f += T) {
                const int g = (int)(f / framesPerBatch), i = (int)(f % framesPerBatch) * framePics;
                uint8_t* dst = p.interlaced ? both.data() : r.frames[g][i].data();
                if (!source.readAt(f0 + f, wide ? file.data() : dst)) { bad[t] = 1; return; }
                if (wide) narrow(file.data(), dst, picBytes / 2 * framePics);
                if (p.interlaced) split_fields(both.data(), r.frames[g][i].data(), r.frames[g][i + 1].data(), format, p.cbytes, p.topFieldFirst);
              }
            });
          for (auto& t : th) t.join();
          for (int t = 0; t < T; ++t) if (bad[t]) eof = true;   // a file that shrank under us: stop after this round
          for (long long f = 0; f < have; ++f) r.count[(int)(f / framesPerBatch)] += framePics;
          if (first && have == 0) r.noFrame0 = true;
        } else {
          for (int g = 0; g < G && !eof; ++g)
            for (int i = 0; i < B; i += framePics) {
              uint8_t* dst = p.interlaced ? fieldsIn.data() : r.frames[g][i].data();
              const bool got = source.readNext(wide ? fileIn.data() : dst);
              if (got && wide) narrow(fileIn.data(), dst, picBytes / 2 * framePics);
              if (!got) {
                if (first && g == 0 && i == 0) r.noFrame0 = true;
                eof = true;
                break;
              }
              if (p.interlaced) split_fields(fieldsIn.data(), r.frames[g][i].data(), r.frames[g][i + 1].data(), format, p.cbytes, p.topFieldFirst);
              r.count[g] += framePics;
            }
        }
        first = false;
        r.last = eof;
        t_read += now_s() - t0;
        readQ.push(ri);
      }
    });

    // stage 2: encode, one host thread per GPU
    std::thread coder([&]() {
      for (;;) {
        const int ri = readQ.pop();
        Round& r = rounds[ri];
        const double t0 = now_s();
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) {
          if (!r.count[g] || r.noFrame0) continue;
          th.emplace_back([&, g]() {
            Worker& w = workers[g];
            if (G > 1) vc2_bind_thread_to_device(g);
            try {
              std::vector<const void*> pics(r.count[g]);
              std::vector<uint8_t*> pay(r.count[g]);
              for (int i = 0; i < r.count[g]; ++i) { pics[i] = r.frames[g][i].data(); pay[i] = r.payload[g][i].data(); }
              const int st = vc2_codec_encode_host(w.codec->handle(), r.count[g], pics.data(), pay.data(), r.payload[g][0].size(), r.len[g].data());
              if (st == VC2_ERR_CAPACITY && r.payload[g][0].size() < cap) {
                vc2_synchronize(w.codec->context());   // whatever of the failed call is still in flight
                for (int i = 0; i < B; ++i) r.payload[g][i].resize(cap);
                for (int i = 0; i < r.count[g]; ++i) pay[i] = r.payload[g][i].data();
                w.codec->encode(r.count[g], pics.data(), pay.data(), cap, r.len[g].data());
              } else {
                w.codec->check(st);
              }
            } catch (const std::exception& e) { w.error = e.what(); }
          });
        }
        for (auto& t : th) t.join();
        t_code += now_s() - t0;
        const bool last = r.last;
        codedQ.push(ri);
        if (last) break;
      }
    });
    // every way out of the loop below ends the two threads first: a stopped reader hands on an empty last round
    auto shutdown = [&]() {
      stop = true;
      for (int r = 0; r < NR; ++r) freeQ.push(r);
      if (reader.joinable()) reader.join();
      if (coder.joinable()) coder.join();
    };

    // stage 3 (this thread): ordered reassembly and output
    unsigned long long frame = 0, psnrFrame = 0;
    std::string failure;
    try {
    for (bool done = false; !done;) {
      const int ri = codedQ.pop();
      Round& r = rounds[ri];
      const double t0 = now_s();
      done = r.last;
      if (r.noFrame0) { failure = "\rFailed to read input frame number 0"; break; }
      std::vector<int>& count = r.count;
      std::vector<std::vector<vc2cli::HostBuf>>& frames = r.frames;
      // ordered reassembly.  Pipelined outputs: the data unit headers of the round as small strings, the slice bytes straight
      // from the pinned payload buffers, all handed to the positional writer as one ordered list of pieces
      std::vector<std::string> heads;
      std::vector<vc2cli::PositionalWriter::Piece> pieces;
      if (pipelined) {
        size_t total = 0;
        for (int g = 0; g < G; ++g) total += (size_t)count[g];
        heads.reserve(total);   // the pieces point into the strings: no reallocation
      }
      for (int g = 0; g < G; ++g) {
        Worker& w = workers[g];
        if (!count[g]) continue;
        if (!w.error.empty()) throw std::logic_error(w.error);
        for (int i = 0; i < count[g]; ++i, ++frame) {
          if (p.verbose) clog << "Encoded frame number " << frame << " (" << r.len[g][i] << " bytes)" << endl;
          if (p.output == STREAM) {
            heads.emplace_back();
            std::string& head = heads.back();
            // picture number = field + frame * fields per frame, wrapping at 2^32 (Utils.cpp:52-63)
            const unsigned long number = (unsigned long)(frame & 0xFFFFFFFFull);
            if (p.fragment > 0) {   // fragments re-chunk the slices: the whole data units are built in the string
              if (r.len[g][i] != sliceOff.back()) throw std::logic_error("fragment writer: payload length does not match the slice table");
              if (ld) writer.ldFragmentedPicture(head, number, pre, r.payload[g][i].data(), sliceOff.data(), p.fragment);
              else writer.hqFragmentedPicture(head, number, pre, r.payload[g][i].data(), sliceOff.data(), p.fragment);
              pieces.push_back({head.data(), head.size()});
            } else {
              if (ld) writer.ldPicture(head, number, pre, nullptr, r.len[g][i]);
              else writer.hqPicture(head, number, pre, nullptr, r.len[g][i]);
              pieces.push_back({head.data(), head.size()});
              pieces.push_back({r.payload[g][i].data(), r.len[g][i]});
            }
          } else if (p.output == PACKAGED) {
            pieces.push_back({r.payload[g][i].data(), r.len[g][i]});
          } else if (taps) {
            // the slots still hold this batch: read the requested intermediate back (EncodeStream -o Transform / Quantised / Indices)
            const size_t ny = (size_t)cp.geom.slices_y * cp.geom.slices_x;
            const size_t nY = (size_t)paddedSize(format.lumaHeight(), p.depth) * paddedSize(format.lumaWidth(), p.depth);
            const size_t nC = (size_t)paddedSize(format.chromaHeight(), p.depth) * paddedSize(format.chromaWidth(), p.depth);
            if (p.output == INDICES) {
              std::vector<int32_t> q(ny);
              w.codec->check(vc2_codec_read_indices(w.codec->handle(), i, q.data()));
              std::string b(ny, '\0');
              for (size_t j = 0; j < ny; ++j) b[j] = (char)q[j];
              out->write(b.data(), (std::streamsize)b.size());
            } else {
              std::vector<int32_t> Y(nY), U(nC), V(nC);
              w.codec->check(p.output == TRANSFORM ? vc2_codec_read_transform(w.codec->handle(), i, Y.data(), U.data(), V.data())
                                                   : vc2_codec_read_quantised(w.codec->handle(), i, Y.data(), U.data(), V.data()));
              write_be32_plane(*out, Y.data(), nY); write_be32_plane(*out, U.data(), nC); write_be32_plane(*out, V.data(), nC);
            }
          }
        }
        if (p.output == PSNR) {   // EncodeStream.cpp:676-767
          std::vector<const uint8_t*> pay(count[g]);
          std::vector<void*> pics(count[g]);
          for (int i = 0; i < count[g]; ++i) { pay[i] = r.payload[g][i].data(); pics[i] = recon[i].data(); }
          w.codec->decode(count[g], pay.data(), r.len[g].data(), pics.data());
          const size_t ns = (size_t)cp.geom.slices_y * cp.geom.slices_x;
          std::vector<int32_t> q(ns);
          const size_t nY = (size_t)format.lumaHeight() * format.lumaWidth(), nC = (size_t)format.chromaHeight() * format.chromaWidth();
          for (int i = 0; i < count[g]; i += framePics) {
            int stats[128] = {0};
            long long ss[3] = {0, 0, 0};
            for (int f = 0; f < framePics; ++f) {
              // the encode_host call has long finished: the slots still hold this batch's indices
              w.codec->check(vc2_codec_read_indices(w.codec->handle(), i + f, q.data()));
              for (size_t j = 0; j < ns; ++j) ++stats[q[j] & 127];
              const uint8_t* a = frames[g][i + f].data();
              const uint8_t* b = recon[i + f].data();
              size_t off = 0;
              for (int c = 0; c < 3; ++c) {
                const size_t n = c == 0 ? nY : nC;
                const int shift = 8 * p.cbytes - (c == 0 ? p.lumaDepth : p.chromaDepth);
                for (size_t j = 0; j < n; ++j, off += p.cbytes) {
                  const int va = p.cbytes == 2 ? ((a[off] << 8 | a[off + 1]) >> shift) : (a[off] >> shift);
                  const int vb = p.cbytes == 2 ? ((b[off] << 8 | b[off + 1]) >> shift) : (b[off] >> shift);
                  const int d = va - vb;
                  ss[c] += d * d;
                }
              }
            }
            const int totalSlices = framePics * (int)ns;
            int topIndex = -1;
            for (int z = 0; z < 128; ++z) if (stats[z] > 0) topIndex = z;
            float mean = 0.0, meanSquare = 0.0;
            for (int z = 0; z <= topIndex; ++z) { mean += (z * stats[z]); meanSquare += (z * z * stats[z]); }
            mean /= totalSlices;
            meanSquare /= totalSlices;
            // the reference's unqualified sqrt / log10 resolve to the C double functions; results are narrowed to float
            const float stdDev = (float)::sqrt((double)(meanSquare - (mean * mean)));
            const int yPixels = p.width * p.height, uvPixels = frameFormat.chromaWidth() * frameFormat.chromaHeight();
            const float YRMS = (float)(::sqrt((double)(float(ss[0]) / float(yPixels))) / (1 << p.lumaDepth));
            const float URMS = (float)(::sqrt((double)(float(ss[1]) / float(uvPixels))) / (1 << p.chromaDepth));
            const float VRMS = (float)(::sqrt((double)(float(ss[2]) / float(uvPixels))) / (1 << p.chromaDepth));
            const float YPSNR = (float)(-20 * ::log10((double)YRMS)), UPSNR = (float)(-20 * ::log10((double)URMS)),
                        VPSNR = (float)(-20 * ::log10((double)VRMS));
            *out << "Frame " << psnrFrame++ << endl;
            *out << std::fixed << std::setprecision(2);
            *out << mean << " " << stdDev << endl;
            *out << std::fixed << std::setprecision(4);
            *out << YPSNR << " " << UPSNR << " " << VPSNR << endl;
          }
        }
        if (p.output == DECODED) {   // local decode of the batch just written (EncodeStream.cpp:649-690)
          std::vector<const uint8_t*> pay(count[g]);
          std::vector<void*> pics(count[g]);
          for (int i = 0; i < count[g]; ++i) { pay[i] = r.payload[g][i].data(); pics[i] = recon[i].data(); }
          w.codec->decode(count[g], pay.data(), r.len[g].data(), pics.data());
          for (int i = 0; i < count[g]; i += framePics) {
            const uint8_t* pic = recon[i].data();
            size_t n = picBytes;
            if (p.interlaced) {
              merge_fields(frameBuf.data(), recon[i].data(), recon[i + 1].data(), format, p.cbytes, p.topFieldFirst);
              pic = frameBuf.data(); n = frameBuf.size();
            }
            if (wide) {
              wideOut.resize(n / 2 * (size_t)wide);
              widen(pic, wideOut.data(), n / 2);
              pic = wideOut.data(); n = wideOut.size();
            }
            out->write(reinterpret_cast<const char*>(pic), (std::streamsize)n);
          }
        }
        if (!*out) { failure = "Failed to write output file \"" + p.outFile + "\""; break; }
      }
      if (pipelined) {
        pw.append(pieces);
        if (!pw.ok()) failure = "Failed to write output file \"" + p.outFile + "\"";
      }
      if (!failure.empty()) break;
      t_write += now_s() - t0;
      if (!done) freeQ.push(ri);   // the round's buffers go back to the reader
    }
    } catch (...) { shutdown(); throw; }
    shutdown();
    if (!failure.empty()) { std::cerr << failure << endl; return EXIT_FAILURE; }
    if (p.verbose) clog << "\rEnd of input reached after " << frame << " frames" << endl;
    if (p.output == STREAM) {
      unit.clear();
      writer.endSequence(unit);
      pw.append({{unit.data(), unit.size()}});
      if (!pw.ok()) { std::cerr << "Failed to write output file \"" << p.outFile << "\"" << endl; return EXIT_FAILURE; }
    }
    pw.close();
    out->flush();
    if (timing)
      std::cerr << "timing: codecs " << t_codecs - t_start << " s, host buffers " << t_buffers - t_codecs << " s, pipeline " << now_s() - t_buffers
                << " s for " << frame << " pictures (busy: read " << t_read << ", code " << t_code << ", write " << t_write << ")" << endl;
    if (pipelined) {
      // everything is written and closed: leave without unpinning gigabytes of host buffers and tearing the CUDA contexts down
      // one by one (about a second for a job that takes a few)
      std::cout.flush(); std::clog.flush(); std::cerr.flush();
      fflush(nullptr);
      _exit(EXIT_SUCCESS);
    }
  } catch (const std::exception& ex) {   // EncodeStream.cpp:782-785: message on standard OUTPUT, failure status
    std::cout << "Error: " << ex.what() << endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
