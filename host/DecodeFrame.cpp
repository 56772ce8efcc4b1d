// DecodeFrame - drop-in for the reference command line (src/DecodeFrame/DecodeFrame.cpp, DecodeParams.cpp):
// decodes bare slice data (EncodeStream -o Packaged: no stream syntax) with every parameter given on the command
// line.  HQ pictures go through the fused batched CUDA codec; the -o Transform / Quantised / Indices taps go through
// the Library-surface calls.  Two behaviours of the reference tool are kept as they are: after the tap output of every
// frame it still writes its (never assigned, all-zero) output frame (the `continue` at DecodeFrame.cpp:263-299 only
// leaves the field loop, :313-333 then runs) - and because that frame write leaves pictureio::bitDepth set on the
// stream, every later tap word is left justified to it (Arrays.cpp:170-172, 391-397) - and LD input fails on the first picture (probed: the reference reports
// "Failed to read the first compressed frame" for streams its own LD encoder wrote).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cmdline.h"
#include "vc2/Codec.h"
#include "vc2/Quantisation.h"
#include "vc2/Slices.h"
#include "vc2/WaveletTransform.h"

using namespace vc2;
using std::clog;
using std::endl;

namespace {

enum Output { TRANSFORM, QUANTISED, INDICES, DECODED };

void write_be32_plane(std::ostream& out, const int* v, size_t n, int shift) {
  std::string buf(n * 4, '\0');
  for (size_t i = 0; i < n; ++i) {
    const uint32_t w = (uint32_t)v[i] << shift;
    buf[4 * i] = (char)(w >> 24); buf[4 * i + 1] = (char)(w >> 16); buf[4 * i + 2] = (char)(w >> 8); buf[4 * i + 3] = (char)w;
  }
  out.write(buf.data(), (std::streamsize)buf.size());
}

// one int plane -> file words: offset binary, left justified, big endian (Arrays.cpp:380-426)
void append_samples(std::string& out, const Array2D& a, int bytes, int depth) {
  const int offset = 1 << (depth - 1), shift = 8 * bytes - depth;
  for (size_t i = 0; i < a.num_elements(); ++i) {
    const uint32_t w = (uint32_t)(a.data()[i] + offset) << shift;
    for (int b = bytes - 1; b >= 0; --b) out.push_back((char)(w >> (8 * b)));
  }
}

// two field pictures (planar bytes) -> frame rows (Frame.cpp:62-110)
void weave(std::string& frame, const std::string& first, const std::string& second, const PictureFormat& ff, int bytes, bool tff) {
  const int h[3] = {ff.lumaHeight(), ff.chromaHeight(), ff.chromaHeight()};
  const size_t w[3] = {(size_t)ff.lumaWidth() * bytes, (size_t)ff.chromaWidth() * bytes, (size_t)ff.chromaWidth() * bytes};
  frame.resize(first.size() + second.size());
  const char* top = tff ? first.data() : second.data();
  const char* bot = tff ? second.data() : first.data();
  char* dst = &frame[0];
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < h[c]; ++y) {
      memcpy(dst, top, w[c]); top += w[c]; dst += w[c];
      memcpy(dst, bot, w[c]); bot += w[c]; dst += w[c];
    }
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    clog << "DecodeFrame (B200 hot path): decodes VC-2 compressed bytes without stream syntax to an uncompressed planar file\n\n"
            "For more details and useage use -h or --help" << endl;
    return EXIT_SUCCESS;
  }
  try {
    vc2cli::CmdLine c;
    const char* value_flags[][2] = {{"m", "mode"}, {"o", "output"}, {"a", "hSlice"}, {"u", "vSlice"}, {"d", "waveletDepth"}, {"k", "kernel"},
                                    {"c", "chromaDepth"}, {"l", "lumaDepth"}, {"z", "bitDepth"}, {"n", "bytes"}, {"f", "format"},
                                    {"x", "width"}, {"y", "height"}, {"S", "scalar"}, {"P", "prefix"}, {"s", "compressedBytes"}, {"B", "batch"}};
    for (auto& f : value_flags) c.add(f[0], f[1], false);
    const char* switches[][2] = {{"v", "verbose"}, {"b", "bottomFieldFirst"}, {"t", "topFieldFirst"}, {"i", "interlace"}, {"p", "progressive"}};
    for (auto& f : switches) c.add(f[0], f[1], true);
    std::string inName, outName;
    bool verbose = false, ld = false, interlaced = false, tff = true;
    int height = 0, width = 0, bytes = 2, lumaDepth = 0, chromaDepth = 0, depth = 0, ySize = 0, xSize = 0, scalar = 1, prefix = 0;
    int compressedBytes = 0, batch = 8;
    ColourFormat cf = CF_UNSET;
    WaveletKernel kernel = NullKernel;
    Output output = DECODED;
    try {   // DecodeParams.cpp:60-206: the same flags, checks and messages
      c.parse(argc, argv);
      if (c.positional().size() != 2) throw std::invalid_argument("Required argument missing: inFile / outFile");
      for (const char* req : {"a", "u", "d", "k", "f", "x", "y"})
        if (!c.isSet(req)) throw std::invalid_argument(std::string("Required argument missing: ") + req);
      inName = c.positional()[0]; outName = c.positional()[1];
      verbose = c.isSet("v");
      height = c.integer("y", 0); width = c.integer("x", 0);
      const std::string f = c.str("f");
      if (f == "4:4:4") cf = CF444; else if (f == "4:2:2") cf = CF422; else if (f == "4:2:0") cf = CF420;
      bytes = c.integer("n", 2);
      int bitDepth = c.integer("z", 0);
      lumaDepth = c.integer("l", 0); chromaDepth = c.integer("c", 0);
      interlaced = c.isSet("i"); tff = !c.isSet("b");
      { std::istringstream ss(c.str("k")); ss >> kernel; }
      depth = c.integer("d", 0); ySize = c.integer("u", 0); xSize = c.integer("a", 0);
      const std::string o = c.str("o", "Decoded");
      if (o == "Transform") output = TRANSFORM; else if (o == "Quantised") output = QUANTISED; else if (o == "Indices") output = INDICES;
      else if (o == "Decoded") output = DECODED;
      else throw std::invalid_argument("Couldn't read argument value from string '" + o + "' for arg -o");
      const std::string m = c.str("m", "HQ");
      if (m == "HQ") ld = false; else if (m == "LD") ld = true;
      else throw std::invalid_argument("Couldn't read argument value from string '" + m + "' for arg -m");
      scalar = c.integer("S", 1); prefix = c.integer("P", 0); compressedBytes = c.integer("s", 0);
      batch = std::max(1, c.integer("B", getenv("VC2_BATCH") ? atoi(getenv("VC2_BATCH")) : 8));
      if (c.isSet("z") && (c.isSet("l") || c.isSet("c")))
        throw std::invalid_argument("bitDepth is incompatible with luma depth (and/or chroma depth): use one or the other");
      if (c.isSet("p") && c.isSet("i")) throw std::invalid_argument("image can't be both interlaced and progressive: specify one or the other");
      if (c.isSet("p") && (c.isSet("t") || c.isSet("b"))) throw std::invalid_argument("field parity is incompatible with progressive image");
      if (c.isSet("t") && c.isSet("b"))
        throw std::invalid_argument("image can't be both top field first and bottom field first: specify one or the other");
      if (!c.isSet("z")) bitDepth = 8 * bytes;
      if (!c.isSet("l")) lumaDepth = bitDepth;
      if (!c.isSet("c")) chromaDepth = lumaDepth;
      if (height < 1) throw std::invalid_argument("picture height must be > 0");
      if (width < 1) throw std::invalid_argument("picture width must be > 0");
      if (cf == CF_UNSET) throw std::invalid_argument("unknown colour format");
      if (bytes < 1 || bytes > 4) throw std::invalid_argument("bytes must be in range 1 to 4");
      if (c.isSet("z")) {
        if (bitDepth < 1 || bitDepth > 8 * bytes) throw std::invalid_argument("bit depth must be in range 1 to 8*(bytes per sample)");
      } else {
        if (lumaDepth < 1 || lumaDepth > 8 * bytes) throw std::invalid_argument("luma bit depth must be in range 1 to 8*(bytes per sample)");
        if (chromaDepth < 1 || chromaDepth > 8 * bytes) throw std::invalid_argument("chroma bit depth must be in range 1 to 8*(bytes per sample)");
      }
      if (kernel == NullKernel) throw std::invalid_argument("invalid wavelet kernel");
      if (depth < 1) throw std::invalid_argument("wavelet depth must be 1 or more");
      if (ld && !c.isSet("s")) throw std::invalid_argument("In LD mode compressedBytes must be set");
      if (ld && c.isSet("P")) throw std::invalid_argument("In LD mode slicePrefix is not required");
      if (ld && c.isSet("S")) throw std::invalid_argument("In LD mode sliceScalar is not required");
      if (!ld && c.isSet("s")) throw std::invalid_argument("In HQ mode compressedBytes is not required");
      if (!ld && scalar < 1) throw std::invalid_argument("Slice Scalar must be 1 or more");
      if (!ld && prefix < 0) throw std::invalid_argument("Slice Prefix must be 0 or more");
    } catch (const std::exception& e) {   // DecodeFrame.cpp:66-69
      std::cerr << "Command line error: " << e.what() << endl;
      return EXIT_FAILURE;
    }
    std::vector<uint8_t> data;
    if (inName == "-") data.assign(std::istreambuf_iterator<char>(std::cin), std::istreambuf_iterator<char>());
    else {
      std::ifstream f(inName.c_str(), std::ios::in | std::ios::binary);
      if (!f) { perror(("Failed to open input file \"" + inName + "\"").c_str()); return EXIT_FAILURE; }
      data.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    if (ld) { std::cerr << "\rFailed to read the first compressed frame" << endl; return EXIT_FAILURE; }
    const size_t dataLen = data.size();
    data.resize(dataLen + 64, 0);   // slack behind the last payload for the parser's word reads
    std::ofstream outF;
    std::ostream* out = &std::cout;
    if (outName != "-") {
      outF.open(outName.c_str(), std::ios::out | std::ios::binary);
      if (!outF) { perror(("Failed to open output file \"" + outName + "\"").c_str()); return EXIT_FAILURE; }
      out = &outF;
    }

    // DecodeFrame.cpp:172-216
    const int yT = ySize << depth, xT = xSize << depth;
    const int pictureHeight = interlaced ? height / 2 : height;
    const int paddedH = paddedSize(pictureHeight, depth), paddedW = paddedSize(width, depth);
    if (yT < 1 || xT < 1) throw std::logic_error("Padded picture height is not divisible by slice height");
    const int ySlices = paddedH / yT, xSlices = paddedW / xT;
    if (paddedH != ySlices * yT) throw std::logic_error("Padded picture height is not divisible by slice height");
    if (paddedW != xSlices * xT) throw std::logic_error("Padded width is not divisible by slice width");
    if (verbose) {
      clog << "Vertical slices per picture          = " << ySlices << endl;
      clog << "Horizontal slices per picture        = " << xSlices << endl;
    }
    const Array1D qMatrix = quantMatrix(kernel, depth);
    const int pictureBytes = interlaced ? compressedBytes / 2 : compressedBytes;
    const PictureFormat transformFormat(paddedH, paddedW, cf), picFormat(pictureHeight, width, cf);
    const int framePics = interlaced ? 2 : 1;
    const int nslices = ySlices * xSlices;

    // split the input into pictures: LD pictures have a fixed size, HQ pictures end where their last slice ends
    std::vector<std::pair<size_t, size_t>> pictures;
    {
      size_t pos = 0;
      std::vector<uint32_t> off(nslices + 1);
      while (pos < dataLen) {
        size_t len;
        if (ld) {
          len = (size_t)pictureBytes;
          if (pos + len > dataLen) break;
        } else {
          if (vc2_hq_index_slices(data.data() + pos, dataLen - pos, nslices, prefix, scalar, off.data()) != VC2_OK) break;
          len = off[nslices];
        }
        pictures.push_back(std::make_pair(pos, len));
        pos += len;
        if (len == 0) break;
      }
    }
    if (pictures.empty()) { std::cerr << "\rFailed to read the first compressed frame" << endl; return EXIT_FAILURE; }

    const bool fused = output == DECODED && bytes <= 2;
    // the reference's zero frame behind every frame of tap output: sample 0 in offset binary, left justified
    std::string zeroFrame;
    if (output != DECODED) {
      const PictureFormat ff(height, width, cf);
      const size_t n = (size_t)ff.lumaHeight() * ff.lumaWidth() + 2 * (size_t)ff.chromaHeight() * ff.chromaWidth();
      zeroFrame.assign(n * bytes, '\0');
      for (size_t i = 0; i < n; ++i) zeroFrame[i * bytes] = (char)0x80;
    }
    std::unique_ptr<Codec> codec;
    int B = interlaced ? (batch + 1) / 2 * 2 : batch;
    if (fused) {
      vc2_codec_params cp;
      if (vc2_make_geom(pictureHeight, width, (int)cf, (int)kernel, depth, ySize, xSize, prefix, scalar, &cp.geom) != VC2_OK)
        throw std::logic_error("Padded picture height is not divisible by slice height");
      cp.fmt.bytes_per_sample = bytes; cp.fmt.luma_depth = lumaDepth; cp.fmt.chroma_depth = chromaDepth;
      cp.mode = VC2_HQ_VBR; cp.qindex = 0; cp.picture_bytes = 0; cp.max_pictures = B;
      codec.reset(new Codec(0, cp));
    }
    std::string first, frameBuf;
    long frame = 0;
    bool sticky = false;   // a frame has been written: the stream now carries pictureio::bitDepth
    const size_t whole = pictures.size() / framePics * framePics;   // a dangling first field is never written (:236-248)
    for (size_t base = 0; base < (output == DECODED ? whole : pictures.size()); base += B) {
      const int n = (int)std::min((size_t)B, (output == DECODED ? whole : pictures.size()) - base);
      std::vector<std::string> pics(n);
      if (fused) {
        std::vector<const uint8_t*> pay(n);
        std::vector<size_t> len(n);
        std::vector<void*> dst(n);
        for (int i = 0; i < n; ++i) {
          pics[i].resize(codec->pictureBytes());
          pay[i] = data.data() + pictures[base + i].first; len[i] = pictures[base + i].second; dst[i] = &pics[i][0];
        }
        codec->decode(n, pay.data(), len.data(), dst.data());
      } else {
        for (int i = 0; i < n; ++i) {
          const uint8_t* p = data.data() + pictures[base + i].first;
          const size_t len = pictures[base + i].second;
          Slices s = readSlicesHQ(p, len, transformFormat, kernel, depth, ySlices, xSlices, prefix, scalar);
          if (output == INDICES) {
            clog << "Writing quantisation indices to output file" << endl;
            std::string b(s.qIndices.num_elements(), '\0');
            // one-byte words: the sticky shift 8 - chromaDepth may be negative; x86 takes the shift count modulo 32
            const int sh = sticky ? ((8 - chromaDepth) & 31) : 0;
            for (size_t j = 0; j < b.size(); ++j) b[j] = (char)((uint32_t)s.qIndices.data()[j] << sh);
            out->write(b.data(), (std::streamsize)b.size());
            if ((base + i) % framePics == (size_t)framePics - 1) { out->write(zeroFrame.data(), (std::streamsize)zeroFrame.size()); sticky = true; }
            continue;
          }
          Picture q = s.yuvCoeffs;
          if (output != QUANTISED) q = inverse_quantise_transform_np(s.yuvCoeffs, s.qIndices, qMatrix);   // DecodeFrame.cpp:290, both modes
          if (output == QUANTISED || output == TRANSFORM) {
            clog << (output == QUANTISED ? "Writing quantised transform coefficients to output file" : "Writing transform coefficients to output file") << endl;
            write_be32_plane(*out, q.y().data(), q.y().num_elements(), sticky ? 32 - lumaDepth : 0);
            write_be32_plane(*out, q.c1().data(), q.c1().num_elements(), sticky ? 32 - chromaDepth : 0);
            write_be32_plane(*out, q.c2().data(), q.c2().num_elements(), sticky ? 32 - chromaDepth : 0);
            if ((base + i) % framePics == (size_t)framePics - 1) { out->write(zeroFrame.data(), (std::streamsize)zeroFrame.size()); sticky = true; }
            continue;
          }
          Picture pic = inverseWaveletTransform(q, kernel, depth, picFormat);
          pic = clip(pic, -(1 << (lumaDepth - 1)), (1 << (lumaDepth - 1)) - 1, -(1 << (chromaDepth - 1)), (1 << (chromaDepth - 1)) - 1);
          append_samples(pics[i], pic.y(), bytes, lumaDepth);
          append_samples(pics[i], pic.c1(), bytes, chromaDepth);
          append_samples(pics[i], pic.c2(), bytes, chromaDepth);
        }
      }
      if (output != DECODED) continue;
      for (int i = 0; i < n; i += framePics, ++frame) {
        if (verbose) clog << "Writing decoded output file" << endl;
        if (interlaced) {
          weave(frameBuf, pics[i], pics[i + 1], picFormat, bytes, tff);
          out->write(frameBuf.data(), (std::streamsize)frameBuf.size());
        } else {
          out->write(pics[i].data(), (std::streamsize)pics[i].size());
        }
      }
      if (!*out) { std::cerr << "Failed to write output file \"" << outName << "\"" << endl; return EXIT_FAILURE; }
    }
    if (verbose) clog << "\rEnd of input reached after " << frame << " frames     " << endl;
    out->flush();
  } catch (const std::exception& ex) {   // DecodeFrame.cpp:352-355
    std::cout << "Error: " << ex.what() << endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
