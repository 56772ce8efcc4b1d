// Calls every function of the C++ Library mirror (include/vc2/*.h, host/vc2_library.cpp) on one random 4:2:2 picture and
// compares each result, bit for bit, with the unmodified reference Library behind oracle/_ref/libvc2ref.so (extern "C"
// taps of oracle/ref_taps.cpp, loaded with dlopen).  Test infrastructure: needs a GPU.  Exit status 0 = all equal.
//   usage: test_library_mirror <path to libvc2ref.so>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "vc2/Quantisation.h"
#include "vc2/Slices.h"
#include "vc2/WaveletTransform.h"

using namespace vc2;

static void* g_ref = nullptr;
template <class F> F sym(const char* name) {
  void* p = dlsym(g_ref, name);
  if (!p) { std::fprintf(stderr, "missing reference tap %s\n", name); std::exit(2); }
  return reinterpret_cast<F>(p);
}
static int g_fail = 0;
static void expect(bool ok, const std::string& what) {
  std::printf("%-58s %s\n", what.c_str(), ok ? "same" : "DIFFERENT");
  if (!ok) ++g_fail;
}
static bool same(const Array2D& a, const std::vector<int>& b) {
  return a.num_elements() == b.size() && std::memcmp(a.data(), b.data(), b.size() * sizeof(int)) == 0;
}
static Array2D randomPlane(int h, int w, int amp, unsigned seed) {
  std::mt19937 rng(seed);
  std::uniform_int_distribution<int> d(-amp, amp);
  Array2D a(h, w);
  for (size_t i = 0; i < a.num_elements(); ++i) a.data()[i] = d(rng);
  return a;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s libvc2ref.so\n", argv[0]); return 2; }
  g_ref = dlopen(argv[1], RTLD_NOW);
  if (!g_ref) { std::fprintf(stderr, "%s\n", dlerror()); return 2; }
  typedef int (*dwt_f)(const int*, int, int, int, int, int*);
  typedef int (*idwt_f)(const int*, int, int, int, int, int*, int, int);
  typedef int (*quant_f)(const int*, int, int, const int*, int, int, const int*, int, int*);
  typedef int (*qm_f)(int, int, int*);
  typedef int (*cbr_f)(const int*, const int*, const int*, int, int, int, int, const int*, int, const int*, int, int, int, int*);
  typedef int (*pack_f)(const int*, const int*, const int*, int, int, int, int, int, const int*, int, int, int, int, int, const int*,
                        unsigned char*, long, long*);
  typedef int (*unpack_f)(const unsigned char*, long, int, int, int, int, int, int, int, int, int, int, const int*, int*, int*, int*, int*);
  typedef int (*sb_f)(int, int, int, int, int*);
  typedef int (*i4_f)(int, int, int, int);
  typedef int (*i2_f)(int, int);
  const dwt_f ref_dwt = sym<dwt_f>("ref_dwt_forward");
  const idwt_f ref_idwt = sym<idwt_f>("ref_dwt_inverse");
  const quant_f ref_q = sym<quant_f>("ref_quantise_np"), ref_dq = sym<quant_f>("ref_dequantise_np"), ref_dqld = sym<quant_f>("ref_dequantise_ld");
  const qm_f ref_qm = sym<qm_f>("ref_quant_matrix");
  const cbr_f ref_cbr = sym<cbr_f>("ref_cbr_qindices");
  const pack_f ref_pack = sym<pack_f>("ref_pack_slices");
  const unpack_f ref_unpack = sym<unpack_f>("ref_unpack_slices");
  const sb_f ref_sb = sym<sb_f>("ref_slice_bytes");
  const i4_f ref_valid = sym<i4_f>("ref_slice_size_is_valid");
  const i2_f ref_padded = sym<i2_f>("ref_padded_size");

  const int H = 70, W = 300, depth = 3, prefix = 1, scalar = 2;   // needs padding: 72 x 304 (chroma 72 x 152)
  const WaveletKernel kernels[] = {DD97, LeGall, DD137, Haar0, Haar1, Fidelity, Daub97};
  try {
    const PictureFormat f(H, W, CF422);
    const Picture pic(f, randomPlane(H, W, 500, 1), randomPlane(H, W / 2, 500, 2), randomPlane(H, W / 2, 500, 3));
    expect(paddedSize(H, depth) == ref_padded(H, depth) && paddedSize(W, depth) == ref_padded(W, depth), "paddedSize");
    const int ny = sliceSizeIsValid(depth, f.lumaHeight(), f.chromaHeight(), 1), nx = sliceSizeIsValid(depth, f.lumaWidth(), f.chromaWidth(), 2);
    expect(ny == ref_valid(depth, f.lumaHeight(), f.chromaHeight(), 1) && nx == ref_valid(depth, f.lumaWidth(), f.chromaWidth(), 2) && ny > 0 && nx > 0,
           "sliceSizeIsValid");
    for (WaveletKernel k : kernels) {
      const std::string kn = " kernel " + std::to_string((int)k);
      // quantMatrix
      const Array1D qm = quantMatrix(k, depth);
      std::vector<int> rqm(3 * depth + 1);
      ref_qm((int)k, depth, rqm.data());
      expect(qm == rqm, "quantMatrix" + kn);
      // waveletTransform (Picture and Array2D overloads)
      const Picture t = waveletTransform(pic, k, depth);
      const int ph = (int)t.y().shape()[0], pw = (int)t.y().shape()[1], ch = (int)t.c1().shape()[0], cw = (int)t.c1().shape()[1];
      std::vector<int> ry((size_t)ph * pw), ru((size_t)ch * cw), rv((size_t)ch * cw);
      ref_dwt(pic.y().data(), H, W, (int)k, depth, ry.data());
      ref_dwt(pic.c1().data(), H, W / 2, (int)k, depth, ru.data());
      ref_dwt(pic.c2().data(), H, W / 2, (int)k, depth, rv.data());
      expect(same(t.y(), ry) && same(t.c1(), ru) && same(t.c2(), rv), "waveletTransform(Picture)" + kn);
      expect(same(waveletTransform(pic.y(), k, depth), ry), "waveletTransform(Array2D)" + kn);
      // quantIndicesConstQ + quantise_transform_np (all overloads)
      const Array2D qc = quantIndicesConstQ(ny, nx, 9 + (int)k);
      const Picture q = quantise_transform_np(t, qc, qm);
      std::vector<int> rq((size_t)ph * pw), rqu((size_t)ch * cw), rqv((size_t)ch * cw);
      ref_q(ry.data(), ph, pw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), rq.data());
      ref_q(ru.data(), ch, cw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), rqu.data());
      ref_q(rv.data(), ch, cw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), rqv.data());
      expect(same(q.y(), rq) && same(q.c1(), rqu) && same(q.c2(), rqv), "quantise_transform_np(Picture, indices)" + kn);
      expect(same(quantise_transform_np(t, 9 + (int)k, qm).y(), rq), "quantise_transform_np(Picture, index)" + kn);
      expect(same(quantise_transform_np(t.y(), qc, qm), rq), "quantise_transform_np(Array2D)" + kn);
      // inverse_quantise_transform_np, inverse_quantise_transform (LD)
      std::vector<int> rd((size_t)ph * pw), rdld((size_t)ph * pw);
      ref_dq(rq.data(), ph, pw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), rd.data());
      ref_dqld(rq.data(), ph, pw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), rdld.data());
      const Picture dq = inverse_quantise_transform_np(q, qc, qm);
      expect(same(dq.y(), rd), "inverse_quantise_transform_np(Picture, indices)" + kn);
      expect(same(inverse_quantise_transform_np(q, 9 + (int)k, qm).y(), rd), "inverse_quantise_transform_np(Picture, index)" + kn);
      expect(same(inverse_quantise_transform_np(q.y(), qc, qm), rd), "inverse_quantise_transform_np(Array2D)" + kn);
      expect(same(inverse_quantise_transform(q, qc, qm).y(), rdld), "inverse_quantise_transform (LD)" + kn);
      // inverseWaveletTransform (both overloads)
      std::vector<int> ri((size_t)H * W);
      ref_idwt(rd.data(), ph, pw, (int)k, depth, ri.data(), H, W);
      expect(same(inverseWaveletTransform(dq, k, depth, f).y(), ri), "inverseWaveletTransform(Picture)" + kn);
      expect(same(inverseWaveletTransform(dq.y(), k, depth, H, W), ri), "inverseWaveletTransform(Array2D)" + kn);
      // slice_bytes, quantIndicesCBR
      const int total = (ph * pw + 2 * ch * cw) * 10 / 8 / 4;
      const Array2D sb = slice_bytes(ny, nx, total, scalar);
      std::vector<int> rsb((size_t)ny * nx), rci((size_t)ny * nx);
      ref_sb(ny, nx, total, scalar, rsb.data());
      expect(same(sb, rsb), "slice_bytes" + kn);
      const Array2D ci = quantIndicesCBR(t, qm, sb, scalar, depth);
      ref_cbr(ry.data(), ru.data(), rv.data(), ph, pw, ch, cw, rqm.data(), (int)rqm.size(), rsb.data(), ny, nx, scalar, rci.data());
      expect(same(ci, rci), "quantIndicesCBR" + kn);
      // writeSlicesHQVBR / writeSlicesHQCBR / readSlicesHQ
      Slices s;
      s.yuvCoeffs = q; s.waveletDepth = depth; s.qIndices = qc;
      const std::string vbr = writeSlicesHQVBR(s, k, prefix, scalar);
      std::vector<unsigned char> rbuf((size_t)8 * (ph * pw + 2 * ch * cw) + 65536);   // also holds the LD picture below
      long rlen = 0;
      ref_pack(rq.data(), rqu.data(), rqv.data(), ph, pw, ch, cw, depth, qc.data(), ny, nx, 0, prefix, scalar, nullptr, rbuf.data(), (long)rbuf.size(), &rlen);
      expect((long)vbr.size() == rlen && std::memcmp(vbr.data(), rbuf.data(), (size_t)rlen) == 0, "writeSlicesHQVBR" + kn);
      const Picture qcbr = quantise_transform_np(t, ci, qm);
      Slices sc;
      sc.yuvCoeffs = qcbr; sc.waveletDepth = depth; sc.qIndices = ci;
      const std::string cbr = writeSlicesHQCBR(sc, k, sb, prefix, scalar);
      ref_pack(qcbr.y().data(), qcbr.c1().data(), qcbr.c2().data(), ph, pw, ch, cw, depth, rci.data(), ny, nx, 1, prefix, scalar, rsb.data(), rbuf.data(),
               (long)rbuf.size(), &rlen);
      expect((long)cbr.size() == rlen && std::memcmp(cbr.data(), rbuf.data(), (size_t)rlen) == 0, "writeSlicesHQCBR" + kn);
      const PictureFormat tf(ph, pw, CF422);
      const Slices back = readSlicesHQ(reinterpret_cast<const uint8_t*>(vbr.data()), vbr.size(), tf, k, depth, ny, nx, prefix, scalar);
      std::vector<int> by((size_t)ph * pw), bu((size_t)ch * cw), bv((size_t)ch * cw), bq((size_t)ny * nx);
      ref_unpack(reinterpret_cast<const unsigned char*>(vbr.data()), (long)vbr.size(), ph, pw, ch, cw, depth, ny, nx, 0, prefix, scalar, nullptr,
                 by.data(), bu.data(), bv.data(), bq.data());
      expect(same(back.yuvCoeffs.y(), by) && same(back.yuvCoeffs.c1(), bu) && same(back.yuvCoeffs.c2(), bv) && same(back.qIndices, bq), "readSlicesHQ" + kn);
      // readSlicesLD: an LD picture written by the reference (quantise_transform + the LD slice writer)
      const quant_f ref_qld = sym<quant_f>("ref_quantise_ld");
      std::vector<int> ly((size_t)ph * pw), lu((size_t)ch * cw), lv((size_t)ch * cw), lsb((size_t)ny * nx);
      ref_qld(ry.data(), ph, pw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), ly.data());
      ref_qld(ru.data(), ch, cw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), lu.data());
      ref_qld(rv.data(), ch, cw, qc.data(), ny, nx, rqm.data(), (int)rqm.size(), lv.data());
      const Array2D ldsb = slice_bytes(ny, nx, 5 * (ph * pw + 2 * ch * cw), 1);   // ample: five bytes per coefficient
      ref_sb(ny, nx, 5 * (ph * pw + 2 * ch * cw), 1, lsb.data());
      if (ref_pack(ly.data(), lu.data(), lv.data(), ph, pw, ch, cw, depth, qc.data(), ny, nx, 2, 0, 1, lsb.data(), rbuf.data(), (long)rbuf.size(), &rlen) == 0) {
        const Slices ld = readSlicesLD(rbuf.data(), (size_t)rlen, tf, k, depth, ny, nx, ldsb);
        ref_unpack(rbuf.data(), rlen, ph, pw, ch, cw, depth, ny, nx, 2, 0, 1, lsb.data(), by.data(), bu.data(), bv.data(), bq.data());
        expect(same(ld.yuvCoeffs.y(), by) && same(ld.yuvCoeffs.c1(), bu) && same(ld.yuvCoeffs.c2(), bv) && same(ld.qIndices, bq), "readSlicesLD" + kn);
      } else {
        expect(false, "reference LD writer failed" + kn);
      }
    }
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
  std::printf("%d difference(s)\n", g_fail);
  return g_fail ? 1 : 0;
}
