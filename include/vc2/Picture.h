// vc2/Picture.h - PictureFormat / Picture as in src/Library/Picture.h:23-70, Picture.cpp:49-73.
#ifndef VC2_PICTURE_H
#define VC2_PICTURE_H
#include <iosfwd>
#include "Arrays.h"

namespace vc2 {

enum ColourFormat { CF_UNSET = -1, CF444 = 0, CF422 = 1, CF420 = 2 };   // Picture.h:21, wire values DataUnit.cpp:789

class PictureFormat {
 public:
  PictureFormat() : h_(0), w_(0), cf_(CF_UNSET) {}
  // throws std::invalid_argument for shapes the chroma format cannot subsample (Picture.cpp:38-47)
  PictureFormat(int height, int width, ColourFormat cf);
  int lumaHeight() const { return h_; }
  int lumaWidth() const { return w_; }
  int chromaHeight() const { return cf_ == CF420 ? h_ / 2 : h_; }
  int chromaWidth() const { return cf_ == CF444 ? w_ : w_ / 2; }
  ColourFormat chromaFormat() const { return cf_; }
  bool operator==(const PictureFormat& o) const { return h_ == o.h_ && w_ == o.w_ && cf_ == o.cf_; }
 private:
  int h_, w_;
  ColourFormat cf_;
};

class Picture {
 public:
  Picture() {}
  explicit Picture(const PictureFormat& f)
      : fmt_(f), y_(f.lumaHeight(), f.lumaWidth()), c1_(f.chromaHeight(), f.chromaWidth()), c2_(f.chromaHeight(), f.chromaWidth()) {}
  Picture(const PictureFormat& f, const Array2D& y, const Array2D& c1, const Array2D& c2) : fmt_(f), y_(y), c1_(c1), c2_(c2) {}
  const PictureFormat& format() const { return fmt_; }
  const Array2D& y() const { return y_; }
  const Array2D& c1() const { return c1_; }
  const Array2D& c2() const { return c2_; }
  Array2D& y() { return y_; }
  Array2D& c1() { return c1_; }
  Array2D& c2() { return c2_; }
 private:
  PictureFormat fmt_;
  Array2D y_, c1_, c2_;
};

// clip(Picture, ...) - Picture.cpp:284-292
const Picture clip(const Picture& p, int luma_min, int luma_max, int chroma_min, int chroma_max);

// planar Y, C1, C2 sample IO (Picture.cpp:410-444); luma/chroma depth as pictureio::bitDepth(l, c)
bool readPicture(std::istream& in, Picture& p, int bytes, int luma_depth, int chroma_depth, bool offset_binary = true);
bool writePicture(std::ostream& out, const Picture& p, int bytes, int luma_depth, int chroma_depth, bool offset_binary = true);

}  // namespace vc2
#endif
