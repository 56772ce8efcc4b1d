// vc2/Arrays.h - the container types of the hot path, same observable layout as the reference's
// Array1D / Array2D (src/Library/Arrays.h:28-31: boost::multi_array<int, N>): row-major, contiguous,
// `int` elements, a[y][x] indexing, shape()[0] = rows, shape()[1] = columns, value semantics and
// auto-resizing assignment.  Own code (no Boost); only what the hot-path callers use.
#ifndef VC2_ARRAYS_H
#define VC2_ARRAYS_H
#include <cstddef>
#include <iosfwd>
#include <vector>

namespace vc2 {

typedef std::vector<int> Array1D;

class Array2D {
 public:
  Array2D() { dims_[0] = dims_[1] = 0; }
  Array2D(int rows, int cols) : v_((size_t)rows * cols, 0) { dims_[0] = rows; dims_[1] = cols; }
  void resize(int rows, int cols) { v_.assign((size_t)rows * cols, 0); dims_[0] = rows; dims_[1] = cols; }
  const size_t* shape() const { return dims_; }          // shape()[0] rows, shape()[1] columns
  int* data() { return v_.data(); }
  const int* data() const { return v_.data(); }
  size_t num_elements() const { return v_.size(); }
  int* operator[](int y) { return v_.data() + (size_t)y * dims_[1]; }
  const int* operator[](int y) const { return v_.data() + (size_t)y * dims_[1]; }
  bool operator==(const Array2D& o) const { return dims_[0] == o.dims_[0] && dims_[1] == o.dims_[1] && v_ == o.v_; }
 private:
  std::vector<int> v_;
  size_t dims_[2];
};

// clip(Array2D, min, max) - src/Library/src/Arrays.cpp:41-53
const Array2D clip(const Array2D& values, int min_value, int max_value);

// Raw sample IO of one plane (src/Library/src/Arrays.cpp:333-426 with the stream state set by
// arrayio::wordWidth / left_justified / offset_binary / bitDepth): big-endian words of `bytes` bytes,
// value = (word >> (8*bytes - depth)) - 2^(depth-1) when offset_binary, two's complement when !offset_binary.
struct SampleFormat {
  int bytes;          // 1..4
  int depth;          // significant bits
  bool left_justified;
  bool offset_binary; // false: signed two's complement
};
bool readArray(std::istream& in, Array2D& a, const SampleFormat& f);     // false on short read (failbit)
bool writeArray(std::ostream& out, const Array2D& a, const SampleFormat& f);

}  // namespace vc2
#endif
