// Stand-in for <boost/multi_array.hpp> (plus the BBC-patched behaviour the
// reference relies on), written from scratch for this repo.  NOT Boost source.
//
// Purpose: TEST INFRASTRUCTURE ONLY.  It lets oracle/build_ref.sh compile the
// unmodified bbc/vc2-reference sources (which need system Boost, absent from
// this image) so that the reference itself can serve as the parity oracle.
// No codec arithmetic lives here: only 1-D/2-D containers and strided views.
//
// Behaviour taken from how the reference uses the API (enumerated by grep):
//   extents[h][w], indices[Range(a,b,s)][Range()], Array(extents|shape|ranges()),
//   a[y][x], a[index_gen] -> (const) view, view[index_gen] -> sub-view,
//   element-wise view assignment, AUTO-RESIZING array assignment
//   (the BBC patch, /root/reference/src/boost/multi_array.hpp:373-389),
//   ranges() (BBC patch, multi_array_ref.hpp:207-218), resize() keeping the
//   overlapping top-left block (used to crop padding,
//   /root/reference/src/Library/src/WaveletTransform.cpp:340).
#ifndef VC2_SHIM_BOOST_MULTI_ARRAY_HPP
#define VC2_SHIM_BOOST_MULTI_ARRAY_HPP

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <functional>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "boost/array.hpp"

namespace boost {

namespace multi_array_types {
typedef std::ptrdiff_t index;
typedef std::size_t size_type;

class index_range {
 public:
  index_range() : start_(0), finish_(0), stride_(1), all_(true) {}
  index_range(index start, index finish, index stride = 1)
      : start_(start), finish_(finish), stride_(stride), all_(false) {}
  index start(index /*extent*/) const { return all_ ? 0 : start_; }
  index stride() const { return all_ ? 1 : stride_; }
  // number of elements selected from a dimension of the given extent
  index count(index extent) const {
    if (all_) return extent;
    if (finish_ <= start_) return 0;
    return (finish_ - start_ + stride_ - 1) / stride_;
  }
 private:
  index start_, finish_, stride_;
  bool all_;
};
}  // namespace multi_array_types

namespace detail {
namespace multi_array {

using boost::multi_array_types::index;
using boost::multi_array_types::index_range;
using boost::multi_array_types::size_type;

template <int NumRanges, int NumDims>
struct index_gen {
  boost::array<index_range, (NumRanges > 0 ? NumRanges : 1)> ranges_;
  index_gen<NumRanges + 1, NumDims + 1> operator[](const index_range& r) const {
    index_gen<NumRanges + 1, NumDims + 1> g;
    for (int i = 0; i < NumRanges; ++i) g.ranges_[i] = ranges_[i];
    g.ranges_[NumRanges] = r;
    return g;
  }
};

template <std::size_t N>
struct extent_gen {
  struct range {
    index start_, finish_;
    range() : start_(0), finish_(0) {}
    range(index s, index f) : start_(s), finish_(f) {}
    index start() const { return start_; }
    index finish() const { return finish_; }
    size_type size() const { return static_cast<size_type>(finish_ - start_); }
  };
  boost::array<range, (N > 0 ? N : 1)> ranges_;
  extent_gen<N + 1> operator[](index n) const {
    extent_gen<N + 1> g;
    for (std::size_t i = 0; i < N; ++i) {
      g.ranges_[i].start_ = ranges_[i].start_;
      g.ranges_[i].finish_ = ranges_[i].finish_;
    }
    g.ranges_[N].start_ = 0;
    g.ranges_[N].finish_ = n;
    return g;
  }
};

}  // namespace multi_array
}  // namespace detail

namespace {
detail::multi_array::extent_gen<0> extents;
detail::multi_array::index_gen<0, 0> indices;
}  // namespace

template <class T, std::size_t N>
class multi_array;

namespace shim_detail {
using multi_array_types::index;
using multi_array_types::size_type;

// Row proxy of a strided 2-D view
template <class T>
class strided_row {
 public:
  strided_row(T* p, index stride) : p_(p), stride_(stride) {}
  T& operator[](index x) const { return p_[x * stride_]; }
 private:
  T* p_;
  index stride_;
};

// Strided 2-D view; TP is "T" for a mutable view, "const T" for a const view
template <class T, class TP>
class view2d {
 public:
  typedef T element;
  typedef multi_array_types::index index;
  typedef multi_array_types::size_type size_type;

  view2d(TP* base, index h, index w, index sy, index sx) : base_(base), sy_(sy), sx_(sx) {
    shape_[0] = static_cast<size_type>(h);
    shape_[1] = static_cast<size_type>(w);
  }
  // const view from mutable view
  template <class TP2>
  view2d(const view2d<T, TP2>& o)
      : base_(o.base()), sy_(o.stride_y()), sx_(o.stride_x()) {
    shape_[0] = o.shape()[0];
    shape_[1] = o.shape()[1];
  }

  const size_type* shape() const { return shape_; }
  size_type size() const { return shape_[0]; }
  size_type num_elements() const { return shape_[0] * shape_[1]; }
  static size_type num_dimensions() { return 2; }
  TP* base() const { return base_; }
  index stride_y() const { return sy_; }
  index stride_x() const { return sx_; }

  strided_row<TP> operator[](index y) const { return strided_row<TP>(base_ + y * sy_, sx_); }

  view2d operator[](const detail::multi_array::index_gen<2, 2>& g) const {
    const index h = static_cast<index>(shape_[0]), w = static_cast<index>(shape_[1]);
    const index y0 = g.ranges_[0].start(h), x0 = g.ranges_[1].start(w);
    return view2d(base_ + y0 * sy_ + x0 * sx_, g.ranges_[0].count(h), g.ranges_[1].count(w),
                  sy_ * g.ranges_[0].stride(), sx_ * g.ranges_[1].stride());
  }

  detail::multi_array::extent_gen<2> ranges() const {
    detail::multi_array::extent_gen<2> g;
    g.ranges_[0].start_ = 0; g.ranges_[0].finish_ = static_cast<index>(shape_[0]);
    g.ranges_[1].start_ = 0; g.ranges_[1].finish_ = static_cast<index>(shape_[1]);
    return g;
  }

  // element-wise assignment (shapes must agree)
  view2d& operator=(const view2d& o) { copy_from(o); return *this; }
  template <class TP2>
  view2d& operator=(const view2d<T, TP2>& o) { copy_from(o); return *this; }
  view2d& operator=(const boost::multi_array<T, 2>& o);

 private:
  template <class V>
  void copy_from(const V& o) {
    assert(o.shape()[0] == shape_[0] && o.shape()[1] == shape_[1]);
    for (size_type y = 0; y < shape_[0]; ++y)
      for (size_type x = 0; x < shape_[1]; ++x)
        base_[static_cast<index>(y) * sy_ + static_cast<index>(x) * sx_] = o[y][x];
  }
  TP* base_;
  index sy_, sx_;
  size_type shape_[2];
};
}  // namespace shim_detail

// ---------------------------------------------------------------- 1-D array
template <class T>
class multi_array<T, 1> {
 public:
  typedef T element;
  typedef multi_array_types::index index;
  typedef multi_array_types::size_type size_type;
  typedef detail::multi_array::extent_gen<1> extent_gen1;

  multi_array() { shape_[0] = 0; }
  explicit multi_array(const extent_gen1& e) : data_(e.ranges_[0].size()) { shape_[0] = e.ranges_[0].size(); }
  multi_array(const multi_array& o) : data_(o.data_) { shape_[0] = o.shape_[0]; }
  multi_array& operator=(const multi_array& o) {
    data_ = o.data_;
    shape_[0] = o.shape_[0];
    return *this;
  }
  const size_type* shape() const { return shape_; }
  size_type size() const { return shape_[0]; }
  size_type num_elements() const { return shape_[0]; }
  static size_type num_dimensions() { return 1; }
  T* data() { return data_.empty() ? 0 : &data_[0]; }
  const T* data() const { return data_.empty() ? 0 : &data_[0]; }
  T& operator[](index i) { return data_[static_cast<size_type>(i)]; }
  const T& operator[](index i) const { return data_[static_cast<size_type>(i)]; }
  extent_gen1 ranges() const {
    extent_gen1 g;
    g.ranges_[0].start_ = 0;
    g.ranges_[0].finish_ = static_cast<index>(shape_[0]);
    return g;
  }
  void resize(const extent_gen1& e) {
    data_.resize(e.ranges_[0].size());
    shape_[0] = e.ranges_[0].size();
  }
 private:
  std::vector<T> data_;
  size_type shape_[1];
};

// ---------------------------------------------------------------- 2-D array
template <class T>
class multi_array<T, 2> {
 public:
  typedef T element;
  typedef multi_array_types::index index;
  typedef multi_array_types::size_type size_type;
  typedef detail::multi_array::extent_gen<2> extent_gen2;
  typedef shim_detail::view2d<T, T> view_type;
  typedef shim_detail::view2d<T, const T> const_view_type;
  template <std::size_t D> struct array_view { typedef view_type type; };
  template <std::size_t D> struct const_array_view { typedef const_view_type type; };

  multi_array() { shape_[0] = shape_[1] = 0; }
  explicit multi_array(const extent_gen2& e) { init(e.ranges_[0].size(), e.ranges_[1].size()); }
  explicit multi_array(const boost::array<index, 2>& s) {
    init(static_cast<size_type>(s[0]), static_cast<size_type>(s[1]));
  }
  multi_array(const multi_array& o) : data_(o.data_) {
    shape_[0] = o.shape_[0];
    shape_[1] = o.shape_[1];
  }
  multi_array(const view_type& v) { init(v.shape()[0], v.shape()[1]); fill_from(v); }
  multi_array(const const_view_type& v) { init(v.shape()[0], v.shape()[1]); fill_from(v); }

  // BBC-patched semantics: assignment resizes the destination when shapes differ
  multi_array& operator=(const multi_array& o) {
    if (this != &o) {
      data_ = o.data_;
      shape_[0] = o.shape_[0];
      shape_[1] = o.shape_[1];
    }
    return *this;
  }
  multi_array& operator=(const view_type& v) { assign_view(v); return *this; }
  multi_array& operator=(const const_view_type& v) { assign_view(v); return *this; }

  const size_type* shape() const { return shape_; }
  size_type size() const { return shape_[0]; }
  size_type num_elements() const { return shape_[0] * shape_[1]; }
  static size_type num_dimensions() { return 2; }
  T* data() { return data_.empty() ? 0 : &data_[0]; }
  const T* data() const { return data_.empty() ? 0 : &data_[0]; }

  T* operator[](index y) { return data() + static_cast<size_type>(y) * shape_[1]; }
  const T* operator[](index y) const { return data() + static_cast<size_type>(y) * shape_[1]; }

  view_type operator[](const detail::multi_array::index_gen<2, 2>& g) {
    return whole()[g];
  }
  const_view_type operator[](const detail::multi_array::index_gen<2, 2>& g) const {
    return const_whole()[g];
  }

  extent_gen2 ranges() const {
    extent_gen2 g;
    g.ranges_[0].start_ = 0; g.ranges_[0].finish_ = static_cast<index>(shape_[0]);
    g.ranges_[1].start_ = 0; g.ranges_[1].finish_ = static_cast<index>(shape_[1]);
    return g;
  }

  // resize keeping the overlapping top-left block
  void resize(const extent_gen2& e) { do_resize(e.ranges_[0].size(), e.ranges_[1].size()); }
  void resize(const boost::array<index, 2>& s) {
    do_resize(static_cast<size_type>(s[0]), static_cast<size_type>(s[1]));
  }

 private:
  void init(size_type h, size_type w) {
    shape_[0] = h;
    shape_[1] = w;
    data_.assign(h * w, T());
  }
  view_type whole() {
    return view_type(data(), static_cast<index>(shape_[0]), static_cast<index>(shape_[1]),
                     static_cast<index>(shape_[1]), 1);
  }
  const_view_type const_whole() const {
    return const_view_type(data(), static_cast<index>(shape_[0]), static_cast<index>(shape_[1]),
                           static_cast<index>(shape_[1]), 1);
  }
  template <class V>
  void fill_from(const V& v) {
    for (size_type y = 0; y < shape_[0]; ++y)
      for (size_type x = 0; x < shape_[1]; ++x) data_[y * shape_[1] + x] = v[y][x];
  }
  template <class V>
  void assign_view(const V& v) {
    // the view may alias this array's own storage: go through a temporary
    multi_array tmp(v);
    data_.swap(tmp.data_);
    shape_[0] = tmp.shape_[0];
    shape_[1] = tmp.shape_[1];
  }
  void do_resize(size_type h, size_type w) {
    if (h == shape_[0] && w == shape_[1]) return;
    std::vector<T> nd(h * w, T());
    const size_type ch = std::min(h, shape_[0]), cw = std::min(w, shape_[1]);
    for (size_type y = 0; y < ch; ++y)
      for (size_type x = 0; x < cw; ++x) nd[y * w + x] = data_[y * shape_[1] + x];
    data_.swap(nd);
    shape_[0] = h;
    shape_[1] = w;
  }
  std::vector<T> data_;
  size_type shape_[2];
};

namespace shim_detail {
template <class T, class TP>
view2d<T, TP>& view2d<T, TP>::operator=(const boost::multi_array<T, 2>& o) {
  copy_from(o);
  return *this;
}
}  // namespace shim_detail

}  // namespace boost

#endif  // VC2_SHIM_BOOST_MULTI_ARRAY_HPP
