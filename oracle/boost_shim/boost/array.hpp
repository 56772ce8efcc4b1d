// Stand-in for <boost/array.hpp>, written for this repo (NOT Boost source).
// Only what bbc/vc2-reference uses: aggregate init, [], size, begin/end, elems.
// Test infrastructure only: lets oracle/build_ref.sh compile the unmodified
// reference in a container that has no system Boost.
#ifndef VC2_SHIM_BOOST_ARRAY_HPP
#define VC2_SHIM_BOOST_ARRAY_HPP
#include <cstddef>
#include <algorithm>
namespace boost {
template <class T, std::size_t N>
struct array {
  T elems[N];
  typedef T value_type;
  typedef T* iterator;
  typedef const T* const_iterator;
  typedef std::size_t size_type;
  T& operator[](size_type i) { return elems[i]; }
  const T& operator[](size_type i) const { return elems[i]; }
  static size_type size() { return N; }
  iterator begin() { return elems; }
  const_iterator begin() const { return elems; }
  iterator end() { return elems + N; }
  const_iterator end() const { return elems + N; }
  T* data() { return elems; }
  const T* data() const { return elems; }
};
template <class T, std::size_t N>
bool operator==(const array<T, N>& a, const array<T, N>& b) {
  return std::equal(a.begin(), a.end(), b.begin());
}
template <class T, std::size_t N>
bool operator!=(const array<T, N>& a, const array<T, N>& b) { return !(a == b); }
}  // namespace boost
#endif
