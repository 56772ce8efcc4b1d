// Stand-in for <boost/scoped_ptr.hpp>, written for this repo (NOT Boost source).
#ifndef VC2_SHIM_BOOST_SCOPED_PTR_HPP
#define VC2_SHIM_BOOST_SCOPED_PTR_HPP
namespace boost {
template <class T>
class scoped_ptr {
 public:
  explicit scoped_ptr(T* p = 0) : p_(p) {}
  ~scoped_ptr() { delete p_; }
  void reset(T* p = 0) { if (p != p_) { delete p_; p_ = p; } }
  T& operator*() const { return *p_; }
  T* operator->() const { return p_; }
  T* get() const { return p_; }
  operator bool() const { return p_ != 0; }
 private:
  scoped_ptr(const scoped_ptr&);
  scoped_ptr& operator=(const scoped_ptr&);
  T* p_;
};
}  // namespace boost
#endif
