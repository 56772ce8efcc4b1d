// Stand-in for <boost/integer/common_factor_rt.hpp>, written for this repo (NOT Boost source).
#ifndef VC2_SHIM_BOOST_COMMON_FACTOR_RT_HPP
#define VC2_SHIM_BOOST_COMMON_FACTOR_RT_HPP
namespace boost { namespace integer {
template <class T>
T gcd(T a, T b) {
  if (a < 0) a = -a;
  if (b < 0) b = -b;
  while (b != 0) { T t = a % b; a = b; b = t; }
  return a;
}
}  // namespace integer
namespace math { using integer::gcd; }
}  // namespace boost
#endif
