/* Oracle stub: included by the reference's c2 file source, nothing used. */
#ifndef ORACLE_STUB_BOOST_ROUND_HPP
#define ORACLE_STUB_BOOST_ROUND_HPP
#include <cmath>
namespace boost { namespace math {
template <class T> inline T round(T v) { return std::round(v); }
template <class T> inline int iround(T v) { return (int)std::lround(v); }
} }
#endif
