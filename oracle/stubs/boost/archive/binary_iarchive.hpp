// No-op stand-in for Boost.Serialization's binary_iarchive: TEST INFRASTRUCTURE ONLY.
#ifndef GM_ORACLE_STUB_BOOST_IARCHIVE_
#define GM_ORACLE_STUB_BOOST_IARCHIVE_
#include <istream>
#include "boost/archive/binary_oarchive.hpp"
namespace boost {
namespace archive {
class binary_iarchive {
 public:
  explicit binary_iarchive(std::istream&) {}
  template <class T> binary_iarchive& operator>>(T&) { return *this; }
  template <class T> binary_iarchive& operator&(T&) { return *this; }
};
}  // namespace archive
}  // namespace boost
#endif
