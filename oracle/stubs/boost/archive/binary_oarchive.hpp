// No-op stand-in for Boost.Serialization's binary_oarchive: TEST INFRASTRUCTURE ONLY.
// The reference only serialises for its GraphMat-binary file format and for
// `Serializable` message types; neither is on the path the oracle runs.
#ifndef GM_ORACLE_STUB_BOOST_OARCHIVE_
#define GM_ORACLE_STUB_BOOST_OARCHIVE_
#include <ostream>
#ifndef BOOST_SERIALIZATION_SPLIT_MEMBER
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#endif
namespace boost {
namespace serialization { class access {}; }
namespace archive {
class binary_oarchive {
 public:
  explicit binary_oarchive(std::ostream&) {}
  template <class T> binary_oarchive& operator<<(const T&) { return *this; }
  template <class T> binary_oarchive& operator&(const T&) { return *this; }
};
}  // namespace archive
}  // namespace boost
#endif
