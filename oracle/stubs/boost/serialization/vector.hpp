// Empty stand-in for <boost/serialization/vector.hpp>: TEST INFRASTRUCTURE ONLY.
#include <vector>
