// Stand-in for <boost/unordered_set.hpp>.
#pragma once
#include <unordered_set>
namespace boost {
using std::unordered_set;
using std::unordered_multiset;
}  // namespace boost
