// Stand-in for <boost/unordered_map.hpp>.
#pragma once
#include <unordered_map>
namespace boost {
using std::unordered_map;
using std::unordered_multimap;
}  // namespace boost
