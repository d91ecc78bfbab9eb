// Stand-in for <boost/atomic.hpp>.
#pragma once
#include <atomic>
namespace boost {
using std::atomic;
}  // namespace boost
