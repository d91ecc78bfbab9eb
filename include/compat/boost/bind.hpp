// Stand-in for <boost/bind.hpp>: std::bind, with the _1.._3 placeholders in the
// global namespace the way Boost.Bind's legacy header exposes them.
#pragma once
#include <functional>
namespace boost {
using std::bind;
}  // namespace boost
using std::placeholders::_1;
using std::placeholders::_2;
using std::placeholders::_3;
