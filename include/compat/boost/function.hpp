// Stand-in for <boost/function.hpp>.
#pragma once
#include <functional>
namespace boost {
using std::function;
}  // namespace boost
